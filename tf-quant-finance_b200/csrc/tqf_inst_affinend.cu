// Explicit instantiations of the fused path kernel for the generic affine
// models of dimension 2..4.
#include "tqf_paths_kernel.cuh"

namespace tqf {
#define TQF_INST(M)                                                                          \
  template int launch_path_kernel<M<double>>(int, bool, int, int, size_t,                    \
                                             const KParams<double>&, cudaStream_t, int*);    \
  template int launch_path_kernel<M<float>>(int, bool, int, int, size_t,                     \
                                            const KParams<float>&, cudaStream_t, int*);
TQF_INST(AffineModel2D)
TQF_INST(AffineModel3D)
TQF_INST(AffineModel4D)
#undef TQF_INST
}  // namespace tqf
