// Explicit instantiations of the fused path kernel for HestonQeModel.
#include "tqf_paths_kernel.cuh"

namespace tqf {
template int launch_path_kernel<HestonQeModel<double>>(int, bool, int, int, size_t,
                                                       const KParams<double>&, cudaStream_t,
                                                       int*);
template int launch_path_kernel<HestonQeModel<float>>(int, bool, int, int, size_t,
                                                      const KParams<float>&, cudaStream_t, int*);
}  // namespace tqf
