// Explicit instantiations of the fused path kernel for HjmModel12 (TQF_MODEL_HJM).
#include "tqf_paths_kernel.cuh"

namespace tqf {
template int launch_path_kernel<HjmModel12<double>>(int, bool, int, int, size_t, const KParams<double>&,
                                             cudaStream_t, int*);
template int launch_path_kernel<HjmModel12<float>>(int, bool, int, int, size_t, const KParams<float>&,
                                            cudaStream_t, int*);
}  // namespace tqf
