// Explicit instantiations of the fused path kernel for TangentHestonModel (one
// translation unit per model so that they compile in parallel).
#include "tqf_paths_kernel.cuh"

namespace tqf {
template int launch_path_kernel<TangentHestonModel<double>>(int, bool, int, int, size_t,
                                            const KParams<double>&, cudaStream_t, int*);
template int launch_path_kernel<TangentHestonModel<float>>(int, bool, int, int, size_t,
                                           const KParams<float>&, cudaStream_t, int*);
}  // namespace tqf
