// Explicit instantiations of the fused path kernel for MilsteinAffine1FModel (one
// translation unit per model so that they compile in parallel).
#include "tqf_paths_kernel.cuh"

namespace tqf {
template int launch_path_kernel<MilsteinAffine1FModel<double>>(int, bool, int, int, size_t,
                                            const KParams<double>&, cudaStream_t, int*);
template int launch_path_kernel<MilsteinAffine1FModel<float>>(int, bool, int, int, size_t,
                                           const KParams<float>&, cudaStream_t, int*);
}  // namespace tqf
