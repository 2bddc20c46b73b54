// Stand-alone generators of libtqf: Philox raw / normal fills and Sobol fills.
// They exist so that the uint32 streams can be checked bit for bit against
// the oracle, and they back `tff_b200.math.random` (stateless_normal,
// sobol.sample, mv_normal_sample).  The path kernels use the same device
// functions (tqf_common.cuh) in registers.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include <cmath>
#include "tqf_common.cuh"

namespace tqf {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return TQF_ERR_CUDA;
}

namespace {
struct DevBlockCache {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_blocks;      // (device, bytes) -> block
  std::map<void*, std::pair<int, size_t>> live;                  // block -> (device, bytes)
  size_t cached_bytes = 0;
  // blocks up to kSlabBlockMax are carved out of 16 MB slabs (one cudaMalloc per slab):
  // the first plans of a process do not pay one cudaMalloc (~0.3-1 ms) per buffer either
  struct Slab {
    char* base;
    size_t used;
  };
  std::map<int, Slab> slab;                                      // device -> current slab
  std::set<void*> carved;                                        // blocks that live inside a slab
};
constexpr size_t kSlabBytes = 16ull << 20, kSlabBlockMax = 4ull << 20;
DevBlockCache& dev_cache() {
  static DevBlockCache* c = new DevBlockCache();   // never destroyed: outlives the CUDA context teardown
  return *c;
}
constexpr size_t kDevCacheLimit = 256ull << 20;
}  // namespace

int dev_alloc(void** out, size_t bytes) {
  size_t size = 512;
  while (size < bytes && size < (64ull << 20)) size <<= 1;
  if (size < bytes) size = (bytes + (16ull << 20) - 1) & ~((16ull << 20) - 1);   // large: 16 MB steps
  int dev = 0;
  TQF_CUDA_OK(cudaGetDevice(&dev));
  DevBlockCache& c = dev_cache();
  {
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.free_blocks.find({dev, size});
    if (it != c.free_blocks.end()) {
      *out = it->second;
      c.free_blocks.erase(it);
      c.cached_bytes -= size;
      c.live[*out] = {dev, size | (c.carved.count(*out) ? 1 : 0)};
      return TQF_OK;
    }
  }
  if (size <= kSlabBlockMax) {
    std::lock_guard<std::mutex> lock(c.mu);
    DevBlockCache::Slab& sl = c.slab[dev];
    if (sl.base == nullptr || sl.used + size > kSlabBytes) {
      void* base = nullptr;
      TQF_CUDA_OK(cudaMalloc(&base, kSlabBytes));     // the rest of an old slab stays with its blocks
      sl.base = static_cast<char*>(base);
      sl.used = 0;
    }
    *out = sl.base + sl.used;
    sl.used += size;
    c.carved.insert(*out);
    c.live[*out] = {dev, size | 1};                   // bit 0: carved from a slab, never cudaFree'd
    return TQF_OK;
  }
  void* p = nullptr;
  TQF_CUDA_OK(cudaMalloc(&p, size));
  std::lock_guard<std::mutex> lock(c.mu);
  c.live[p] = {dev, size};
  *out = p;
  return TQF_OK;
}

void dev_release(void* const* ptrs, int count) {
  bool any = false;
  for (int i = 0; i < count; ++i) any = any || ptrs[i] != nullptr;
  if (!any) return;
  // what cudaFree did implicitly: nothing in flight may still read the blocks
  cudaDeviceSynchronize();
  DevBlockCache& c = dev_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  for (int i = 0; i < count; ++i) {
    void* p = ptrs[i];
    if (!p) continue;
    auto it = c.live.find(p);
    if (it == c.live.end()) {          // not one of ours (allocated with cudaMalloc)
      cudaFree(p);
      continue;
    }
    const bool carved = (it->second.second & 1) != 0;
    const std::pair<int, size_t> key = {it->second.first, it->second.second & ~static_cast<size_t>(1)};
    c.live.erase(it);
    if (!carved && c.cached_bytes + key.second > kDevCacheLimit) {
      cudaFree(p);
    } else {
      c.free_blocks.insert({key, p});
      c.cached_bytes += key.second;
    }
  }
}

// Device-global copy of the {T, 1/c} table of the table logarithm
// (tqf_math.cuh); uploaded once per device and kept for the process lifetime.
// Two tables back to back: the one of the central ndtri branch (T carries the
// -MID offset) and the MID-free one behind the Box-Muller logarithm of the Philox
// path (kernels address it at +2 * TQF_LOGTAB_COUNT doubles).
static const double kLogTabHost[4 * TQF_LOGTAB_COUNT] = {
#include "tqf_logtab.inc"
#include "tqf_logtab0.inc"
};

int device_logtab(const double** out) {
  static std::mutex mu;
  static const double* tables[64] = {nullptr};
  int dev = 0;
  TQF_CUDA_OK(cudaGetDevice(&dev));
  TQF_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  std::lock_guard<std::mutex> lock(mu);
  if (!tables[dev]) {
    double* d = nullptr;
    TQF_CUDA_OK(cudaMalloc(&d, sizeof(kLogTabHost)));
    cudaError_t e = cudaMemcpy(d, kLogTabHost, sizeof(kLogTabHost), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(d);
      return cuda_fail(e, "cudaMemcpy(log table)");
    }
    tables[dev] = d;
  }
  *out = tables[dev];
  return TQF_OK;
}

static const float kNdtriF32TabHost[] = {
#include "tqf_ndtri_f32_tab.inc"
};

int device_ndtri_f32_tab(const float** out) {
  static std::mutex mu;
  static const float* tables[64] = {nullptr};
  int dev = 0;
  TQF_CUDA_OK(cudaGetDevice(&dev));
  TQF_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  std::lock_guard<std::mutex> lock(mu);
  if (!tables[dev]) {
    float* d = nullptr;
    TQF_CUDA_OK(cudaMalloc(&d, sizeof(kNdtriF32TabHost)));
    cudaError_t e = cudaMemcpy(d, kNdtriF32TabHost, sizeof(kNdtriF32TabHost), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(d);
      return cuda_fail(e, "cudaMemcpy(ndtri table)");
    }
    tables[dev] = d;
  }
  *out = tables[dev];
  return TQF_OK;
}

// ------------------------------------------------------------- kernels ----
__global__ void philox_raw_kernel(PhiloxKey key, PhiloxCtr ctr, uint64_t first_group,
                                  uint64_t num_groups, uint4* __restrict__ out) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       g < num_groups; g += stride) {
    out[g] = philox_group(ctr, key, first_group + g);
  }
}

// One thread per Philox group; writes the elements of the group that fall into
// [first_element, first_element + n).
__global__ void philox_normal_f64_kernel(PhiloxKey key, PhiloxCtr ctr,
                                         uint64_t first_element, uint64_t n,
                                         double* __restrict__ out) {
  const uint64_t g0 = first_element >> 1;
  const uint64_t g1 = (first_element + n + 1) >> 1;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t g = g0 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       g < g1; g += stride) {
    const uint4 w = philox_group(ctr, key, g);
    double a, b;
    box_muller(w.x, w.y, w.z, w.w, &a, &b);
    const uint64_t e = g << 1;
    if (e >= first_element && e < first_element + n) out[e - first_element] = a;
    if (e + 1 >= first_element && e + 1 < first_element + n) out[e + 1 - first_element] = b;
  }
}

__global__ void philox_normal_f32_kernel(PhiloxKey key, PhiloxCtr ctr,
                                         uint64_t first_element, uint64_t n,
                                         float* __restrict__ out) {
  const uint64_t g0 = first_element >> 2;
  const uint64_t g1 = (first_element + n + 3) >> 2;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t g = g0 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       g < g1; g += stride) {
    const uint4 w = philox_group(ctr, key, g);
    float v[4];
    box_muller(w.x, w.y, &v[0], &v[1]);
    box_muller(w.z, w.w, &v[2], &v[3]);
    const uint64_t e = g << 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (e + k >= first_element && e + k < first_element + n) out[e + k - first_element] = v[k];
    }
  }
}

// tf.random.stateless_uniform / tf.random.uniform on [0, 1)
// (random_distributions.h: UniformDistribution<PhiloxRandom, T>): float32 takes
// four Uint32ToFloat per group, float64 two Uint64ToDouble.
__global__ void philox_uniform_f64_kernel(PhiloxKey key, PhiloxCtr ctr, uint64_t first_element,
                                          uint64_t n, double* __restrict__ out) {
  const uint64_t g0 = first_element >> 1;
  const uint64_t g1 = (first_element + n + 1) >> 1;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t g = g0 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       g < g1; g += stride) {
    const uint4 w = philox_group(ctr, key, g);
    const double a = uint64_to_double(w.x, w.y), b = uint64_to_double(w.z, w.w);
    const uint64_t e = g << 1;
    if (e >= first_element && e < first_element + n) out[e - first_element] = a;
    if (e + 1 >= first_element && e + 1 < first_element + n) out[e + 1 - first_element] = b;
  }
}

__global__ void philox_uniform_f32_kernel(PhiloxKey key, PhiloxCtr ctr, uint64_t first_element,
                                          uint64_t n, float* __restrict__ out) {
  const uint64_t g0 = first_element >> 2;
  const uint64_t g1 = (first_element + n + 3) >> 2;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t g = g0 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       g < g1; g += stride) {
    const uint4 w = philox_group(ctr, key, g);
    const float v[4] = {uint32_to_float(w.x), uint32_to_float(w.y), uint32_to_float(w.z),
                        uint32_to_float(w.w)};
    const uint64_t e = g << 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (e + k >= first_element && e + k < first_element + n) out[e + k - first_element] = v[k];
    }
  }
}

// Sobol fill: one thread per (row, dim) element, rows fastest within a warp
// would make stores strided, so threads run over the flattened [count][dim]
// output (coalesced stores); the XOR walks the set bits of the index.
template <int KIND, typename Out>
__global__ void sobol_fill_kernel(const uint32_t* __restrict__ v_table, int dim,
                                  uint64_t first_index, uint64_t count,
                                  int num_digits, const double* __restrict__ logtab,
                                  Out* __restrict__ out) {
  const uint64_t total = count * static_cast<uint64_t>(dim);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       e < total; e += stride) {
    const uint64_t row = e / dim;
    const int d = static_cast<int>(e - row * dim);
    uint32_t i = static_cast<uint32_t>(first_index + row);
    const uint32_t* v = v_table + static_cast<size_t>(d) * 32;
    uint32_t x = 0;
    while (i) {
      const int b = __ffs(i) - 1;
      x ^= __ldg(v + b);
      i &= i - 1;
    }
    if constexpr (KIND == 0) {
      out[e] = static_cast<Out>(x >> (32 - num_digits));
    } else if constexpr (KIND == 1) {
      out[e] = RealTraits<Out>::sobol_uniform(x);
    } else {
      out[e] = ndtri(RealTraits<Out>::sobol_uniform(x), logtab);
    }
  }
}

// Non-randomized Halton sequence in the reference's floating-point arithmetic
// (halton_impl.py:250-288): digit j of index i in base p is
// floor(i / p^j) mod p, the point is sum_j (digit_j / p) / p^j, everything in
// `Real`.  One thread per (row, dim) element of the [count][dim] output.
template <int KIND, typename Real>
__global__ void halton_fill_kernel(const Real* __restrict__ weights, const int* __restrict__ sizes,
                                   const Real* __restrict__ radixes, int dim, int max_size,
                                   uint64_t first_index, uint64_t count,
                                   const double* __restrict__ logtab, Real* __restrict__ out) {
  const uint64_t total = count * static_cast<uint64_t>(dim);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       e < total; e += stride) {
    const uint64_t row = e / dim;
    const int d = static_cast<int>(e - row * dim);
    // indices = cast(sequence_indices, dtype) + 1   (halton_impl.py:404-411)
    const Real idx = static_cast<Real>(first_index + row) + Real(1);
    const Real p = radixes[d];
    const Real* w = weights + static_cast<size_t>(d) * max_size;
    const int n = sizes[d];
    Real sum = 0;
    for (int j = 0; j < n; ++j) {
      const Real wj = w[j];
      Real c = floor(idx / wj);
      c = fmod(c, p);
      c = c / p;
      sum += c / wj;
    }
    if constexpr (KIND == 1) {
      out[e] = sum;
    } else {
      out[e] = ndtri(sum, logtab);
    }
  }
}

// Owen-randomized Halton point (halton_impl.py:290-322): digit j of axis d goes
// through permutation `perms[j][offset_d ..]` of range(p_d) before it is scaled; all
// `num_coeffs` digit positions are looked up as the reference does, the ones beyond
// the axis' own size are then masked, and the random tail `zero_correction[d]` is added.
template <int KIND, typename Real>
__global__ void halton_randomized_kernel(const Real* __restrict__ weights,
                                         const int* __restrict__ sizes,
                                         const Real* __restrict__ radixes,
                                         const int* __restrict__ radix_offsets,
                                         const int* __restrict__ perms, int radix_sum,
                                         const Real* __restrict__ zero_correction, int dim,
                                         int max_size, uint64_t first_index, uint64_t count,
                                         const double* __restrict__ logtab,
                                         Real* __restrict__ out) {
  const uint64_t total = count * static_cast<uint64_t>(dim);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t e = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       e < total; e += stride) {
    const uint64_t row = e / dim;
    const int d = static_cast<int>(e - row * dim);
    const Real idx = static_cast<Real>(first_index + row) + Real(1);
    const Real p = radixes[d];
    const Real* w = weights + static_cast<size_t>(d) * max_size;
    const int n = sizes[d];
    const int* perm = perms + radix_offsets[d];
    Real sum = 0;
    for (int j = 0; j < n; ++j) {
      const Real wj = w[j];
      Real c = floor(idx / wj);
      c = fmod(c, p);
      c = static_cast<Real>(perm[static_cast<size_t>(j) * radix_sum + static_cast<int>(c)]);
      c = c / p;
      sum += c / wj;
    }
    sum += zero_correction[d];
    if constexpr (KIND == 1) {
      out[e] = sum;
    } else {
      out[e] = ndtri(sum, logtab);
    }
  }
}

__global__ void math_eval_kernel(int fn, const double* __restrict__ in, double* __restrict__ out,
                                 uint64_t n, const double* __restrict__ logtab) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const double x = in[i];
    double r, s, c;
    switch (fn) {
      case 0: r = fm::log_pos(x); break;
      case 1: r = fm::sqrt_pos(x); break;
      case 2: r = ndtri(x, logtab); break;
      case 3: fm::sincos_2pi(x, &s, &c); r = s; break;
      default: fm::sincos_2pi(x, &s, &c); r = c; break;
    }
    out[i] = r;
  }
}

static int grid_for(uint64_t work_items, int block) {
  uint64_t blocks = (work_items + block - 1) / block;
  const uint64_t cap = static_cast<uint64_t>(kSMs) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return static_cast<int>(blocks);
}

// Host: left-aligned device table from m[dim][32].
int upload_sobol_table(const int32_t* direction_numbers, int dim, uint32_t** out_dev,
                       cudaStream_t stream) {
  std::vector<uint32_t> v(static_cast<size_t>(dim) * 32);
  for (int d = 0; d < dim; ++d) {
    for (int b = 0; b < 32; ++b) {
      const uint32_t m = static_cast<uint32_t>(direction_numbers[static_cast<size_t>(d) * 32 + b]);
      v[static_cast<size_t>(d) * 32 + b] = m << (31 - b);
    }
  }
  uint32_t* dev = nullptr;
  {
    void* p = nullptr;
    const int rc = dev_alloc(&p, v.size() * sizeof(uint32_t));
    if (rc != TQF_OK) return rc;
    dev = static_cast<uint32_t*>(p);
  }
  cudaError_t e = cudaMemcpyAsync(dev, v.data(), v.size() * sizeof(uint32_t),
                                  cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // v is a local
  if (e != cudaSuccess) {
    void* p = dev;
    dev_release(&p, 1);
    return cuda_fail(e, "upload_sobol_table");
  }
  *out_dev = dev;
  return TQF_OK;
}

int sobol_num_digits(uint64_t skip, uint64_t num_results) {
  // ceil(log2(skip + num_results + 1)) (sobol_impl.py:118-123).
  const uint64_t max_index = skip + num_results + 1;
  int nd = 0;
  while ((1ull << nd) < max_index) ++nd;
  return nd;
}

}  // namespace tqf

using namespace tqf;

extern "C" {

const char* tqf_last_error(void) { return g_last_error.c_str(); }

int tqf_version(void) { return TQF_VERSION; }

int tqf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int tqf_abi_sizes(int32_t out[4]) {
  TQF_REQUIRE(out, "null argument");
  out[0] = static_cast<int32_t>(sizeof(tqf_rng_desc));
  out[1] = static_cast<int32_t>(sizeof(tqf_model_desc));
  out[2] = static_cast<int32_t>(sizeof(tqf_payoff_desc));
  out[3] = static_cast<int32_t>(sizeof(tqf_lsm_desc));
  return TQF_OK;
}

int tqf_philox_stateless_key_counter(const int64_t seed[2], uint32_t key[2],
                                     uint32_t counter[4]) {
  TQF_REQUIRE(seed && key && counter, "null argument");
  const uint64_t s0 = static_cast<uint64_t>(seed[0]);
  const uint64_t s1 = static_cast<uint64_t>(seed[1]);
  const uint4 mix = philox4x32_10(static_cast<uint32_t>(s0), static_cast<uint32_t>(s0 >> 32),
                                  static_cast<uint32_t>(s1), static_cast<uint32_t>(s1 >> 32),
                                  0x3ec8f720u, 0x02461e29u);
  key[0] = mix.x;
  key[1] = mix.y;
  counter[0] = 0;
  counter[1] = 0;
  counter[2] = mix.z;
  counter[3] = mix.w;
  return TQF_OK;
}

int tqf_philox_stateful_key_counter(int64_t op_seed, uint32_t key[2], uint32_t counter[4]) {
  TQF_REQUIRE(key && counter, "null argument");
  const int64_t kMaxInt32 = 2147483647;
  int64_t a = 87654321 % kMaxInt32;        // DEFAULT_GRAPH_SEED
  int64_t b = ((op_seed % kMaxInt32) + kMaxInt32) % kMaxInt32;  // python %
  if (a == 0 && b == 0) b = kMaxInt32;
  key[0] = static_cast<uint32_t>(a);
  key[1] = static_cast<uint32_t>(static_cast<uint64_t>(a) >> 32);
  counter[0] = 0;
  counter[1] = 0;
  counter[2] = static_cast<uint32_t>(b);
  counter[3] = static_cast<uint32_t>(static_cast<uint64_t>(b) >> 32);
  return TQF_OK;
}

int tqf_philox_raw_fill(const uint32_t key[2], const uint32_t counter[4],
                        uint64_t first_group, uint64_t num_groups, uint32_t* out_dev,
                        void* stream) {
  TQF_NVTX("tqf_philox_raw_fill");
  TQF_REQUIRE(key && counter, "null key/counter");
  if (num_groups == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  const PhiloxKey k{key[0], key[1]};
  const PhiloxCtr c{counter[0], counter[1], counter[2], counter[3]};
  philox_raw_kernel<<<grid_for(num_groups, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      k, c, first_group, num_groups, reinterpret_cast<uint4*>(out_dev));
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

int tqf_philox_normal_fill(const uint32_t key[2], const uint32_t counter[4],
                           uint64_t first_element, uint64_t num_elements, int dtype,
                           void* out_dev, void* stream) {
  TQF_NVTX("tqf_philox_normal_fill");
  TQF_REQUIRE(key && counter, "null key/counter");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "dtype must be TQF_F32 or TQF_F64");
  if (num_elements == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  const PhiloxKey k{key[0], key[1]};
  const PhiloxCtr c{counter[0], counter[1], counter[2], counter[3]};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == TQF_F64) {
    philox_normal_f64_kernel<<<grid_for(num_elements / 2 + 1, 256), 256, 0, s>>>(
        k, c, first_element, num_elements, static_cast<double*>(out_dev));
  } else {
    philox_normal_f32_kernel<<<grid_for(num_elements / 4 + 1, 256), 256, 0, s>>>(
        k, c, first_element, num_elements, static_cast<float*>(out_dev));
  }
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

}  // extern "C"

template <typename Real>
static int halton_fill_impl(const double* weights, const int32_t* sizes, const int32_t* radixes,
                            int dim, int max_size, uint64_t first_index, uint64_t count, int kind,
                            void* out_dev, cudaStream_t s) {
  std::vector<Real> w(static_cast<size_t>(dim) * max_size), r(dim);
  for (size_t i = 0; i < w.size(); ++i) w[i] = static_cast<Real>(weights[i]);
  for (int d = 0; d < dim; ++d) r[d] = static_cast<Real>(radixes[d]);
  Real *w_dev = nullptr, *r_dev = nullptr;
  int* n_dev = nullptr;
  const double* logtab = nullptr;
  int rc = device_logtab(&logtab);
  if (rc != TQF_OK) return rc;
  cudaError_t e = cudaMalloc(&w_dev, w.size() * sizeof(Real));
  if (e == cudaSuccess) e = cudaMalloc(&r_dev, r.size() * sizeof(Real));
  if (e == cudaSuccess) e = cudaMalloc(&n_dev, dim * sizeof(int));
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(w_dev, w.data(), w.size() * sizeof(Real), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(r_dev, r.data(), r.size() * sizeof(Real), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(n_dev, sizes, dim * sizeof(int), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    const int grid = grid_for(count * static_cast<uint64_t>(dim), 256);
    if (kind == 1)
      halton_fill_kernel<1, Real><<<grid, 256, 0, s>>>(w_dev, n_dev, r_dev, dim, max_size,
                                                       first_index, count, logtab,
                                                       static_cast<Real*>(out_dev));
    else
      halton_fill_kernel<2, Real><<<grid, 256, 0, s>>>(w_dev, n_dev, r_dev, dim, max_size,
                                                       first_index, count, logtab,
                                                       static_cast<Real*>(out_dev));
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // the host staging buffers go away
  }
  cudaFree(w_dev);
  cudaFree(r_dev);
  cudaFree(n_dev);
  if (e != cudaSuccess) return cuda_fail(e, "tqf_halton_fill");
  return TQF_OK;
}

template <typename Real>
static int halton_randomized_impl(const double* weights, const int32_t* sizes,
                                  const int32_t* radixes, int dim, int max_size,
                                  const int32_t* perms_dev, const double* zero_correction,
                                  uint64_t first_index, uint64_t count, int kind, void* out_dev,
                                  cudaStream_t s) {
  std::vector<Real> w(static_cast<size_t>(dim) * max_size), r(dim), z(dim);
  std::vector<int> offs(dim);
  int radix_sum = 0;
  for (size_t i = 0; i < w.size(); ++i) w[i] = static_cast<Real>(weights[i]);
  for (int d = 0; d < dim; ++d) {
    r[d] = static_cast<Real>(radixes[d]);
    z[d] = static_cast<Real>(zero_correction[d]);
    offs[d] = radix_sum;
    radix_sum += radixes[d];
  }
  const double* logtab = nullptr;
  int rc = device_logtab(&logtab);
  if (rc != TQF_OK) return rc;
  // one staging allocation: weights | radixes | zero correction | sizes | offsets
  const size_t nreal = w.size() + 2 * static_cast<size_t>(dim);
  const size_t bytes = nreal * sizeof(Real) + 2 * static_cast<size_t>(dim) * sizeof(int);
  unsigned char* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "tqf_halton_randomized_fill");
  Real* w_dev = reinterpret_cast<Real*>(dev);
  Real* r_dev = w_dev + w.size();
  Real* z_dev = r_dev + dim;
  int* n_dev = reinterpret_cast<int*>(z_dev + dim);
  int* o_dev = n_dev + dim;
  e = cudaMemcpyAsync(w_dev, w.data(), w.size() * sizeof(Real), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r_dev, r.data(), dim * sizeof(Real), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(z_dev, z.data(), dim * sizeof(Real), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(n_dev, sizes, dim * sizeof(int), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(o_dev, offs.data(), dim * sizeof(int), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    const int grid = grid_for(count * static_cast<uint64_t>(dim), 256);
    if (kind == 1)
      halton_randomized_kernel<1, Real><<<grid, 256, 0, s>>>(
          w_dev, n_dev, r_dev, o_dev, perms_dev, radix_sum, z_dev, dim, max_size, first_index,
          count, logtab, static_cast<Real*>(out_dev));
    else
      halton_randomized_kernel<2, Real><<<grid, 256, 0, s>>>(
          w_dev, n_dev, r_dev, o_dev, perms_dev, radix_sum, z_dev, dim, max_size, first_index,
          count, logtab, static_cast<Real*>(out_dev));
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // the host staging buffers go away
  }
  cudaFree(dev);
  if (e != cudaSuccess) return cuda_fail(e, "tqf_halton_randomized_fill");
  return TQF_OK;
}

// Uint64ToDouble on the host (the device twin is uint64_to_double).
static double host_uint64_to_double(uint32_t x0, uint32_t x1) {
  const uint64_t bits = (static_cast<uint64_t>((x0 & 0xFFFFFu) | 0x3FF00000u) << 32) | x1;
  double d;
  std::memcpy(&d, &bits, sizeof(d));
  return d - 1.0;
}

extern "C" {

int tqf_halton_permutations(int64_t seed, const int32_t* radixes, int dim, int num_coeffs,
                            int32_t* perms) {
  TQF_NVTX("tqf_halton_permutations");
  TQF_REQUIRE(radixes && perms && dim >= 1 && dim <= 1000, "bad radixes / dim");
  TQF_REQUIRE(num_coeffs >= 1 && num_coeffs <= 64, "bad num_coeffs");
  size_t radix_sum = 0;
  for (int d = 0; d < dim; ++d) {
    TQF_REQUIRE(radixes[d] >= 2, "bad radix");
    radix_sum += radixes[d];
  }
  std::vector<double> u;
  std::vector<int32_t> order;
  size_t offset = 0;
  for (int d = 0; d < dim; ++d) {
    const int p = radixes[d];
    u.resize(p);
    order.resize(p);
    for (int i = 0; i < num_coeffs; ++i) {
      // stateless_random_shuffle(range(p), seed=(seed + i, p)): float64 stateless
      // uniforms (two per Philox group) and a stable argsort
      const int64_t sd[2] = {seed + i, p};
      uint32_t key[2], ctr[4];
      tqf_philox_stateless_key_counter(sd, key, ctr);
      const PhiloxKey k{key[0], key[1]};
      const PhiloxCtr c{ctr[0], ctr[1], ctr[2], ctr[3]};
      for (int g = 0; 2 * g < p; ++g) {
        const uint4 w = philox_group(c, k, static_cast<uint64_t>(g));
        u[2 * g] = host_uint64_to_double(w.x, w.y);
        if (2 * g + 1 < p) u[2 * g + 1] = host_uint64_to_double(w.z, w.w);
      }
      for (int q = 0; q < p; ++q) order[q] = q;
      std::stable_sort(order.begin(), order.end(),
                       [&u](int32_t a, int32_t b) { return u[a] < u[b]; });
      std::memcpy(perms + static_cast<size_t>(i) * radix_sum + offset, order.data(),
                  static_cast<size_t>(p) * sizeof(int32_t));
    }
    offset += p;
  }
  return TQF_OK;
}

int tqf_halton_randomized_fill(const double* weights, const int32_t* sizes,
                               const int32_t* radixes, int dim, int max_size,
                               const int32_t* perms_dev, const double* zero_correction,
                               uint64_t first_index, uint64_t count, int kind, int dtype,
                               void* out_dev, void* stream) {
  TQF_NVTX("tqf_halton_randomized_fill");
  TQF_REQUIRE(weights && sizes && radixes && perms_dev && zero_correction, "null table");
  TQF_REQUIRE(dim >= 1 && dim <= 1000 && max_size >= 1 && max_size <= 64, "bad dim / max_size");
  TQF_REQUIRE(kind == 1 || kind == 2, "kind must be 1 (uniform) or 2 (normal)");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "dtype must be TQF_F32 or TQF_F64");
  for (int d = 0; d < dim; ++d)
    TQF_REQUIRE(sizes[d] >= 1 && sizes[d] <= max_size && radixes[d] >= 2, "bad sizes / radixes");
  if (count == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dtype == TQF_F64
             ? halton_randomized_impl<double>(weights, sizes, radixes, dim, max_size, perms_dev,
                                              zero_correction, first_index, count, kind,
                                              out_dev, s)
             : halton_randomized_impl<float>(weights, sizes, radixes, dim, max_size, perms_dev,
                                             zero_correction, first_index, count, kind, out_dev,
                                             s);
}

int tqf_halton_fill(const double* weights, const int32_t* sizes, const int32_t* radixes, int dim,
                    int max_size, uint64_t first_index, uint64_t count, int kind, int dtype,
                    void* out_dev, void* stream) {
  TQF_NVTX("tqf_halton_fill");
  TQF_REQUIRE(weights && sizes && radixes, "null table");
  TQF_REQUIRE(dim >= 1 && dim <= 1000 && max_size >= 1 && max_size <= 64, "bad dim / max_size");
  TQF_REQUIRE(kind == 1 || kind == 2, "kind must be 1 (uniform) or 2 (normal)");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "dtype must be TQF_F32 or TQF_F64");
  for (int d = 0; d < dim; ++d)
    TQF_REQUIRE(sizes[d] >= 1 && sizes[d] <= max_size && radixes[d] >= 2, "bad sizes / radixes");
  if (count == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dtype == TQF_F64
             ? halton_fill_impl<double>(weights, sizes, radixes, dim, max_size, first_index,
                                        count, kind, out_dev, s)
             : halton_fill_impl<float>(weights, sizes, radixes, dim, max_size, first_index, count,
                                       kind, out_dev, s);
}

int tqf_philox_uniform_fill(const uint32_t key[2], const uint32_t counter[4],
                            uint64_t first_element, uint64_t num_elements, int dtype,
                            void* out_dev, void* stream) {
  TQF_NVTX("tqf_philox_uniform_fill");
  TQF_REQUIRE(key && counter, "null key/counter");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "dtype must be TQF_F32 or TQF_F64");
  if (num_elements == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  const PhiloxKey k{key[0], key[1]};
  const PhiloxCtr c{counter[0], counter[1], counter[2], counter[3]};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == TQF_F64) {
    philox_uniform_f64_kernel<<<grid_for(num_elements / 2 + 1, 256), 256, 0, s>>>(
        k, c, first_element, num_elements, static_cast<double*>(out_dev));
  } else {
    philox_uniform_f32_kernel<<<grid_for(num_elements / 4 + 1, 256), 256, 0, s>>>(
        k, c, first_element, num_elements, static_cast<float*>(out_dev));
  }
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

int tqf_sobol_direction_numbers(const uint32_t* poly_a, const uint8_t* degree,
                                const uint32_t* m_init, int num_rows, int dim,
                                int32_t* out) {
  TQF_NVTX("tqf_sobol_direction_numbers");
  TQF_REQUIRE(out && dim >= 1, "bad output / dim");
  TQF_REQUIRE(dim - 1 <= num_rows, "dim exceeds the direction-number table");
  TQF_REQUIRE(dim == 1 || (poly_a && degree && m_init), "null table");
  for (int j = 0; j < 32; ++j) out[j] = 1;  // dimension 0 (sobol_impl.py:186)
  for (int k = 0; k + 1 < dim; ++k) {
    const int deg = degree[k];
    TQF_REQUIRE(deg >= 1 && deg <= 18, "degree out of range");
    // a_k = 2^s + 2a + 1 (sobol_impl.py:257); bit i of a_k for 0 <= i < deg.
    const uint64_t a_k = (1ull << deg) + 2ull * poly_a[k] + 1ull;
    uint64_t m[32];
    for (int j = 0; j < deg; ++j) m[j] = m_init[static_cast<size_t>(k) * 18 + j];
    for (int j = deg; j < 32; ++j) {
      uint64_t v = m[j - deg];
      for (int i = 0; i < deg; ++i) {
        if ((a_k >> i) & 1ull) v ^= m[j - deg + i] << (deg - i);
      }
      m[j] = v;
    }
    // int32 storage like the reference (column 31 wraps there and is unused).
    for (int j = 0; j < 32; ++j)
      out[static_cast<size_t>(k + 1) * 32 + j] = static_cast<int32_t>(static_cast<uint32_t>(m[j]));
  }
  return TQF_OK;
}

int tqf_sobol_direction_numbers_from_file(const char* path, int dim, int32_t* out) {
  TQF_REQUIRE(path && out && dim >= 1, "bad arguments");
  FILE* f = std::fopen(path, "r");
  if (!f) {
    set_error(std::string("cannot open ") + path);
    return TQF_ERR_IO;
  }
  std::vector<uint32_t> a;
  std::vector<uint8_t> s;
  std::vector<uint32_t> m;
  char line[1024];
  bool header = true;
  while (std::fgets(line, sizeof line, f) && static_cast<int>(a.size()) + 1 < dim) {
    if (header) {
      header = false;
      continue;
    }
    unsigned long d = 0, sv = 0, av = 0;
    int pos = 0;
    if (std::sscanf(line, "%lu %lu %lu%n", &d, &sv, &av, &pos) < 3) continue;
    a.push_back(static_cast<uint32_t>(av));
    s.push_back(static_cast<uint8_t>(sv));
    size_t base = m.size();
    m.resize(base + 18, 0u);
    const char* p = line + pos;
    for (int i = 0; i < 18; ++i) {
      unsigned long mv = 0;
      int adv = 0;
      if (std::sscanf(p, "%lu%n", &mv, &adv) < 1) break;
      m[base + i] = static_cast<uint32_t>(mv);
      p += adv;
    }
  }
  std::fclose(f);
  return tqf_sobol_direction_numbers(a.data(), s.data(), m.data(), static_cast<int>(a.size()),
                                     dim, out);
}

int tqf_math_eval(int fn, const double* in_dev, double* out_dev, uint64_t n, void* stream) {
  TQF_REQUIRE(fn >= 0 && fn <= 4, "fn must be in [0, 4]");
  if (n == 0) return TQF_OK;
  TQF_REQUIRE(in_dev && out_dev, "null argument");
  const double* logtab = nullptr;
  int rc = device_logtab(&logtab);
  if (rc != TQF_OK) return rc;
  math_eval_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      fn, in_dev, out_dev, n, logtab);
  TQF_CUDA_OK(cudaGetLastError());
  return TQF_OK;
}

int tqf_sobol_fill(const int32_t* direction_numbers, int dim, uint64_t num_results,
                   uint64_t skip, uint64_t first_result, uint64_t count, int kind, int dtype,
                   void* out_dev, void* stream) {
  TQF_NVTX("tqf_sobol_fill");
  TQF_REQUIRE(direction_numbers && dim >= 1, "null direction numbers / bad dim");
  TQF_REQUIRE(kind >= 0 && kind <= 2, "kind must be 0, 1 or 2");
  TQF_REQUIRE(dtype == TQF_F32 || dtype == TQF_F64, "dtype must be TQF_F32 or TQF_F64");
  TQF_REQUIRE(first_result + count <= num_results, "shard exceeds num_results");
  // skip + num_results must stay below 2^31 - 1 (sobol_impl.py:98-104).
  TQF_REQUIRE(skip + num_results < 2147483647ull, "skip + num_results too large");
  if (count == 0) return TQF_OK;
  TQF_REQUIRE(out_dev, "null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nd = sobol_num_digits(skip, num_results);
  const double* logtab = nullptr;
  int rc = device_logtab(&logtab);
  if (rc != TQF_OK) return rc;
  uint32_t* table = nullptr;
  rc = upload_sobol_table(direction_numbers, dim, &table, s);
  if (rc != TQF_OK) return rc;
  const uint64_t first_index = skip + 1 + first_result;
  const int grid = grid_for(count * dim, 256);
  if (kind == 0) {
    sobol_fill_kernel<0, int32_t><<<grid, 256, 0, s>>>(table, dim, first_index, count, nd, logtab,
                                                        static_cast<int32_t*>(out_dev));
  } else if (kind == 1 && dtype == TQF_F64) {
    sobol_fill_kernel<1, double><<<grid, 256, 0, s>>>(table, dim, first_index, count, nd, logtab,
                                                       static_cast<double*>(out_dev));
  } else if (kind == 1) {
    sobol_fill_kernel<1, float><<<grid, 256, 0, s>>>(table, dim, first_index, count, nd, logtab,
                                                      static_cast<float*>(out_dev));
  } else if (dtype == TQF_F64) {
    sobol_fill_kernel<2, double><<<grid, 256, 0, s>>>(table, dim, first_index, count, nd, logtab,
                                                       static_cast<double*>(out_dev));
  } else {
    sobol_fill_kernel<2, float><<<grid, 256, 0, s>>>(table, dim, first_index, count, nd, logtab,
                                                      static_cast<float*>(out_dev));
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // the table is released below
  {
    void* p = table;
    dev_release(&p, 1);
  }
  if (e != cudaSuccess) return cuda_fail(e, "sobol_fill_kernel");
  return TQF_OK;
}

}  // extern "C"
