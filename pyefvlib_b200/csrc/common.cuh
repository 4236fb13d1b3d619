// common.cuh -- shared device/host helpers of libefv_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace efv {

// ---- error handling: internals throw, the C-ABI layer converts to return codes -----------------------------
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
void set_last_error(const std::string &msg);
#define EFV_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      throw efv::Error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" + __FILE__ +   \
                       ":" + std::to_string(__LINE__) + ")");                                           \
  } while (0)
#define EFV_REQUIRE(cond, msg)                                                                           \
  do { if (!(cond)) throw efv::Error(std::string(msg)); } while (0)

cudaStream_t stream();          // the one stream of the library
// device -> host copy that returns when the data is in `host`; large copies into pageable memory are pipelined through
// two pinned bounce buffers (a pageable cudaMemcpy D2H runs at ~3.5 GB/s on this platform, pinned at > 20 GB/s)
void download_bytes(void *host, const void *dev, size_t bytes);
void count_launch(int n = 1);   // bookkeeping for efv_kernel_launch_count
int64_t launch_count();
void count_field_copy(int d2h, size_t bytes);   // whole-field uploads / downloads through efv_vec_upload / efv_vec_download
int num_sms();

#define EFV_LAUNCH(kernel, grid, block, smem, ...)                                                       \
  do {                                                                                                   \
    kernel<<<(grid), (block), (smem), efv::stream()>>>(__VA_ARGS__);                                     \
    efv::count_launch();                                                                                 \
    EFV_CUDA(cudaGetLastError());                                                                        \
  } while (0)

// ---- RAII device buffer -------------------------------------------------------------------------------------
template <typename T> struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  DBuf() = default;
  explicit DBuf(size_t count) { alloc(count); }
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf &operator=(DBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~DBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    // stream-ordered allocation from the device pool (release threshold = infinity, see runtime.cu): re-creating a
    // mesh / system of the same size re-uses the previous blocks instead of paying cudaMalloc for multi-GB buffers
    if (count) EFV_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), stream()));
  }
  void release() { if (p) cudaFreeAsync(p, stream()); p = nullptr; n = 0; }
  void zero() { if (n) EFV_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), stream())); }
  void upload(const T *host, size_t count) {
    if (n < count) alloc(count);
    if (count) EFV_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, stream()));
  }
  void download(T *host, size_t count) const { download_bytes(host, p, count * sizeof(T)); }
  size_t bytes() const { return n * sizeof(T); }
};

inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int grid_for(int64_t work_items, int block, int max_blocks_per_sm = 16) {
  int64_t g = div_up(work_items, block);
  int64_t cap = (int64_t)num_sms() * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- device helpers --------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int L> __device__ __forceinline__ double group_sum(double v) {   // sum over aligned groups of L lanes
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// deterministic block sum (fixed tree); result valid in thread 0.  red: >= 32 doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double *red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}
__device__ __forceinline__ double block_max(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  return v;
}

// exclusive prefix sum of int32 counts into int64 offsets (n+1 entries, last = total); returns total (host)
int64_t exclusive_scan_i32_to_i64(const int32_t *counts, int64_t *offsets, int64_t n);

}  // namespace efv
