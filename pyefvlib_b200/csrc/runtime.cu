// runtime.cu -- process-level state (device, stream, last error, launch counter) and the device-wide scan.
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace efv {

static thread_local std::string g_last_error;
static cudaStream_t g_stream = nullptr;
static std::atomic<int64_t> g_launches{0};
static int g_num_sms = 0;
static int g_device = -1;

void set_last_error(const std::string &msg) { g_last_error = msg; }
const std::string &last_error() { return g_last_error; }

static void ensure_init() {
  if (g_stream) return;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error("libefv_b200: no CUDA device available (this library has no CPU fallback)");
  if (g_device < 0) { EFV_CUDA(cudaGetDevice(&g_device)); }
  cudaDeviceProp prop;
  EFV_CUDA(cudaGetDeviceProperties(&prop, g_device));
  g_num_sms = prop.multiProcessorCount;
  EFV_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  if (const char *cc = getenv("EFV_CACHE_CONFIG")) {      // experiment: one shared-memory carveout for every kernel
    if (strcmp(cc, "shared") == 0) cudaDeviceSetCacheConfig(cudaFuncCachePreferShared);
    if (strcmp(cc, "l1") == 0) cudaDeviceSetCacheConfig(cudaFuncCachePreferL1);
  }
  cudaMemPool_t pool;
  EFV_CUDA(cudaDeviceGetDefaultMemPool(&pool, g_device));
  uint64_t keep = UINT64_MAX;
  EFV_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
}

void set_device(int device) {
  EFV_CUDA(cudaSetDevice(device));
  if (g_stream && device != g_device) {
    cudaStreamDestroy(g_stream);
    g_stream = nullptr;
  }
  g_device = device;
  ensure_init();
}

cudaStream_t stream() { ensure_init(); return g_stream; }
int num_sms() { ensure_init(); return g_num_sms; }
void count_launch(int n) { g_launches += n; }
static std::atomic<int64_t> g_field_copies[2]{{0}, {0}}, g_field_bytes[2]{{0}, {0}};
void count_field_copy(int d2h, size_t bytes) { g_field_copies[d2h ? 1 : 0] += 1; g_field_bytes[d2h ? 1 : 0] += (int64_t)bytes; }
void field_copy_counters(int64_t *out4) {
  out4[0] = g_field_copies[0].load(); out4[1] = g_field_bytes[0].load();
  out4[2] = g_field_copies[1].load(); out4[3] = g_field_bytes[1].load();
}
int64_t launch_count() { return g_launches.load(); }

// ---- device -> host copies ----------------------------------------------------------------------------------------
constexpr size_t kBounceBytes = 8u << 20;
static char *g_bounce[2] = {nullptr, nullptr};
static cudaEvent_t g_bounce_ev[2] = {nullptr, nullptr};

void download_bytes(void *host, const void *dev, size_t bytes) {
  cudaStream_t st = stream();
  bool pinned = false;
  if (bytes >= 2 * kBounceBytes) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost);
    else cudaGetLastError();
  }
  if (bytes < 2 * kBounceBytes || pinned) {
    if (bytes) EFV_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
    EFV_CUDA(cudaStreamSynchronize(st));
    return;
  }
  if (!g_bounce[0]) {
    for (int k = 0; k < 2; ++k) {
      EFV_CUDA(cudaMallocHost((void **)&g_bounce[k], kBounceBytes));
      EFV_CUDA(cudaEventCreateWithFlags(&g_bounce_ev[k], cudaEventDisableTiming));
    }
  }
  const size_t nchunks = (bytes + kBounceBytes - 1) / kBounceBytes;
  auto chunk_size = [&](size_t k) { return std::min(kBounceBytes, bytes - k * kBounceBytes); };
  for (size_t k = 0; k <= nchunks; ++k) {
    if (k < nchunks) {            // buffer k&1 was drained two rounds ago
      EFV_CUDA(cudaMemcpyAsync(g_bounce[k & 1], (const char *)dev + k * kBounceBytes, chunk_size(k), cudaMemcpyDeviceToHost, st));
      EFV_CUDA(cudaEventRecord(g_bounce_ev[k & 1], st));
    }
    if (k > 0) {                  // drain the previous chunk while the next one is in flight
      EFV_CUDA(cudaEventSynchronize(g_bounce_ev[(k - 1) & 1]));
      memcpy((char *)host + (k - 1) * kBounceBytes, g_bounce[(k - 1) & 1], chunk_size(k - 1));
    }
  }
}

// ---- exclusive scan int32 -> int64 -------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void scan_tile_sums(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ tile_sums) {
  int64_t base = (int64_t)blockIdx.x * kScanTile;
  int64_t local = 0;
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
    if (i < n) local += in[i];
  }
  // integer block reduction
  __shared__ int64_t wsum[kScanThreads / 32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t s = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) s += wsum[w];
    tile_sums[blockIdx.x] = s;
  }
}

__global__ void scan_tile_offsets(int64_t *tile_sums, int64_t ntiles, int64_t *total) {
  // single block; sequential over chunks of blockDim.x tiles
  __shared__ int64_t buf[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < ntiles; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    int64_t v = (i < ntiles) ? tile_sums[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    // Hillis-Steele inclusive scan
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
      int64_t t = (threadIdx.x >= (unsigned)o) ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    int64_t incl = buf[threadIdx.x];
    if (i < ntiles) tile_sums[i] = carry + incl - v;   // exclusive
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void scan_apply(const int32_t *__restrict__ in, int64_t n, const int64_t *__restrict__ tile_offsets,
                           int64_t *__restrict__ out, const int64_t *__restrict__ total) {
  // thread t owns kScanItems consecutive items of the tile
  __shared__ int64_t wsum[kScanThreads / 32];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int64_t local = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    v[k] = (i < n) ? in[i] : 0;
    local += v[k];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t incl = local;
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  int64_t woff = 0;
  for (int k = 0; k < w; ++k) woff += wsum[k];
  int64_t run = tile_offsets[blockIdx.x] + woff + incl - local;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

int64_t exclusive_scan_i32_to_i64(const int32_t *counts, int64_t *offsets, int64_t n) {
  int64_t ntiles = div_up(n, kScanTile);
  if (ntiles < 1) ntiles = 1;
  DBuf<int64_t> tiles(ntiles + 1);
  EFV_LAUNCH(scan_tile_sums, (unsigned)ntiles, kScanThreads, 0, counts, n, tiles.p);
  EFV_LAUNCH(scan_tile_offsets, 1, 1024, 0, tiles.p, ntiles, tiles.p + ntiles);
  EFV_LAUNCH(scan_apply, (unsigned)ntiles, kScanThreads, 0, counts, n, tiles.p, offsets, tiles.p + ntiles);
  int64_t total = 0;
  EFV_CUDA(cudaMemcpyAsync(&total, tiles.p + ntiles, sizeof(int64_t), cudaMemcpyDeviceToHost, stream()));
  EFV_CUDA(cudaStreamSynchronize(stream()));
  return total;
}

}  // namespace efv
