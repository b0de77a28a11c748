// mem_probe.cu -- measures the memory-side roofline denominators of the traversal kernels on the box the
// bench runs on ("vs measured L2 bandwidth for C1-C4 ... measure L2 size/bandwidth on the box").
// Not part of the render path; built in-tree as libcrt_probe.so and called by bench.py / tools/l2_probe.py.
//
// Two access patterns over a working set of `bytes`:
//   sequential : every thread streams 32-byte vectors, grid-stride (what a copy roofline measures, read only)
//   records    : every LANE reads one 64-byte record at a pseudo-random index per step, the next index depends
//                on nothing it loaded (independent loads, the throughput limit) -- the shape of a BVH node fetch
//                by 32 incoherent rays: two 256-bit loads per lane, one L1 wavefront per lane and load
// and a dependent variant of `records` (next index derived from the loaded data: the latency-bound limit).
// Working sets below the L2 capacity measure L2; far above it, HBM.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace {

__device__ __forceinline__ void ld256(const float4* p, float4& a, float4& b)
{
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

__global__ void __launch_bounds__(256) k_seq(const float4* __restrict__ buf, size_t n_vec8, int passes, float* sink)
{
  float acc = 0.0f;
  for (int p = 0; p < passes; ++p)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec8; i += (size_t)gridDim.x * blockDim.x) {
      float4 a, b;
      ld256(buf + 2 * i, a, b);
      acc += a.x + b.w;
    }
  if (acc == 123.456f) *sink = acc;
}

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// DEP = 0: index stream independent of the data (throughput); DEP = 1: next index = f(loaded word) (latency chain)
template <int DEP>
__global__ void __launch_bounds__(128, 9) k_records(const float4* __restrict__ buf, uint32_t n_rec, int steps, float* sink)
{
  uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u);
  float acc = 0.0f;
  for (int k = 0; k < steps; ++k) {
    const uint32_t r = (uint32_t)(((uint64_t)s * n_rec) >> 32);
    float4 a, b, c, d;
    ld256(buf + 4 * (size_t)r, a, b);
    ld256(buf + 4 * (size_t)r + 2, c, d);
    acc += a.x + d.w;
    if (DEP) s = mix(s + __float_as_uint(b.y) + __float_as_uint(c.z));
    else s = mix(s + (uint32_t)k);
  }
  if (acc == 123.456f) *sink = acc;
}

}  // namespace

extern "C" {

// mode 0 sequential, 1 independent 64-byte records, 2 dependent 64-byte records.  Returns GB/s (bytes actually
// requested / device time, best of `reps`), or a negative CUDA error code.
double crt_probe_bandwidth(int device, size_t bytes, int mode, int reps)
{
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2.0;
  float4* buf = nullptr;
  float* sink = nullptr;
  bytes &= ~(size_t)255;
  if (bytes < 4096 || cudaMalloc(&buf, bytes) != cudaSuccess) return -3.0;
  cudaMalloc(&sink, 4);
  // non-zero, data-independent content (the dependent chain hashes it)
  cudaMemset(buf, 0x3c, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int sm = prop.multiProcessorCount;
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {      // first repetition warms the cache
    double moved = 0.0;
    cudaEventRecord(e0);
    if (mode == 0) {
      const size_t n8 = bytes / 32;
      const int passes = (int)(((size_t)4 << 30) / bytes) + 1;     // about 4 GB per timing
      k_seq<<<sm * 8, 256>>>(buf, n8, passes, sink);
      moved = (double)n8 * 32.0 * passes;
    } else {
      const uint32_t n_rec = (uint32_t)(bytes / 64);
      const int steps = mode == 1 ? 256 : 64;
      const int grid = sm * 9, block = 128;
      if (mode == 1) k_records<0><<<grid, block>>>(buf, n_rec, steps, sink);
      else k_records<1><<<grid, block>>>(buf, n_rec, steps, sink);
      moved = (double)grid * block * steps * 64.0;
    }
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -4.0; break; }
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms > 0.0f) best = best > moved / (ms * 1e6) ? best : moved / (ms * 1e6);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink);
  return best;
}

// L2 capacity the driver reports (bytes), SM count and the SM clock (kHz) of `device`.
int crt_probe_device(int device, size_t* l2_bytes, int* sm_count, int* sm_clock_khz)
{
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  if (sm_clock_khz) *sm_clock_khz = khz;
  return 0;
}

}  // extern "C"
