// microbench.cu — B200 throughput probes that decide the pooling strategy of the KDE kernel.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int THREADS = 256;
constexpr int ITERS = 4096;

// mode 0: atomicAdd smem spread (lane-distinct random bins, 4096 bins)
// mode 1: red.shared via asm
// mode 2: plain LDS + STS RMW (racy, for throughput only)
// mode 3: atomicAdd smem, same bin for pairs of lanes (2-way collisions)
template <int MODE>
__global__ void smem_kernel (const uint32_t* __restrict__ idx, uint32_t* out) {
  __shared__ uint32_t bins[4096];
  for (int i = threadIdx.x; i < 4096; i += THREADS) bins[i] = 0;
  __syncthreads();
  uint32_t b = idx[blockIdx.x * THREADS + threadIdx.x] & 4095;
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) atomicAdd(&bins[b], 1u);
    else if (MODE == 1) { uint32_t a = (uint32_t)__cvta_generic_to_shared(&bins[b]); asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(1u) : "memory"); }
    else if (MODE == 2) { volatile uint32_t* p = bins; p[b] = p[b] + 1; }
    else { atomicAdd(&bins[b & ~1u], 1u); }
    b = (b * 1664525u + 1013904223u) & 4095;
  }
  __syncthreads();
  uint32_t s = 0;
  for (int i = threadIdx.x; i < 4096; i += THREADS) s += bins[i];
  atomicAdd(out, s);
}

// FP64 pipe: dependent-free dmul+dadd streams (8 accumulators per thread)
__global__ void fp64_kernel (double* out, double k) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __dadd_rn(a[i], __dmul_rn(a[(i + 1) & 7], k));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
}

// u32 -> f64 conversion + dmul + dadd (the dense exact-order tap)
__global__ void tap_kernel (double* out, const uint32_t* __restrict__ w, double k) {
  __shared__ uint32_t cell[4096];
  __shared__ double K[512];
  for (int i = threadIdx.x; i < 4096; i += THREADS) cell[i] = w[i];
  for (int i = threadIdx.x; i < 512; i += THREADS) K[i] = k * (i + 1);
  __syncthreads();
  double s = 0;
  for (int it = 0; it < ITERS / 128; ++it) {
    const int base = (threadIdx.x + it * 37) & 2047;
#pragma unroll 4
    for (int j = 0; j < 128; ++j) s = __dadd_rn(s, __dmul_rn(K[j], (double)cell[base + j]));
  }
  if (s == 12345.678) out[0] = s;
}

// global RED (L2) spread over a 1 MB window per CTA
__global__ void red_global_kernel (uint32_t* buf, const uint32_t* __restrict__ idx) {
  uint32_t b = idx[blockIdx.x * THREADS + threadIdx.x];
  uint32_t* mine = buf + (size_t)(blockIdx.x % 1024) * 65536;
  for (int it = 0; it < ITERS / 16; ++it) {
    atomicAdd(&mine[b & 65535], 1u);
    b = b * 1664525u + 1013904223u;
  }
}

template <class F> static float time_it (F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main () {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  const int grid = sms * 8;
  printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  uint32_t* idx; CK(cudaMalloc(&idx, (size_t)grid * THREADS * 4));
  { uint32_t* h = new uint32_t[(size_t)grid * THREADS]; uint32_t x = 12345; for (size_t i = 0; i < (size_t)grid * THREADS; ++i) { x = x * 1664525u + 1013904223u; h[i] = x >> 8; } CK(cudaMemcpy(idx, h, (size_t)grid * THREADS * 4, cudaMemcpyHostToDevice)); delete[] h; }
  uint32_t* out; CK(cudaMalloc(&out, 64)); CK(cudaMemset(out, 0, 64));
  double* dout; CK(cudaMalloc(&dout, 64));
  uint32_t* big; CK(cudaMalloc(&big, (size_t)1024 * 65536 * 4)); CK(cudaMemset(big, 0, (size_t)1024 * 65536 * 4));
  const double ops = (double)grid * THREADS * ITERS;
  const char* names[4] = {"atomicAdd smem spread", "red.shared spread", "plain LDS+STS RMW", "atomicAdd smem 2-way"};
  float ms[4];
  ms[0] = time_it([&] { smem_kernel<0><<<grid, THREADS>>>(idx, out); });
  ms[1] = time_it([&] { smem_kernel<1><<<grid, THREADS>>>(idx, out); });
  ms[2] = time_it([&] { smem_kernel<2><<<grid, THREADS>>>(idx, out); });
  ms[3] = time_it([&] { smem_kernel<3><<<grid, THREADS>>>(idx, out); });
  for (int i = 0; i < 4; ++i) printf("%-24s %8.3f ms  %8.1f Gop/s  %.3f lane-ops/clk/SM (at 1.9 GHz)\n", names[i], ms[i], ops / ms[i] / 1e6, ops / (ms[i] * 1e-3) / sms / 1.9e9);
  float f = time_it([&] { fp64_kernel<<<grid, THREADS>>>(dout, 1.0000001); });
  printf("%-24s %8.3f ms  %8.1f G(dmul+dadd)/s  %.3f pairs/clk/SM\n", "fp64 dmul+dadd", f, ops * 8 / f / 1e6, ops * 8 / (f * 1e-3) / sms / 1.9e9);
  float t = time_it([&] { tap_kernel<<<grid, THREADS>>>(dout, idx, 1.0000001); });
  printf("%-24s %8.3f ms  %8.1f Gtap/s  %.3f taps/clk/SM\n", "dense tap (cvt+mul+add)", t, ops / t / 1e6, ops / (t * 1e-3) / sms / 1.9e9);
  float g = time_it([&] { red_global_kernel<<<grid, THREADS>>>(big, idx); });
  printf("%-24s %8.3f ms  %8.1f Gop/s\n", "global RED spread", g, ops / 16 / g / 1e6);
  return 0;
}
