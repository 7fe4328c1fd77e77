// fp64_peak.cu -- microbenchmarks of the FP64 pipes and of shared memory on B200 (sm_100a): the denominators of the
// cell-integration kernels' "FP64 pipe utilisation" (BASELINE.md section 2 left them "to be measured").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/fp64_peak scratch/fp64_peak.cu && scratch/fp64_peak
// Prints one JSON line per experiment.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
  } while (0)

__constant__ double c_tab[64];

// NCH independent FMA chains per thread, register operands
template <int NCH>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// the multiplier comes from the constant bank (uniform address): K[i,j] += X[c] * T[...] with compile-time c
template <int NCH>
__global__ void dfma_const_kernel(double* out, int iters, double b) {
  double acc[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = fma(acc[i], c_tab[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NCH>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double d0[NCH], d1[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { d0[i] = threadIdx.x * 1e-3 + i; d1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) dmma884(d0[i], d1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += d0[i] + d1[i];
  if (s == 12345.678) out[0] = s;
}

// shared-memory loads: MODE 0 = 64-bit conflict free (lane-contiguous), 1 = 64-bit broadcast (4 distinct addresses per warp),
// 2 = 128-bit conflict free, 3 = 128-bit broadcast (4 distinct), 4 = 64-bit all lanes one address
template <int MODE>
__global__ void lds_kernel(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int base;
  if (MODE == 0) base = lane;
  else if (MODE == 1) base = (lane >> 3) * 9;
  else if (MODE == 2) base = 2 * lane;
  else if (MODE == 3) base = (lane >> 3) * 18;
  else base = 0;
  base += warp * 64;
  for (int it = 0; it < iters; ++it) {
    const int o = (it & 7) * 256;
    if (MODE == 2 || MODE == 3) {
      double2 a = *reinterpret_cast<const double2*>(sm + ((base + o) & 4094));
      double2 b = *reinterpret_cast<const double2*>(sm + ((base + o + 64) & 4094));
      double2 c = *reinterpret_cast<const double2*>(sm + ((base + o + 128) & 4094));
      double2 d = *reinterpret_cast<const double2*>(sm + ((base + o + 192) & 4094));
      s0 += a.x + a.y; s1 += b.x + b.y; s2 += c.x + c.y; s3 += d.x + d.y;
    } else {
      s0 += sm[(base + o) & 4095];
      s1 += sm[(base + o + 64) & 4095];
      s2 += sm[(base + o + 128) & 4095];
      s3 += sm[(base + o + 192) & 4095];
    }
  }
  double s = s0 + s1 + s2 + s3;
  if (s == 12345.678) out[0] = s;
}

// DFMA fed from shared memory: R loads (64-bit, 4 distinct addresses per warp = the sum-factorisation pattern) per F FMAs
template <int F>
__global__ void dfma_lds_kernel(double* out, int iters, double b) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = (lane >> 3) * 9 + warp * 64;
  double acc[F];
#pragma unroll
  for (int i = 0; i < F; ++i) acc[i] = i;
  for (int it = 0; it < iters; ++it) {
    const double v = sm[(base + it * 37) & 4095];
#pragma unroll
    for (int i = 0; i < F; ++i) acc[i] = fma(acc[i], v, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < F; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out;
  CK(cudaMalloc(&out, 8));
  double h[64];
  for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-12 * i;
  CK(cudaMemcpyToSymbol(c_tab, h, sizeof(h)));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d}\n", p.name, sms, p.clockRate / 1000);
  const int iters = 20000;
  for (int wps : {4, 8, 16, 32}) {   // warps per SM
    const int threads = 32 * (wps > 16 ? 16 : wps), blocks = sms * (wps > 16 ? wps / 16 : 1);
    {
      float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 8 * iters * (double)threads * blocks;
      printf("{\"exp\": \"dfma_reg\", \"chains\": 8, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", wps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dfma_const_kernel<8><<<blocks, threads>>>(out, iters, 1e-9); });
      double fl = 2.0 * 8 * iters * (double)threads * blocks;
      printf("{\"exp\": \"dfma_const_operand\", \"chains\": 8, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", wps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
      double fl = 2.0 * 256 * 8 * (iters / 4) * (double)(threads / 32) * blocks;
      printf("{\"exp\": \"dmma_m8n8k4\", \"chains\": 8, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", wps, ms, fl / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dfma_kernel<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      double fl = 2.0 * 2 * iters * (double)threads * blocks;
      printf("{\"exp\": \"dfma_reg\", \"chains\": 2, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", wps, ms, fl / ms * 1e-9);
    }
  }
  // shared memory (16 warps per SM)
  {
    const int threads = 512, blocks = sms;
    const size_t smem = 4096 * 8;
    const char* names[5] = {"lds64_conflict_free", "lds64_4_addresses", "lds128_conflict_free", "lds128_4_addresses", "lds64_1_address"};
    float ms[5];
    ms[0] = time_ms([&] { lds_kernel<0><<<blocks, threads, smem>>>(out, iters); });
    ms[1] = time_ms([&] { lds_kernel<1><<<blocks, threads, smem>>>(out, iters); });
    ms[2] = time_ms([&] { lds_kernel<2><<<blocks, threads, smem>>>(out, iters); });
    ms[3] = time_ms([&] { lds_kernel<3><<<blocks, threads, smem>>>(out, iters); });
    ms[4] = time_ms([&] { lds_kernel<4><<<blocks, threads, smem>>>(out, iters); });
    for (int m = 0; m < 5; ++m) {
      const double insts = 4.0 * iters * (threads / 32) * blocks;   // warp-level load instructions
      const double bytes = insts * 32 * ((m == 2 || m == 3) ? 16 : 8);
      printf("{\"exp\": \"%s\", \"ms\": %.4f, \"warp_loads_per_clk_per_sm\": %.3f, \"lane_TBps\": %.2f}\n", names[m], ms[m],
             insts / sms / (ms[m] * 1e-3 * p.clockRate * 1e3), bytes / ms[m] * 1e-9);
    }
    float m1 = time_ms([&] { dfma_lds_kernel<1><<<blocks, threads, smem>>>(out, iters, 1e-9); });
    float m2 = time_ms([&] { dfma_lds_kernel<2><<<blocks, threads, smem>>>(out, iters, 1e-9); });
    float m4 = time_ms([&] { dfma_lds_kernel<4><<<blocks, threads, smem>>>(out, iters, 1e-9); });
    float m8 = time_ms([&] { dfma_lds_kernel<8><<<blocks, threads, smem>>>(out, iters, 1e-9); });
    const double fl = 2.0 * iters * (double)threads * blocks;
    printf("{\"exp\": \"dfma_per_lds64\", \"fma_per_load\": [1,2,4,8], \"tflops\": [%.2f, %.2f, %.2f, %.2f]}\n", fl / m1 * 1e-9, 2 * fl / m2 * 1e-9,
           4 * fl / m4 * 1e-9, 8 * fl / m8 * 1e-9);
  }
  return 0;
}
