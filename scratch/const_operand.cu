// const_operand.cu -- how fast is a DFMA stream whose multiplier is a compile-time indexed __constant__ (uniform register
// operand, loaded by LDCU) when every constant is used only ONCE per pass (the stage-3 contraction of sweep.cu)?
// Variants: constants from the constant bank, from shared memory (broadcast loads), and from registers (upper bound).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__constant__ double c_X[108];

template <int MODE>   // 0 constant bank, 1 shared memory, 2 registers (12 constants reused)
__global__ void __launch_bounds__(512, 1) k_stage3(double* out, int iters, const double* gX) {
  __shared__ double Xs[108];
  for (int i = threadIdx.x; i < 108; i += blockDim.x) Xs[i] = gX[i];
  __syncthreads();
  double in[12];
  for (int i = 0; i < 12; ++i) in[i] = threadIdx.x * 1e-3 + i;
  double acc[9];
  for (int i = 0; i < 9; ++i) acc[i] = 0.0;
  double r[12];
  for (int i = 0; i < 12; ++i) r[i] = gX[i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 9; ++u) {
      double v = acc[u];
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        if (MODE == 0) v += c_X[u * 12 + k] * in[k];
        else if (MODE == 1) v += Xs[u * 12 + k] * in[k];
        else v += r[k] * in[k];
      }
      acc[u] = v;
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) in[k] += 1e-9;
  }
  double s = 0;
  for (int i = 0; i < 9; ++i) s += acc[i];
  if (s == 1.2345) out[0] = s;
}
template <typename F> static float time_ms(F launch) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) { CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
  CK(cudaGetLastError());
  return best;
}
int main() {
  double h[108]; for (int i = 0; i < 108; ++i) h[i] = 1.0 + 1e-9 * i;
  CK(cudaMemcpyToSymbol(c_X, h, sizeof(h)));
  double *gX, *out; CK(cudaMalloc(&gX, sizeof(h))); CK(cudaMalloc(&out, 8)); CK(cudaMemcpy(gX, h, sizeof(h), cudaMemcpyHostToDevice));
  const int iters = 2000, blocks = 148;
  for (int threads : {128, 256, 512}) {
    const double fl = 2.0 * 108 * iters * (double)threads * blocks;
    float m0 = time_ms([&] { k_stage3<0><<<blocks, threads>>>(out, iters, gX); });
    float m1 = time_ms([&] { k_stage3<1><<<blocks, threads>>>(out, iters, gX); });
    float m2 = time_ms([&] { k_stage3<2><<<blocks, threads>>>(out, iters, gX); });
    printf("{\"exp\": \"stage3_like\", \"threads_per_sm\": %d, \"tflops_constant_bank\": %.2f, \"tflops_shared_broadcast\": %.2f, \"tflops_registers\": %.2f}\n",
           threads, fl / m0 * 1e-9, fl / m1 * 1e-9, fl / m2 * 1e-9);
  }
  return 0;
}
