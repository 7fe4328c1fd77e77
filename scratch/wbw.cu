// write-bandwidth microbenchmark: memset, aligned streaming store kernel, misaligned row stores
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill_kernel(double* p, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] = v;
}
__global__ void fill4_kernel(double4* p, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  double4 w = make_double4(v, v, v, v);
  for (; i < n; i += st) p[i] = w;
}
// rows of len doubles, warp per 8 rows, like the stream kernel
__global__ void rows_kernel(double* p, size_t nrows, int len, double v) {
  int lane = threadIdx.x & 31;
  size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t b = w; b * 8 < nrows; b += nw) {
    for (int g = 0; g < 8; ++g) {
      double* out = p + (b * 8 + g) * len + lane;
      for (int q = 0; q * 32 + lane < len; ++q) out[q * 32] = v;
    }
  }
}
__global__ void copy_kernel(const double4* a, double4* b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}
int main() {
  size_t n = (size_t)1060000000;  // doubles, ~8.5 GB
  double *p, *q;
  cudaMalloc(&p, n * 8); cudaMalloc(&q, n * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); cudaMemsetAsync(p, 0, n * 8); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("memset        %.3f ms  %.1f GB/s\n", ms, n * 8 / ms / 1e6);
    for (int grid : {148 * 4, 148 * 8, 148 * 16, 148 * 64}) {
      cudaEventRecord(e0); fill_kernel<<<grid, 256>>>(p, n, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      printf("fill   g=%5d %.3f ms  %.1f GB/s\n", grid, ms, n * 8 / ms / 1e6);
      cudaEventRecord(e0); fill4_kernel<<<grid, 256>>>((double4*)p, n / 4, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      printf("fill4  g=%5d %.3f ms  %.1f GB/s\n", grid, ms, n * 8 / ms / 1e6);
    }
    for (int len : {27, 45, 63, 64, 75, 125, 128}) {
      size_t nrows = n / len;
      cudaEventRecord(e0); rows_kernel<<<148 * 8, 256>>>(p, nrows, len, 2.0); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      printf("rows len=%3d  %.3f ms  %.1f GB/s\n", len, ms, nrows * len * 8 / ms / 1e6);
    }
    cudaEventRecord(e0); copy_kernel<<<148 * 16, 256>>>((double4*)p, (double4*)q, n / 4); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("copy          %.3f ms  %.1f GB/s (r+w)\n", ms, 2 * n * 8 / ms / 1e6);
  }
  return 0;
}
