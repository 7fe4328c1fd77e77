// write-stream microbenchmark for the T1 group-template copy: which store pattern reaches the HBM write roof?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scratch/wbw2 scratch/wbw2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

struct Args {
  long long ngroups, nnz;
  const long long* gstart; const int* gcls; const long long* goff; const double* gtmpl; double* vals;
  const int* piece_group; long long npieces;
};

// A: persistent, warp per group (the current library kernel)
template <bool CS>
__global__ void __launch_bounds__(256) kA(Args a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long long stride = (long long)gridDim.x * wpb;
  for (long long b = (long long)blockIdx.x * wpb + warp; b < a.ngroups; b += stride) {
    const long long s0 = __ldg(a.gstart + b), s1 = __ldg(a.gstart + b + 1);
    const int c0 = __ldg(a.gcls + b), c1 = __ldg(a.gcls + b + 1);
    const double* src = a.gtmpl + (__ldg(a.goff + c0) - s0);
    const double* srcn = a.gtmpl + (__ldg(a.goff + c1) - s1);
    long long p = ((s0 + 3) & ~3ll) + lane;
    const long long pend = min((s1 + 3) & ~3ll, a.nnz);
    for (; p + 96 < s1; p += 128) {
      double v0 = __ldg(src + p), v1 = __ldg(src + p + 32), v2 = __ldg(src + p + 64), v3 = __ldg(src + p + 96);
      if (CS) { __stcs(a.vals + p, v0); __stcs(a.vals + p + 32, v1); __stcs(a.vals + p + 64, v2); __stcs(a.vals + p + 96, v3); }
      else { a.vals[p] = v0; a.vals[p + 32] = v1; a.vals[p + 64] = v2; a.vals[p + 96] = v3; }
    }
    for (; p < pend; p += 32) {
      double v = p < s1 ? __ldg(src + p) : __ldg(srcn + p);
      if (CS) __stcs(a.vals + p, v); else a.vals[p] = v;
    }
  }
}
// B: one warp per group, not persistent
template <bool CS>
__global__ void __launch_bounds__(256) kB(Args a) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= a.ngroups) return;
  const long long s0 = __ldg(a.gstart + b), s1 = __ldg(a.gstart + b + 1);
  const int c0 = __ldg(a.gcls + b), c1 = __ldg(a.gcls + b + 1);
  const double* src = a.gtmpl + (__ldg(a.goff + c0) - s0);
  const double* srcn = a.gtmpl + (__ldg(a.goff + c1) - s1);
  long long p = ((s0 + 3) & ~3ll) + lane;
  const long long pend = min((s1 + 3) & ~3ll, a.nnz);
  for (; p + 96 < s1; p += 128) {
    double v0 = __ldg(src + p), v1 = __ldg(src + p + 32), v2 = __ldg(src + p + 64), v3 = __ldg(src + p + 96);
    if (CS) { __stcs(a.vals + p, v0); __stcs(a.vals + p + 32, v1); __stcs(a.vals + p + 64, v2); __stcs(a.vals + p + 96, v3); }
    else { a.vals[p] = v0; a.vals[p + 32] = v1; a.vals[p + 64] = v2; a.vals[p + 96] = v3; }
  }
  for (; p < pend; p += 32) {
    double v = p < s1 ? __ldg(src + p) : __ldg(srcn + p);
    if (CS) __stcs(a.vals + p, v); else a.vals[p] = v;
  }
}

// B2: one warp per group, 16-byte stores (lane writes 2 consecutive doubles), 4 pairs in flight
template <bool CS>
__global__ void __launch_bounds__(256) kB2(Args a) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= a.ngroups) return;
  const long long s0 = __ldg(a.gstart + b), s1 = __ldg(a.gstart + b + 1);
  const int c0 = __ldg(a.gcls + b), c1 = __ldg(a.gcls + b + 1);
  const double* src = a.gtmpl + (__ldg(a.goff + c0) - s0);
  const double* srcn = a.gtmpl + (__ldg(a.goff + c1) - s1);
  long long p = ((s0 + 3) & ~3ll) + 2 * lane;
  const long long pend = min((s1 + 3) & ~3ll, a.nnz & ~1ll);
  for (; p < pend; p += 256) {
    double2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long q = p + 64 * u;
      v[u].x = q < s1 ? __ldg(src + q) : (q < pend ? __ldg(srcn + q) : 0.0);
      v[u].y = q + 1 < s1 ? __ldg(src + q + 1) : (q + 1 < pend ? __ldg(srcn + q + 1) : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long q = p + 64 * u;
      if (q < pend) { if (CS) __stcs(reinterpret_cast<double2*>(a.vals + q), v[u]); else *reinterpret_cast<double2*>(a.vals + q) = v[u]; }
    }
  }
}
// B3: like the library kernel now (8-byte stores, 4 in flight in every iteration)
template <bool CS>
__global__ void __launch_bounds__(256) kB3(Args a) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= a.ngroups) return;
  const long long s0 = __ldg(a.gstart + b), s1 = __ldg(a.gstart + b + 1);
  const int c0 = __ldg(a.gcls + b), c1 = __ldg(a.gcls + b + 1);
  const double* src = a.gtmpl + (__ldg(a.goff + c0) - s0);
  const double* srcn = a.gtmpl + (__ldg(a.goff + c1) - s1);
  long long p = ((s0 + 3) & ~3ll) + lane;
  const long long pend = min((s1 + 3) & ~3ll, a.nnz);
  for (; p < pend; p += 128) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long q = p + 32 * u;
      v[u] = q < s1 ? __ldg(src + q) : (q < pend ? __ldg(srcn + q) : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long q = p + 32 * u;
      if (q < pend) { if (CS) __stcs(a.vals + q, v[u]); else a.vals[q] = v[u]; }
    }
  }
}

// E: interleaved 256-byte pieces (the pattern of a plain fill kernel): piece -> group table
template <bool CS>
__global__ void __launch_bounds__(256) kE(Args a) {
  const int lane = threadIdx.x & 31;
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long piece = w; piece < a.npieces; piece += nw) {
    long long g = __ldg(a.piece_group + piece);
    const long long p = piece * 32 + lane;
    long long s1 = __ldg(a.gstart + g + 1);
    if (p >= s1) { ++g; }
    const long long s0 = __ldg(a.gstart + g);
    const int c = __ldg(a.gcls + g);
    if (p < a.nnz) {
      double v = __ldg(a.gtmpl + __ldg(a.goff + c) + (p - s0));
      if (CS) __stcs(a.vals + p, v); else a.vals[p] = v;
    }
  }
}
// F: CTA per 8 consecutive groups, pieces of 2 KB written by the whole CTA (thread t -> element base + it*256 + t)
template <bool CS>
__global__ void __launch_bounds__(256) kF(Args a) {
  __shared__ long long sg[10];
  __shared__ long long so[10];
  for (long long b0 = (long long)blockIdx.x * 8; b0 < a.ngroups; b0 += (long long)gridDim.x * 8) {
    __syncthreads();
    if (threadIdx.x < 10) {
      long long b = min(b0 + threadIdx.x, a.ngroups);
      sg[threadIdx.x] = __ldg(a.gstart + b);
      so[threadIdx.x] = __ldg(a.goff + __ldg(a.gcls + min(b, a.ngroups - 1)));
    }
    __syncthreads();
    const int ng = (int)min(8ll, a.ngroups - b0);
    const long long pbeg = (sg[0] + 3) & ~3ll, pend = min((sg[ng] + 3) & ~3ll, a.nnz);
    int g = 0;
    for (long long p = pbeg + threadIdx.x; p < pend; p += 256) {
      while (g < ng && p >= sg[g + 1]) ++g;
      double v = __ldg(a.gtmpl + so[g] + (p - sg[g]));
      if (CS) __stcs(a.vals + p, v); else a.vals[p] = v;
    }
  }
}

int main() {
  const long long ngroups = 2072672;
  const int ncls = 324;
  std::vector<long long> gstart(ngroups + 2), goff(ncls + 1);
  std::vector<int> gcls(ngroups + 2), clen(ncls);
  srand(1);
  for (int c = 0; c < ncls; ++c) clen[c] = c < 8 ? 8 * (c % 4 == 0 ? 125 : c % 4 == 1 ? 75 : c % 4 == 2 ? 45 : 27) : 8 * 27 + rand() % 700;
  clen[0] = 8 * 63; clen[1] = 8 * 63 + 3;
  long long tot = 0;
  for (int c = 0; c < ncls; ++c) { goff[c] = tot; tot += (clen[c] + 7) & ~3; }
  long long nnz = 0;
  for (long long b = 0; b < ngroups; ++b) {
    int c = (rand() % 100 < 85) ? rand() % 8 : rand() % ncls;
    gcls[b] = c; gstart[b] = nnz; nnz += clen[c];
  }
  gstart[ngroups] = nnz; gstart[ngroups + 1] = nnz; gcls[ngroups] = 0; gcls[ngroups + 1] = 0;
  const long long npieces = (nnz + 31) / 32;
  std::vector<int> pg(npieces);
  { long long g = 0; for (long long i = 0; i < npieces; ++i) { while (gstart[g + 1] <= i * 32) ++g; pg[i] = (int)g; } }
  printf("nnz %lld (%.2f GB), templates %.2f MB\n", nnz, nnz * 8 / 1e9, tot * 8 / 1e6);
  Args a{};
  a.ngroups = ngroups; a.nnz = nnz; a.npieces = npieces;
  long long *d_gs, *d_go; int *d_gc, *d_pg; double *d_t, *d_v;
  cudaMalloc(&d_gs, (ngroups + 2) * 8); cudaMalloc(&d_go, (ncls + 1) * 8); cudaMalloc(&d_gc, (ngroups + 2) * 4); cudaMalloc(&d_pg, npieces * 4);
  cudaMalloc(&d_t, (tot + 8) * 8); cudaMalloc(&d_v, (nnz + 64) * 8);
  cudaMemcpy(d_gs, gstart.data(), (ngroups + 2) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_go, goff.data(), (ncls + 1) * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_gc, gcls.data(), (ngroups + 2) * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_pg, pg.data(), npieces * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_t, 0, (tot + 8) * 8);
  a.gstart = d_gs; a.goff = d_go; a.gcls = d_gc; a.piece_group = d_pg; a.gtmpl = d_t; a.vals = d_v;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
#define TIME(name, launch)                                                                   \
  for (int rep = 0; rep < 3; ++rep) {                                                        \
    cudaEventRecord(e0); launch; cudaEventRecord(e1); cudaEventSynchronize(e1);              \
    cudaEventElapsedTime(&ms, e0, e1);                                                       \
    if (rep == 2) printf("%-28s %.3f ms  %.1f GB/s  (%s)\n", name, ms, nnz * 8 / ms / 1e6, cudaGetErrorString(cudaGetLastError())); \
  }
  TIME("memset", cudaMemsetAsync(d_v, 0, nnz * 8));
  for (int bps : {4, 8, 16, 32}) {
    char nm[64];
    snprintf(nm, 64, "A persistent g=148x%d", bps); TIME(nm, (kA<false><<<148 * bps, 256>>>(a)));
    snprintf(nm, 64, "A persistent cs g=148x%d", bps); TIME(nm, (kA<true><<<148 * bps, 256>>>(a)));
  }
  TIME("B warp/group", (kB<false><<<(unsigned)((ngroups + 7) / 8), 256>>>(a)));
  TIME("B warp/group cs", (kB<true><<<(unsigned)((ngroups + 7) / 8), 256>>>(a)));
  TIME("B2 warp/group 16B cs", (kB2<true><<<(unsigned)((ngroups + 7) / 8), 256>>>(a)));
  TIME("B2 warp/group 16B", (kB2<false><<<(unsigned)((ngroups + 7) / 8), 256>>>(a)));
  TIME("B3 warp/group 4-in-flight cs", (kB3<true><<<(unsigned)((ngroups + 7) / 8), 256>>>(a)));
  TIME("B3 cs, 128-thread CTAs", (kB3<true><<<(unsigned)((ngroups + 3) / 4), 128>>>(a)));
  TIME("B3 cs, 512-thread CTAs", (kB3<true><<<(unsigned)((ngroups + 15) / 16), 512>>>(a)));
  for (int bps : {8, 64}) {
    char nm[64];
    snprintf(nm, 64, "E pieces g=148x%d", bps); TIME(nm, (kE<false><<<148 * bps, 256>>>(a)));
    snprintf(nm, 64, "E pieces cs g=148x%d", bps); TIME(nm, (kE<true><<<148 * bps, 256>>>(a)));
  }
  for (int bps : {8, 64}) {
    char nm[64];
    snprintf(nm, 64, "F cta/8groups g=148x%d", bps); TIME(nm, (kF<false><<<148 * bps, 256>>>(a)));
    snprintf(nm, 64, "F cta/8groups cs g=148x%d", bps); TIME(nm, (kF<true><<<148 * bps, 256>>>(a)));
  }
  TIME("F cta/8groups full grid", (kF<false><<<(unsigned)((ngroups + 7) / 8), 256>>>(a)));
  TIME("F cta/8groups cs full grid", (kF<true><<<(unsigned)((ngroups + 7) / 8), 256>>>(a)));
  return 0;
}
