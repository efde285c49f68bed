// Microbenchmark: throughput of a dependent-free stream of modular butterflies on B200,
//   (a) integer Shoup with lazy quotient (what tb200_fast.cuh uses for the 40-bit primes)
//   (b) FP64: h = a*w; l = fma(a,w,-h); c = rint(h*qinv); r = fma(-c,q,h) + l
// 16 independent butterflies per thread per iteration (like one NTT round), values kept in registers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
  const unsigned xh = (unsigned)(x >> 32), xl = (unsigned)x, wh = (unsigned)(ws >> 32), wl = (unsigned)ws;
  const u64 h = (u64)xh * wh + __umulhi(xh, wl) + __umulhi(xl, wh);
  return x * w - h * q;
}
__global__ void k_int(u64* out, u64 q, u64 w, u64 ws, int iters) {
  u64 x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 977 + i * 131 + blockIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const u64 v = shoup_lazy(x[i + 8], w + it, ws, q);
      const u64 u = x[i];
      x[i] = u + v;
      x[i + 8] = u + 4 * q - v;
      x[i] = x[i] >= (q << 6) ? x[i] - (q << 6) : x[i];          // keep values bounded (not in the real kernel)
      x[i + 8] = x[i + 8] >= (q << 6) ? x[i + 8] - (q << 6) : x[i + 8];
    }
  }
  u64 s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ double dmodmul(double a, double w, double q, double qinv) {
  const double h = a * w;
  const double l = fma(a, w, -h);
  const double c = rint(h * qinv);
  return fma(-c, q, h) + l;   // in (-q, q) + small
}
__global__ void k_f64(double* out, double q, double qinv, double w, int iters) {
  double x[16];
  for (int i = 0; i < 16; ++i) x[i] = (double)(threadIdx.x * 977 + i * 131 + blockIdx.x);
  const double two_q = 2.0 * q;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double v = dmodmul(x[i + 8], w + it, q, qinv);
      const double u = x[i];
      double a = u + v, b = u - v;
      a = a >= two_q ? a - two_q : a;     // one lazy correction each, as a Harvey butterfly would do
      b = b < 0 ? b + two_q : b;
      x[i] = a;
      x[i + 8] = b;
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  u64* o1; double* o2;
  cudaMalloc(&o1, blocks * threads * 8); cudaMalloc(&o2, blocks * threads * 8);
  const u64 q = 1099510054913ull;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const double nb = (double)blocks * threads * iters * 8;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_int<<<blocks, threads>>>(o1, q, 123456789ull, (u64)(((unsigned __int128)123456789ull << 64) / q), iters); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("int  shoup_lazy butterflies: %.3f ms  -> %.2f T butterflies/s\n", ms, nb / ms / 1e9);
    cudaEventRecord(e0); k_f64<<<blocks, threads>>>(o2, (double)q, 1.0 / (double)q, 123456789.0, iters); cudaEventRecord(e1);
    cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("fp64 fma butterflies:        %.3f ms  -> %.2f T butterflies/s\n", ms, nb / ms / 1e9);
  }
  return 0;
}
