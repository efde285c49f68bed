// FP64 pipe throughput on the device: independent DFMA / DADD / DMUL chains, per-SM rate per clock, the modular
// product and butterfly sequences of FastF64Pol, and how much integer work issues beside them.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

#define MAGIC 6755399441055744.0
#define QD 1099510054913.0

__device__ __forceinline__ double mulmod(double a, double w, double qinv) {
  const double h = __dmul_rn(a, w);
  const double l = __fma_rn(a, w, -h);
  const double c = __dadd_rn(__fma_rn(h, qinv, MAGIC), -MAGIC);
  return __dadd_rn(__fma_rn(-c, QD, h), l);
}
// the same with every DADD / DMUL written as a DFMA (does the pipe care?)
__device__ __forceinline__ double mulmod_fma(double a, double w, double qinv, double one) {
  const double h = __fma_rn(a, w, 0.0);
  const double l = __fma_rn(a, w, -h);
  const double c = __fma_rn(__fma_rn(h, qinv, MAGIC), one, -MAGIC);
  return __fma_rn(__fma_rn(-c, QD, h), one, l);
}

template <int OP, int ILP>
__global__ void k(double* out, double a, double b, int iters, unsigned long long ia) {
  double x[ILP];
  unsigned long long n[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = threadIdx.x + i;
    n[i] = threadIdx.x * 7 + i;
  }
  const double one = a > 0 ? 1.0 : 2.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) x[i] = __fma_rn(x[i], a, b);
      if (OP == 1) x[i] = __dadd_rn(x[i], b);
      if (OP == 2) x[i] = __dmul_rn(x[i], a);
      if (OP == 3) x[i] = mulmod(x[i], a, b);
      if (OP == 7) x[i] = mulmod_fma(x[i], a, b, one);
    }
    if (OP == 4 || OP == 5 || OP == 6 || OP == 8 || OP == 9) {  // butterflies on pairs (i, i + ILP/2)
#pragma unroll
      for (int i = 0; i < ILP / 2; ++i) {
        const double u = x[i];
        const double v = OP == 5 ? mulmod_fma(x[i + ILP / 2], a, b, one) : mulmod(x[i + ILP / 2], a, b);
        if (OP == 5) {
          x[i] = __fma_rn(u, one, v);
          x[i + ILP / 2] = __fma_rn(u, one, -v);
        } else {
          x[i] = __dadd_rn(u, v);
          x[i + ILP / 2] = __dadd_rn(u, -v);
        }
        if (OP == 6) {  // + 4 integer ALU instructions per butterfly
          n[i] = (n[i] ^ ia) + (n[i] >> 3);
          n[i + ILP / 2] = (n[i + ILP / 2] ^ ia) + (n[i + ILP / 2] >> 5);
        }
        if (OP == 8) {  // + 8 integer instructions per butterfly incl. 2 IMAD
          n[i] = (n[i] ^ ia) + (n[i] >> 3);
          n[i + ILP / 2] = (n[i + ILP / 2] ^ ia) + (n[i + ILP / 2] >> 5);
          n[i] = n[i] * 0x9E3779B1ull + n[i + ILP / 2];
        }
        if (OP == 9) {  // + 16 integer instructions
          n[i] = (n[i] ^ ia) + (n[i] >> 3);
          n[i + ILP / 2] = (n[i + ILP / 2] ^ ia) + (n[i + ILP / 2] >> 5);
          n[i] = n[i] * 0x9E3779B1ull + n[i + ILP / 2];
          n[i + ILP / 2] = (n[i + ILP / 2] ^ (n[i] >> 7)) + ia;
          n[i] = (n[i] & ~ia) | (n[i + ILP / 2] << 2);
        }
      }
    }
  }
  double s = 0;
  unsigned long long m = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    s += x[i];
    m += n[i];
  }
  if (s == 12345.678 || m == 99) out[0] = s + (double)m;
}

template <int OP, int ILP>
void run(const char* name, int ctas_per_sm, int threads, double ops_per_elem) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<OP, ILP><<<sms * ctas_per_sm, threads>>>(out, 1.0000001, 9.0949e-13, 100, 12345);
  cudaEventRecord(e0);
  k<OP, ILP><<<sms * ctas_per_sm, threads>>>(out, 1.0000001, 9.0949e-13, iters, 12345);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * ctas_per_sm * threads * iters * ILP * ops_per_elem;
  printf("%-44s ilp %2d, %d x %d thr/SM: %6.2f T FP64 inst/s = %5.1f lane-inst/clk/SM (at %d MHz nominal)\n", name, ILP,
         ctas_per_sm, threads, ops / ms / 1e9, ops / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
  cudaFree(out);
}

int main() {
  run<0, 8>("DFMA", 4, 256, 1);
  run<1, 8>("DADD", 4, 256, 1);
  run<2, 8>("DMUL", 4, 256, 1);
  run<0, 2>("DFMA", 4, 256, 1);
  run<0, 8>("DFMA", 1, 128, 1);
  run<0, 8>("DFMA", 2, 256, 1);
  run<3, 8>("mulmod (6 FP64 inst)", 4, 256, 6);
  run<7, 8>("mulmod, all DFMA", 4, 256, 6);
  run<3, 8>("mulmod (6 FP64 inst)", 3, 256, 6);
  run<3, 4>("mulmod (6 FP64 inst)", 2, 256, 6);
  // butterflies: 8 FP64 instructions per pair = 4 per element
  run<4, 16>("butterfly (8 FP64 inst / pair)", 4, 256, 4);
  run<4, 16>("butterfly (8 FP64 inst / pair)", 2, 256, 4);
  run<4, 16>("butterfly (8 FP64 inst / pair)", 1, 256, 4);
  run<5, 16>("butterfly, all DFMA", 4, 256, 4);
  run<5, 16>("butterfly, all DFMA", 2, 256, 4);
  run<6, 16>("butterfly + 8 int ALU inst / pair", 4, 256, 4);
  run<6, 16>("butterfly + 8 int ALU inst / pair", 2, 256, 4);
  run<8, 16>("butterfly + ~14 int inst / pair (IMAD)", 4, 256, 4);
  run<8, 16>("butterfly + ~14 int inst / pair (IMAD)", 2, 256, 4);
  run<9, 16>("butterfly + ~24 int inst / pair", 4, 256, 4);
  run<9, 16>("butterfly + ~24 int inst / pair", 2, 256, 4);
  return 0;
}
