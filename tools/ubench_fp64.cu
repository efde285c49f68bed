// FP64 pipe throughput on the device: independent DFMA / DADD / DMUL chains, per-SM rate per clock.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void k(double* out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) x[i] = __fma_rn(x[i], a, b);
      if (OP == 1) x[i] = __dadd_rn(x[i], b);
      if (OP == 2) x[i] = __dmul_rn(x[i], a);
      if (OP == 3) {  // the modular-product sequence of FastF64Pol::mulmod
        const double h = __dmul_rn(x[i], a);
        const double l = __fma_rn(x[i], a, -h);
        const double c = __dadd_rn(__fma_rn(h, b, 6755399441055744.0), -6755399441055744.0);
        x[i] = __dadd_rn(__fma_rn(-c, 1099510054913.0, h), l);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int OP, int ILP>
void run(const char* name, int ctas_per_sm, int threads, int ops_per_iter) {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<OP, ILP><<<sms * ctas_per_sm, threads>>>(out, 1.0000001, 9.0949e-13, 100);
  cudaEventRecord(e0);
  k<OP, ILP><<<sms * ctas_per_sm, threads>>>(out, 1.0000001, 9.0949e-13, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * ctas_per_sm * threads * iters * ILP * ops_per_iter;
  printf("%-28s ilp %d, %d x %d thr/SM: %.2f T inst/s = %.1f lane-inst/clk/SM (at %d MHz nominal)\n", name, ILP,
         ctas_per_sm, threads, ops / ms / 1e9, ops / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
}

int main() {
  run<0, 8>("DFMA", 4, 256, 1);
  run<1, 8>("DADD", 4, 256, 1);
  run<2, 8>("DMUL", 4, 256, 1);
  run<0, 2>("DFMA", 4, 256, 1);
  run<0, 8>("DFMA", 1, 128, 1);
  run<3, 8>("mulmod (6 FP64 inst)", 4, 256, 6);
  run<3, 8>("mulmod (6 FP64 inst)", 3, 256, 6);
  run<3, 4>("mulmod (6 FP64 inst)", 2, 256, 6);
  return 0;
}
