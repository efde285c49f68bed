// tb200_platform.h -- CUDA build glue.
//
// The product is compiled by nvcc for sm_100a only.  The same kernel sources can additionally be
// compiled by g++ with -DTB200_HOST_EMU for tests/emu (a thread-per-CUDA-thread interpreter that
// lets the indexing logic of the kernels be checked against the oracle in a container without a
// GPU).  That build is test infrastructure: the product loader never loads it and there is no
// CPU fallback in the product path.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>

#ifndef TB200_HOST_EMU
// ------------------------------------------------------------------ real CUDA
#include <cuda_runtime.h>
#define TB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define TB_LAUNCH_DYN(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define TB_KERNEL_SHARED __shared__
#define TB_SET_MAX_DYN_SMEM(kernel, bytes) \
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
static __device__ __forceinline__ double tb_rint(double a) { return rint(a); }
static __device__ __forceinline__ bool tb_block_any(bool pred) { return __syncthreads_or(pred ? 1 : 0) != 0; }
// streaming (L2-only) accesses for data that is touched once per kernel
static __device__ __forceinline__ long tb_ldcg(const long* p) { return __ldcg(p); }
static __device__ __forceinline__ longlong2 tb_ldcg2(const longlong2* p) { return __ldcg(p); }
static __device__ __forceinline__ void tb_stcg(long* p, long v) { __stcg(p, v); }
static __device__ __forceinline__ void tb_stcg2(longlong2* p, longlong2 v) { __stcg(p, v); }
#else
// ------------------------------------------------------------------ host emulation (tests only)
#include <cstdlib>
#include <cstring>
#include <functional>
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct emu_uint3 {
  unsigned x, y, z;
};
extern thread_local emu_uint3 threadIdx;
extern thread_local emu_uint3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;
void emu_syncthreads();
void emu_launch(dim3 grid, dim3 block, const std::function<void()>& body);
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ inline __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define TB_KERNEL_SHARED static
#define TB_SET_MAX_DYN_SMEM(kernel, bytes) ((void)0)
#define __syncthreads() emu_syncthreads()
#include <atomic>
static inline bool tb_block_any(bool pred) {
  static std::atomic<int> flag{0};
  if (threadIdx.x == 0) flag.store(0);
  emu_syncthreads();
  if (pred) flag.store(1);
  emu_syncthreads();
  const bool r = flag.load() != 0;
  emu_syncthreads();
  return r;
}
template <class T>
static inline T __ldg(const T* p) {
  return *p;
}
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
}
#include <cmath>
static inline double __longlong_as_double(long long v) {
  double d;
  std::memcpy(&d, &v, 8);
  return d;
}
static inline long long __double_as_longlong(double d) {
  long long v;
  std::memcpy(&v, &d, 8);
  return v;
}
static inline double __ll2double_rn(long long v) { return (double)v; }
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }
static inline long long __double2ll_rn(double d) { return std::llrint(d); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double tb_rint(double a) { return std::nearbyint(a); }
static inline unsigned __umulhi(unsigned a, unsigned b) {
  return (unsigned)(((unsigned long long)a * b) >> 32);
}
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline long long __mul64hi(long long a, long long b) {
  return (long long)(((__int128)a * (__int128)b) >> 64);
}
struct alignas(16) longlong2 {
  long long x, y;
};
static inline long tb_ldcg(const long* p) { return *p; }
static inline longlong2 tb_ldcg2(const longlong2* p) { return *p; }
static inline void tb_stcg(long* p, long v) { *p = v; }
static inline void tb_stcg2(longlong2* p, longlong2 v) { *p = v; }
#include <algorithm>
using std::min;
typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = std::malloc(n ? n : 1);
  return *p ? 0 : 2;
}
static inline cudaError_t cudaFree(void* p) {
  std::free(p);
  return 0;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memcpy(d, s, n);
  return 0;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memcpy(d, s, n);
  return 0;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
  std::memset(d, v, n);
  return 0;
}
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
typedef void* cudaEvent_t;
#define cudaStreamNonBlocking 1
#define cudaEventDisableTiming 2
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = nullptr;
  return 0;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = nullptr;
  return 0;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return 0;
}
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
#define TB_LAUNCH(kernel, grid, block, stream, ...) \
  emu_launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
#define TB_LAUNCH_DYN(kernel, grid, block, smem, stream, ...) \
  emu_launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
#endif
