// warp_emul.h — just enough of the CUDA programming model to run the device code of mjlab_b200/csrc on the
// host: a CTA is a pool of host threads, each warp of 32 has its own pthread barrier on which shuffles,
// ballots and __syncwarp are built, __syncthreads is one more barrier.  Test infrastructure only
// (tests/test_warp_emul.py, tests/test_kernel_emul.py); never part of the product.
// Requirement inherited from the GPU code: every lane reaches every warp-synchronous call.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace warp_emul {
struct WarpBar {
  pthread_barrier_t bar;
  unsigned long long slot[32];
};
struct Ctx {
  WarpBar w[32];          // up to 1024 threads per CTA
  pthread_barrier_t cta;
  void* dyn_smem = nullptr;
};
inline Ctx& ctx() { static Ctx c; return c; }
struct Tl { int lane = 0, warp = 0; };
inline Tl& tl() { static thread_local Tl t; return t; }
inline int& lane() { return tl().lane; }
inline void sync() { pthread_barrier_wait(&ctx().w[tl().warp].bar); }
template <class T>
inline T exchange(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  WarpBar& wb = ctx().w[tl().warp];
  memcpy(&wb.slot[tl().lane], &v, sizeof(T));
  sync();
  T r;
  memcpy(&r, &wb.slot[src & 31], sizeof(T));
  sync();
  return r;
}
}  // namespace warp_emul

inline uint3& emul_threadIdx() { static thread_local uint3 v = {0, 0, 0}; return v; }
inline uint3& emul_blockIdx() { static thread_local uint3 v = {0, 0, 0}; return v; }
inline dim3& emul_blockDim() { static dim3 v; return v; }
inline dim3& emul_gridDim() { static dim3 v; return v; }
#define threadIdx (emul_threadIdx())
#define blockIdx (emul_blockIdx())
#define blockDim (emul_blockDim())
#define gridDim (emul_gridDim())

inline void __syncwarp(unsigned = 0xffffffffu) { warp_emul::sync(); }
inline void __syncthreads() { pthread_barrier_wait(&warp_emul::ctx().cta); }
inline int __syncthreads_or(int pred) {
  static int vote = 0;
  if (pred) __atomic_store_n(&vote, 1, __ATOMIC_RELAXED);
  __syncthreads();
  int r = __atomic_load_n(&vote, __ATOMIC_RELAXED);
  __syncthreads();
  __atomic_store_n(&vote, 0, __ATOMIC_RELAXED);
  __syncthreads();
  return r;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return warp_emul::exchange(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return warp_emul::exchange(v, warp_emul::lane() ^ m); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) {
  int src = warp_emul::lane() - d;
  return warp_emul::exchange(v, src < 0 ? warp_emul::lane() : src);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned bits = 0;
  int p = pred ? 1 : 0;
  for (int l = 0; l < 32; l++) bits |= (unsigned)warp_emul::exchange(p, l) << l;
  return bits;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __frcp_rn(float x) { return 1.f / x; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __powf(float a, float b) { return powf(a, b); }
inline float __expf(float a) { return expf(a); }
inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float atomicAdd(float* p, float v) {
  unsigned old, neu;
  float f;
  do {
    old = __atomic_load_n((unsigned*)p, __ATOMIC_RELAXED);
    memcpy(&f, &old, 4);
    f += v;
    memcpy(&neu, &f, 4);
  } while (!__atomic_compare_exchange_n((unsigned*)p, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  memcpy(&f, &old, 4);
  return f;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
