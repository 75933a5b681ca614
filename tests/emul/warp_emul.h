// warp_emul.h — just enough of the CUDA warp programming model to run the warp-level helpers of
// mjlab_b200/csrc/b2_kernel.cuh on the host: 32 threads in lock step, shuffles / ballots / __syncwarp built
// on one pthread barrier.  Test infrastructure (tests/test_warp_emul.py); never part of the product.
// Requirement inherited from the GPU code: every lane reaches every warp-synchronous call.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__

struct float4 { float x, y, z, w; };

namespace warp_emul {
struct Ctx {
  pthread_barrier_t bar;
  unsigned long long slot[32];
};
inline Ctx& ctx() { static Ctx c; return c; }
inline int& lane() { static thread_local int l = 0; return l; }
inline void sync() { pthread_barrier_wait(&ctx().bar); }
template <class T>
inline T exchange(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  memcpy(&ctx().slot[lane()], &v, sizeof(T));
  sync();
  T r;
  memcpy(&r, &ctx().slot[src & 31], sizeof(T));
  sync();
  return r;
}
}  // namespace warp_emul

inline void __syncwarp(unsigned = 0xffffffffu) { warp_emul::sync(); }
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
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __frcp_rn(float x) { return 1.f / x; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __powf(float a, float b) { return powf(a, b); }
inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
