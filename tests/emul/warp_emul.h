// warp_emul.h — just enough of the CUDA programming model to run the device code of mjlab_b200/csrc on the
// host: the threads of a CTA are fibers on one host thread, each warp of 32 has its own barrier on which shuffles,
// ballots and __syncwarp are built, __syncthreads is one more barrier.  Test infrastructure only
// (tests/test_warp_emul.py, tests/test_kernel_emul.py); never part of the product.
// Requirement inherited from the GPU code: every lane reaches every warp-synchronous call.
#pragma once
#include <math.h>
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

// ---- fibers: every CUDA thread of a CTA is a user-level context on ONE host thread; a warp- or CTA-level
// synchronisation point is a cooperative yield until all participants have arrived.  (The first version ran 32 host
// threads per warp on pthread barriers: every shuffle cost a round of kernel context switches, ~100x slower.)
extern "C" void b2_fiber_switch(void** save_sp, void* load_sp);
#if defined(__x86_64__)
asm(R"(
.text
.globl b2_fiber_switch
.type b2_fiber_switch,@function
b2_fiber_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size b2_fiber_switch,.-b2_fiber_switch
)");
#else
#error "tests/emul needs x86-64 (hand-written context switch)"
#endif

#include <functional>
#include <sys/mman.h>

namespace warp_emul {
struct Bar { int count = 0, n = 0; unsigned gen = 0; };
struct WarpBar {
  Bar bar;
  unsigned long long slot[32];
};
struct Ctx {
  WarpBar w[32];          // up to 1024 threads per CTA
  Bar cta;
  void* dyn_smem = nullptr;
};
inline Ctx& ctx() { static Ctx c; return c; }
struct Tl { int lane = 0, warp = 0; };
struct Fiber { void* sp = nullptr; Tl tl; uint3 tid = {0, 0, 0}; bool done = false; };
struct Sched {
  Fiber* f = nullptr; int n = 0, cur = 0; void* main_sp = nullptr; std::function<void()>* body = nullptr;
  Fiber host;  // state seen by code that runs outside any launch
};
inline Sched& sched() { static Sched s; return s; }
inline Fiber& self() { Sched& s = sched(); return s.f ? s.f[s.cur] : s.host; }
inline Tl& tl() { return self().tl; }
inline int& lane() { return tl().lane; }
inline void yield() { Sched& s = sched(); b2_fiber_switch(&s.f[s.cur].sp, s.main_sp); }
inline void bar_wait(Bar& b) {
  const unsigned g = b.gen;
  if (++b.count == b.n) { b.count = 0; b.gen++; }
  else while (b.gen == g) yield();
}
inline void fiber_entry() {
  Sched& s = sched();
  (*s.body)();
  s.f[s.cur].done = true;
  for (;;) yield();
}
// run `body` on nt fibers (tid 0..nt-1) to completion, round robin
inline void run_fibers(int nt, std::function<void()> body) {
  Sched& s = sched();
  const size_t stack = 512 << 10;
  char* mem = (char*)mmap(nullptr, stack * (size_t)nt, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  Fiber* f = new Fiber[nt];
  for (int t = 0; t < nt; t++) {
    f[t].tl.lane = t & 31; f[t].tl.warp = t >> 5; f[t].tid = {(unsigned)t, 0, 0};
    void** top = (void**)(mem + stack * (size_t)(t + 1));  // 16-byte aligned
    top[-1] = nullptr;                     // alignment slot: after `ret` rsp is 8 mod 16, as after a call
    top[-2] = (void*)&fiber_entry;
    for (int k = 3; k <= 8; k++) top[-k] = nullptr;  // rbp rbx r12 r13 r14 r15
    f[t].sp = (void*)(top - 8);
  }
  s.f = f; s.n = nt; s.body = &body;
  for (int left = nt; left > 0;) {
    left = 0;
    for (int t = 0; t < nt; t++) {
      if (f[t].done) continue;
      s.cur = t;
      b2_fiber_switch(&s.main_sp, f[t].sp);
      left += !f[t].done;
    }
  }
  s.f = nullptr; s.n = 0; s.body = nullptr;
  delete[] f;
  munmap(mem, stack * (size_t)nt);
}
inline void sync() { bar_wait(ctx().w[tl().warp].bar); }
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

inline uint3& emul_threadIdx() { return warp_emul::self().tid; }
inline uint3& emul_blockIdx() { static uint3 v = {0, 0, 0}; return v; }
inline dim3& emul_blockDim() { static dim3 v; return v; }
inline dim3& emul_gridDim() { static dim3 v; return v; }
#define threadIdx (emul_threadIdx())
#define blockIdx (emul_blockIdx())
#define blockDim (emul_blockDim())
#define gridDim (emul_gridDim())

inline void __syncwarp(unsigned = 0xffffffffu) { warp_emul::sync(); }
inline void __syncthreads() { warp_emul::bar_wait(warp_emul::ctx().cta); }
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
