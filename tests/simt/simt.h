// simt.h -- TEST INFRASTRUCTURE: a tiny single-warp SIMT emulator so that the *same* device source
// (soft-grip_b200/csrc/*.cuh) can be compiled with g++ and executed on the CPU by the `-m "not gpu"` tests.
//
// It is NOT a CPU backend of the product: libsoftgrip.so contains no host path and the package never
// loads anything built from this header.  The emulator exists because the kernels are warp-synchronous
// code (sub-warp shuffles, ballots, __syncwarp-ordered shared memory) whose logic errors are otherwise only
// observable on a GPU box.  Each lane of a warp is a ucontext fiber; every warp-collective is a rendezvous of
// all 32 lanes, and the emulator aborts when lanes meet at different collectives (undefined behaviour on a
// GPU).  One warp per CTA (the kernels use blockDim.x == 32); CTAs run one after another.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace simt {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };
inline Dim3 threadIdx, blockIdx, blockDim, gridDim;

struct WarpState {
  ucontext_t sched;
  ucontext_t ctx[32];
  char* stack[32] = {nullptr};
  bool done[32];
  unsigned long long slot[2][32];
  int site[2][32];
  int gen = 0;
  int cur = 0;
  unsigned char* smem = nullptr;
  size_t smem_bytes = 0;
  void (*entry)(void*) = nullptr;
  void* arg = nullptr;
};
inline WarpState W;
constexpr size_t STACK_BYTES = 512 * 1024;

inline void fiber_main() {
  W.entry(W.arg);
  W.done[W.cur] = true;
  swapcontext(&W.ctx[W.cur], &W.sched);
}

// every lane calls this at a collective; returns the generation whose slots hold all lanes' values
inline int rendezvous(unsigned long long v, int site) {
  const int lane = W.cur, g = W.gen;
  W.slot[g & 1][lane] = v;
  W.site[g & 1][lane] = site;
  swapcontext(&W.ctx[lane], &W.sched);
  return g;
}

inline void run_warp() {
  for (int l = 0; l < 32; l++) {
    if (!W.stack[l]) W.stack[l] = (char*)malloc(STACK_BYTES);
    getcontext(&W.ctx[l]);
    W.ctx[l].uc_stack.ss_sp = W.stack[l];
    W.ctx[l].uc_stack.ss_size = STACK_BYTES;
    W.ctx[l].uc_link = &W.sched;
    makecontext(&W.ctx[l], (void (*)())fiber_main, 0);
    W.done[l] = false;
  }
  for (;;) {
    int ndone = 0;
    for (int l = 0; l < 32; l++) {
      if (W.done[l]) { ndone++; continue; }
      W.cur = l;
      threadIdx.x = (unsigned)l;
      swapcontext(&W.sched, &W.ctx[l]);
      if (W.done[l]) ndone++;
    }
    if (ndone == 32) break;
    if (ndone != 0) { fprintf(stderr, "simt: %d lanes exited while others wait at a collective\n", ndone); abort(); }
    const int g = W.gen & 1;
    for (int l = 1; l < 32; l++)
      if (W.site[g][l] != W.site[g][0]) {
        fprintf(stderr, "simt: divergent collectives: lane 0 at site %d, lane %d at site %d (block %u)\n", W.site[g][0], l, W.site[g][l], blockIdx.x);
        abort();
      }
    W.gen++;
  }
}

template <typename F, typename A>
struct Thunk { F f; A a; static void call(void* p) { Thunk* t = (Thunk*)p; t->f(t->a); } };

template <typename F, typename A>
inline void launch(F kernel, unsigned grid, unsigned block, size_t smem, const A& arg) {
  if (block != 32) { fprintf(stderr, "simt: only blockDim.x == 32 is emulated\n"); abort(); }
  if (smem > W.smem_bytes) { free(W.smem); W.smem = (unsigned char*)malloc(smem + 64); W.smem_bytes = smem; }
  Thunk<F, A> t{kernel, arg};
  W.entry = &Thunk<F, A>::call;
  W.arg = &t;
  gridDim.x = grid; blockDim.x = block;
  for (unsigned b = 0; b < grid; b++) {
    blockIdx.x = b;
    memset(W.smem, 0xff, smem);   // NaN pattern: reads of never-written shared memory show up as NaNs
    run_warp();
  }
}
inline unsigned char* smem_ptr() { return W.smem; }

template <typename T> inline unsigned long long to_bits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace simt

// ---- CUDA spellings -------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__
using simt::blockDim;
using simt::blockIdx;
using simt::gridDim;
using simt::threadIdx;

inline void __syncwarp(unsigned = 0xffffffffu) { simt::rendezvous(0, 1); }
inline void __syncthreads() { simt::rendezvous(0, 2); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o) {
  const int lane = simt::W.cur, g = simt::rendezvous(simt::to_bits(v), 3);
  return simt::from_bits<T>(simt::W.slot[g & 1][(lane ^ o) & 31]);
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) {
  const int g = simt::rendezvous(simt::to_bits(v), 4);
  return simt::from_bits<T>(simt::W.slot[g & 1][src & 31]);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  const int g = simt::rendezvous(pred ? 1 : 0, 5);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) if (simt::W.slot[g & 1][l]) m |= 1u << l;
  return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  const int g = simt::rendezvous(v, 6);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) r |= (unsigned)simt::W.slot[g & 1][l];
  return r;
}
inline int __reduce_max_sync(unsigned, int v) {
  const int g = simt::rendezvous(simt::to_bits(v), 7);
  int r = simt::from_bits<int>(simt::W.slot[g & 1][0]);
  for (int l = 1; l < 32; l++) { int x = simt::from_bits<int>(simt::W.slot[g & 1][l]); if (x > r) r = x; }
  return r;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
template <typename T> inline T __ldg(const T* p) { return *p; }
