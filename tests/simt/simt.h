// simt.h -- TEST INFRASTRUCTURE: a tiny single-warp SIMT emulator so that the *same* device source
// (soft-grip_b200/csrc/*.cuh) can be compiled with g++ and executed on the CPU by the `-m "not gpu"` tests.
//
// It is NOT a CPU backend of the product: libsoftgrip.so contains no host path and the package never
// loads anything built from this header.  The emulator exists because the kernels are warp-synchronous
// code (sub-warp shuffles, ballots, __syncwarp-ordered shared memory) whose logic errors are otherwise only
// observable on a GPU box.  Each lane of a warp is a ucontext fiber; every warp-collective is a rendezvous of
// all 32 lanes, and the emulator aborts when lanes meet at different collectives (undefined behaviour on a
// GPU).  A CTA may hold several warps (__syncthreads is a rendezvous of all of them); CTAs run one after another.
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
  ucontext_t ctx[32];
  char* stack[32] = {nullptr};
  bool done[32];
  unsigned long long slot[2][32];
  int site[2][32];
  int gen = 0;
  bool finished = false, at_cta_barrier = false;
  int named_id = -1, named_warps = 0;     // waiting at bar.sync id, nthreads (a subset of the CTA's warps)
  // tensor-memory accesses since the warp's last collective, per lane: count and rolling hash of (kind, address).  The
  // instructions are warp-collective with a uniform address; each lane only touches its own TMEM lane, so the emulator
  // executes them lane by lane and compares the logs at the next rendezvous instead of paying one per access.
  unsigned tm_n[32] = {0};
  unsigned long long tm_hash[32] = {0};
};
constexpr int MAX_WARPS = 32;
struct CtaState {
  ucontext_t sched;
  WarpState warp[MAX_WARPS];
  int nwarps = 1;
  int cur_warp = 0, cur = 0;          // running fiber
  unsigned char* smem = nullptr;
  size_t smem_bytes = 0;
  void (*entry)(void*) = nullptr;
  void* arg = nullptr;
};
inline CtaState CTA;
constexpr size_t STACK_BYTES = 512 * 1024;
enum { SITE_SYNCWARP = 1, SITE_SYNCTHREADS = 2, SITE_NAMED = 1000 };

inline void fiber_main() {
  CTA.entry(CTA.arg);
  WarpState& W = CTA.warp[CTA.cur_warp];
  W.done[CTA.cur] = true;
  swapcontext(&W.ctx[CTA.cur], &CTA.sched);
}

// every lane of a warp calls this at a collective; returns the generation whose slots hold all lanes' values
inline int rendezvous(unsigned long long v, int site) {
  WarpState& W = CTA.warp[CTA.cur_warp];
  const int lane = CTA.cur, g = W.gen;
  W.slot[g & 1][lane] = v;
  W.site[g & 1][lane] = site;
  swapcontext(&W.ctx[lane], &CTA.sched);
  return g;
}
inline WarpState& cur_warp() { return CTA.warp[CTA.cur_warp]; }
inline int cur_lane() { return CTA.cur; }

// resumes every lane of warp `wi` until its next collective (or exit)
inline void run_round(int wi) {
  WarpState& W = CTA.warp[wi];
  int ndone = 0;
  for (int l = 0; l < 32; l++) {
    if (W.done[l]) { ndone++; continue; }
    CTA.cur_warp = wi; CTA.cur = l;
    threadIdx.x = (unsigned)(wi * 32 + l);
    swapcontext(&CTA.sched, &W.ctx[l]);
    if (W.done[l]) ndone++;
  }
  for (int l = 1; l < 32; l++)
    if (W.tm_n[l] != W.tm_n[0] || W.tm_hash[l] != W.tm_hash[0]) {
      fprintf(stderr, "simt: tensor-memory accesses differ between lane 0 (%u) and lane %d (%u) of warp %d since the last collective (not warp-uniform)\n",
              W.tm_n[0], l, W.tm_n[l], wi);
      abort();
    }
  for (int l = 0; l < 32; l++) { W.tm_n[l] = 0; W.tm_hash[l] = 0; }
  if (ndone == 32) { W.finished = true; return; }
  if (ndone != 0) { fprintf(stderr, "simt: %d lanes of warp %d exited while others wait at a collective\n", ndone, wi); abort(); }
  const int g = W.gen & 1;
  for (int l = 1; l < 32; l++)
    if (W.site[g][l] != W.site[g][0]) {
      fprintf(stderr, "simt: divergent collectives: lane 0 at site %d, lane %d at site %d (block %u warp %d)\n", W.site[g][0], l, W.site[g][l], blockIdx.x, wi);
      abort();
    }
  if (W.site[g][0] == SITE_SYNCTHREADS) W.at_cta_barrier = true;
  if (W.site[g][0] >= SITE_NAMED) { W.named_id = W.site[g][0] - SITE_NAMED; W.named_warps = (int)W.slot[g][0]; }
  W.gen++;
}

inline void run_cta() {
  for (int wi = 0; wi < CTA.nwarps; wi++) {
    WarpState& W = CTA.warp[wi];
    W.gen = 0; W.finished = false; W.at_cta_barrier = false; W.named_id = -1;
    for (int l = 0; l < 32; l++) { W.tm_n[l] = 0; W.tm_hash[l] = 0; }
    for (int l = 0; l < 32; l++) {
      if (!W.stack[l]) W.stack[l] = (char*)malloc(STACK_BYTES);
      getcontext(&W.ctx[l]);
      W.ctx[l].uc_stack.ss_sp = W.stack[l];
      W.ctx[l].uc_stack.ss_size = STACK_BYTES;
      W.ctx[l].uc_link = &CTA.sched;
      makecontext(&W.ctx[l], (void (*)())fiber_main, 0);
      W.done[l] = false;
    }
  }
  for (;;) {
    int nfin = 0, nbar = 0;
    for (int wi = 0; wi < CTA.nwarps; wi++) {
      WarpState& W = CTA.warp[wi];
      if (W.finished) { nfin++; continue; }
      if (W.at_cta_barrier) { nbar++; continue; }
      if (W.named_id >= 0) continue;
      run_round(wi);
      if (W.finished) nfin++;
      else if (W.at_cta_barrier) nbar++;
    }
    if (nfin == CTA.nwarps) break;
    // named barriers: release an id once the expected number of warps waits on it
    int nnamed = 0;
    for (int id = 0; id < 16; id++) {
      int cnt = 0, want = 0;
      for (int wi = 0; wi < CTA.nwarps; wi++) if (CTA.warp[wi].named_id == id) { cnt++; want = CTA.warp[wi].named_warps; }
      if (cnt && cnt == want) { for (int wi = 0; wi < CTA.nwarps; wi++) if (CTA.warp[wi].named_id == id) CTA.warp[wi].named_id = -1; }
      else nnamed += cnt;
      if (cnt > want && want) { fprintf(stderr, "simt: %d warps at named barrier %d that expects %d\n", cnt, id, want); abort(); }
    }
    if (nbar > 0 && nbar + nfin == CTA.nwarps) {
      if (nfin) { fprintf(stderr, "simt: __syncthreads with %d exited warps\n", nfin); abort(); }
      for (int wi = 0; wi < CTA.nwarps; wi++) CTA.warp[wi].at_cta_barrier = false;
    } else {
      int runnable = 0;
      for (int wi = 0; wi < CTA.nwarps; wi++) if (!CTA.warp[wi].finished && !CTA.warp[wi].at_cta_barrier && CTA.warp[wi].named_id < 0) runnable++;
      if (!runnable) { fprintf(stderr, "simt: barrier deadlock (named-waiting %d, cta-waiting %d, finished %d)\n", nnamed, nbar, nfin); abort(); }
    }
  }
}

// tensor memory of the running CTA (see simt_tm_ld / simt_tm_st below); poisoned with NaN patterns at CTA start
inline unsigned tmem[128][512];
inline void tm_poison() { memset(tmem, 0xff, sizeof(tmem)); }

template <typename F, typename A>
struct Thunk { F f; A a; static void call(void* p) { Thunk* t = (Thunk*)p; t->f(t->a); } };

template <typename F, typename A>
inline void launch(F kernel, unsigned grid, unsigned block, size_t smem, const A& arg) {
  if (block % 32 != 0 || block / 32 > (unsigned)MAX_WARPS) { fprintf(stderr, "simt: blockDim.x must be a multiple of 32 (<= %d)\n", 32 * MAX_WARPS); abort(); }
  if (smem > CTA.smem_bytes) { free(CTA.smem); CTA.smem = (unsigned char*)malloc(smem + 64); CTA.smem_bytes = smem; }
  Thunk<F, A> t{kernel, arg};
  CTA.entry = &Thunk<F, A>::call;
  CTA.arg = &t;
  CTA.nwarps = (int)(block / 32);
  gridDim.x = grid; blockDim.x = block;
  for (unsigned b = 0; b < grid; b++) {
    blockIdx.x = b;
    if (smem) memset(CTA.smem, 0xff, smem);   // NaN pattern: reads of never-written shared memory show up as NaNs
    tm_poison();
    run_cta();
  }
}
inline unsigned char* smem_ptr() { return CTA.smem; }

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

inline void __syncwarp(unsigned = 0xffffffffu) { simt::rendezvous(0, simt::SITE_SYNCWARP); }
inline void __syncthreads() { simt::rendezvous(0, simt::SITE_SYNCTHREADS); }
inline int __syncthreads_or(int p) {
  static int acc = 0;
  if (p) acc = 1;
  __syncthreads();
  const int r = acc;
  __syncthreads();
  acc = 0;
  __syncthreads();
  return r;
}
// bar.sync id, nthreads: a barrier over nthreads/32 whole warps of the CTA
inline void simt_named_barrier(int id, int nthreads) { simt::rendezvous((unsigned long long)(nthreads / 32), simt::SITE_NAMED + id); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o) {
  const int lane = simt::cur_lane(), g = simt::rendezvous(simt::to_bits(v), 3);
  return simt::from_bits<T>(simt::cur_warp().slot[g & 1][(lane ^ o) & 31]);
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) {
  const int g = simt::rendezvous(simt::to_bits(v), 4);
  return simt::from_bits<T>(simt::cur_warp().slot[g & 1][src & 31]);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  const int g = simt::rendezvous(pred ? 1 : 0, 5);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) if (simt::cur_warp().slot[g & 1][l]) m |= 1u << l;
  return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  const int g = simt::rendezvous(v, 6);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) r |= (unsigned)simt::cur_warp().slot[g & 1][l];
  return r;
}
inline int __reduce_max_sync(unsigned, int v) {
  const int g = simt::rendezvous(simt::to_bits(v), 7);
  int r = simt::from_bits<int>(simt::cur_warp().slot[g & 1][0]);
  for (int l = 1; l < 32; l++) { int x = simt::from_bits<int>(simt::cur_warp().slot[g & 1][l]); if (x > r) r = x; }
  return r;
}
// tensor memory in the 32x32b access shape (tcgen05.ld / tcgen05.st): [128 lanes][512 columns] of 32-bit words per CTA; a thread
// of warp w owns lane 32 (w % 4) + its lane id.  The instructions are warp-collective with a warp-uniform address: the
// emulator aborts on a lane quarter that is not the warp's or a column past the end at once, and on accesses that are
// not uniform over the warp at the warp's next collective (see WarpState::tm_hash).
inline void simt_tm_check(unsigned ta, int n, int site) {
  simt::WarpState& W = simt::cur_warp();
  const int lane = simt::cur_lane();
  W.tm_n[lane]++;
  W.tm_hash[lane] = W.tm_hash[lane] * 1000003ull + (((unsigned long long)(site * 16 + n)) << 32 | ta);     // compared at the next rendezvous
  const unsigned lane0 = ta >> 16, col = ta & 0xffffu;
  if (lane0 != 32u * (unsigned)(simt::CTA.cur_warp & 3)) { fprintf(stderr, "simt: warp %d accesses tensor-memory lanes %u..\n", simt::CTA.cur_warp, lane0); abort(); }
  if (col + (unsigned)n > 512u) { fprintf(stderr, "simt: tensor-memory column %u + %d out of range\n", col, n); abort(); }
}
inline void simt_tm_ld(unsigned ta, int n, unsigned* out) {
  simt_tm_check(ta, n, 8);
  memcpy(out, &simt::tmem[(ta >> 16) + (unsigned)simt::cur_lane()][ta & 0xffffu], 4 * (size_t)n);
}
inline void simt_tm_st(unsigned ta, int n, const unsigned* in) {
  simt_tm_check(ta, n, 9);
  memcpy(&simt::tmem[(ta >> 16) + (unsigned)simt::cur_lane()][ta & 0xffffu], in, 4 * (size_t)n);
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
template <typename T> inline T __ldg(const T* p) { return *p; }
