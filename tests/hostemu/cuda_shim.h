// cuda_shim.h -- TEST INFRASTRUCTURE: a minimal CUDA execution model on CPU fibers, just enough to run the product's
// step / gradient kernels (dojo.jl_b200/csrc/dojo_kernels.cuh, dojo_grad.cuh, dojo_linalg.cuh and the kernel body extracted
// from dojo_b200.cu) unmodified in the CPU test-suite.
//
//   * one CTA at a time; every CUDA thread is a ucontext fiber with its own stack, scheduled round-robin on ONE OS thread;
//   * threads interact only at synchronisation points, each implemented as a counting barrier keyed by what CUDA keys it by:
//       __syncwarp(mask) / __shfl*_sync(mask) / __all_sync(mask)   -> (warp, mask)
//       bar.sync id, count (slot_sync) / __syncthreads*()          -> (barrier id)
//     a thread that arrives early yields until the barrier's generation changes; a full scheduling round without any
//     progress aborts with a deadlock report;
//   * shuffles go through a per-warp exchange buffer between two barriers; atomics are plain operations (one OS thread).
// Scheduling is deterministic, so a missing synchronisation in the kernel shows up as a reproducible wrong answer here;
// HOSTEMU_ORDER=reverse|random changes the order in which the threads of a round run (see run_cta): a race detector.
// This is not a performance model and not a CPU fallback: nothing in the product library includes it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

namespace emu {

struct Dim3 { unsigned x, y, z; };

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  int or_gen = 0;
};

struct Barrier { int arrived = 0; unsigned gen = 0; };

struct State {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1, nthreads = 0;
  unsigned block = 0, grid = 1;
  std::vector<double> smem;
  std::vector<int> sstatic;       // statically declared __shared__ ints of the kernel (s_env)
  std::unordered_map<uint64_t, Barrier> bars;
  double xd[64][32];              // per-warp shuffle exchange
  int xi[64][32];
  int or_val[2] = {0, 0};
  unsigned long progress = 0;     // bumped whenever a barrier completes or a fiber finishes
  std::function<void()> body;
};

inline State& S() { static State s; return s; }

static const size_t kStack = 1u << 20;

inline void yield() {
  State& s = S();
  swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline void barrier(uint64_t key, int count) {
  State& s = S();
  Barrier& b = s.bars[key];
  const unsigned gen = b.gen;
  if (++b.arrived >= count) { b.arrived = 0; b.gen++; s.progress++; return; }
  while (s.bars[key].gen == gen) yield();
}

inline void trampoline() {
  State& s = S();
  s.body();
  s.fibers[s.cur].done = true;
  s.progress++;
  swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

// run one CTA of `nthreads` threads over `body` (which reads threadIdx etc. through the accessors below)
inline void run_cta(unsigned block, unsigned grid, int nthreads, size_t smem_bytes, std::function<void()> body) {
  State& s = S();
  s.block = block; s.grid = grid; s.nthreads = nthreads; s.body = body;
  s.smem.assign(smem_bytes / sizeof(double) + 2, 0.0);
  s.sstatic.assign(128, 0);
  s.bars.clear();
  s.or_val[0] = s.or_val[1] = 0;
  if ((int)s.fibers.size() < nthreads) s.fibers.resize(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = s.fibers[t];
    if (!f.stack) {
      f.stack = (char*)mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (f.stack == (char*)MAP_FAILED) { perror("mmap"); abort(); }
    }
    f.done = false; f.or_gen = 0;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &s.sched;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  int alive = nthreads;
  unsigned long last_progress = s.progress;
  int idle_rounds = 0;
  // Scheduling order of a round: ascending thread index by default.  HOSTEMU_ORDER=reverse / random runs the threads of every round in
  // descending / a pseudo-random order instead: between two synchronisation points the result must not depend on which thread
  // runs first, so a test that passes in one order and fails in another has found a missing barrier (race detection for free).
  static const int order_mode = [] { const char* e = getenv("HOSTEMU_ORDER"); return !e ? 0 : (e[0] == 'r' && e[1] == 'e') ? 1 : (e[0] == 'r') ? 2 : 0; }();
  static const bool order_announced = [] { if (order_mode) fprintf(stderr, "hostemu: thread order of a round = %s\n", order_mode == 1 ? "reverse" : "random"); return true; }();
  (void)order_announced;
  static uint64_t lcg = 0x9E3779B97F4A7C15ull;
  std::vector<int> order(nthreads);
  while (alive > 0) {
    alive = 0;
    for (int t = 0; t < nthreads; ++t) order[t] = order_mode == 1 ? nthreads - 1 - t : t;
    if (order_mode == 2)
      for (int t = nthreads - 1; t > 0; --t) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; std::swap(order[t], order[(lcg >> 33) % (t + 1)]); }
    for (int k = 0; k < nthreads; ++k) {
      const int t = order[k];
      if (s.fibers[t].done) continue;
      s.cur = t;
      swapcontext(&s.sched, &s.fibers[t].ctx);
      if (!s.fibers[t].done) alive++;
    }
    if (s.progress == last_progress) {
      if (++idle_rounds > 1000) {
        fprintf(stderr, "hostemu: deadlock in block %u: %d threads wait at barriers that never complete\n", block, alive);
        for (auto& kv : s.bars) if (kv.second.arrived) fprintf(stderr, "  barrier key %016llx: %d arrived\n", (unsigned long long)kv.first, kv.second.arrived);
        abort();
      }
    } else { idle_rounds = 0; last_progress = s.progress; }
  }
  s.cur = -1;
}

inline int tid() { return S().cur; }
inline uint64_t warp_key(unsigned mask, int phase) { return (uint64_t(1) << 62) | (uint64_t(phase) << 48) | (uint64_t(tid() >> 5) << 32) | mask; }
inline uint64_t named_key(int id, int phase) { return (uint64_t(2) << 62) | (uint64_t(phase) << 48) | (uint64_t)(unsigned)id; }
inline int popc(unsigned m) { return __builtin_popcount(m); }

}  // namespace emu

// ------------------------------------------------------------------------------------------------ CUDA vocabulary
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__  // (library headers spell the attribute __attribute__((__noinline__)): keep every std include above this line)
#define __launch_bounds__(...)

struct int4 { int x, y, z, w; };

struct EmuIdx { unsigned x, y, z; };
#define threadIdx (EmuIdx{(unsigned)emu::tid(), 0u, 0u})
#define blockIdx (EmuIdx{emu::S().block, 0u, 0u})
#define blockDim (EmuIdx{(unsigned)emu::S().nthreads, 1u, 1u})
#define gridDim (EmuIdx{emu::S().grid, 1u, 1u})

inline double* hostemu_smem() { return emu::S().smem.data(); }
inline int* hostemu_sstatic(int n) { (void)n; return emu::S().sstatic.data(); }

inline void hostemu_bar_sync(int id, int count) { emu::barrier(emu::named_key(id, 0), count); }
inline void __syncthreads() { emu::barrier(emu::named_key(0, 0), emu::S().nthreads); }
inline int __syncthreads_or(int p) {
  emu::State& s = emu::S();
  emu::Fiber& f = s.fibers[s.cur];
  const int g = f.or_gen++;
  s.or_val[(g + 1) & 1] = 0;  // the slot of the next call: everybody has finished reading it (barrier B of call g - 1)
  s.or_val[g & 1] |= (p != 0);
  emu::barrier(emu::named_key(0, 1), s.nthreads);
  const int r = s.or_val[g & 1];
  emu::barrier(emu::named_key(0, 2), s.nthreads);
  return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::barrier(emu::warp_key(mask, 0), emu::popc(mask)); }

inline double __shfl_sync(unsigned mask, double v, int src, int width = 32) {
  emu::State& s = emu::S();
  const int t = s.cur, w = t >> 5, l = t & 31;
  s.xd[w][l] = v;
  emu::barrier(emu::warp_key(mask, 1), emu::popc(mask));
  const double r = s.xd[w][(l & ~(width - 1)) + (src & (width - 1))];
  emu::barrier(emu::warp_key(mask, 2), emu::popc(mask));
  return r;
}
inline double __shfl_xor_sync(unsigned mask, double v, int lanemask, int width = 32) {
  emu::State& s = emu::S();
  const int t = s.cur, w = t >> 5, l = t & 31;
  s.xd[w][l] = v;
  emu::barrier(emu::warp_key(mask, 1), emu::popc(mask));
  const int src = l ^ lanemask;
  const double r = ((src & ~(width - 1)) == (l & ~(width - 1))) ? s.xd[w][src] : v;
  emu::barrier(emu::warp_key(mask, 2), emu::popc(mask));
  return r;
}
inline int __shfl_sync(unsigned mask, int v, int src, int width = 32) {
  emu::State& s = emu::S();
  const int t = s.cur, w = t >> 5, l = t & 31;
  s.xi[w][l] = v;
  emu::barrier(emu::warp_key(mask, 3), emu::popc(mask));
  const int r = s.xi[w][(l & ~(width - 1)) + (src & (width - 1))];
  emu::barrier(emu::warp_key(mask, 4), emu::popc(mask));
  return r;
}
inline int __all_sync(unsigned mask, int p) {
  emu::State& s = emu::S();
  const int t = s.cur, w = t >> 5, l = t & 31;
  s.xi[w][l] = (p != 0);
  emu::barrier(emu::warp_key(mask, 3), emu::popc(mask));
  int r = 1;
  for (int k = 0; k < 32; ++k) if (mask & (1u << k)) r &= s.xi[w][k];
  emu::barrier(emu::warp_key(mask, 4), emu::popc(mask));
  return r;
}
inline int __any_sync(unsigned mask, int p) {
  emu::State& s = emu::S();
  const int t = s.cur, w = t >> 5, l = t & 31;
  s.xi[w][l] = (p != 0);
  emu::barrier(emu::warp_key(mask, 3), emu::popc(mask));
  int r = 0;
  for (int k = 0; k < 32; ++k) if (mask & (1u << k)) r |= s.xi[w][k];
  emu::barrier(emu::warp_key(mask, 4), emu::popc(mask));
  return r;
}

template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
inline void __threadfence() {}
inline void __threadfence_system() {}
template <class T> inline T atomicAdd_system(T* p, T v) { T o = *p; *p = o + v; return o; }
inline void __threadfence_block() {}
inline void __nanosleep(unsigned) { emu::yield(); }
inline long long clock64() { return 0; }
using std::max;
using std::min;
