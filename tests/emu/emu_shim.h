// emu_shim.h -- host stand-ins for the CUDA builtins the tree kernels use (TEST INFRASTRUCTURE).
// A kernel launch becomes emu::launch(grid, block, body).  The threads of a block are FIBERS (own
// stacks, cooperative switches in user space) inside one host thread, grouped in warps of 32 that
// exchange values through a barrier for the warp collectives (__shfl_*_sync, __ballot_sync,
// __match_any_sync, __reduce_*_sync, __any_sync, __syncwarp); __syncthreads is a barrier over the
// block; a barrier is "yield until everybody has arrived".  __shared__ arrays are thread-local
// statics (one block is resident per host thread), and the blocks of a grid are dealt to a few host
// threads.  Kernels without collectives can be launched `serial`.  -DGH_EMU_THREADS selects the
// older one-host-thread-per-CUDA-thread implementation (std::barrier), which the sanitizers
// understand (scripts/cpu_sanitize.sh); it runs the blocks one after another.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <semaphore>
#include <thread>
#include <vector>

#define GH_HOST_EMU 1
#define __launch_bounds__(...)
#define __grid_constant__
#include <cuda_runtime.h>
#undef __shared__
#ifdef GH_EMU_THREADS
#define __shared__ static
#else
#define __shared__ static thread_local
#endif

struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
template <class T> static inline uint64_t bits(T v) { static_assert(sizeof(T) <= 8, ""); uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
static unsigned g_block_y = 0, g_grid_y = 1;  // row of a 2-D grid being run
static thread_local int t_lane = 0;

#ifdef GH_EMU_THREADS
// ---- one host thread per CUDA thread (sanitizer builds) ------------------------------------------
struct Warp {
  std::barrier<> bar;
  uint64_t slot[2][32];
  explicit Warp(int n) : bar(n) {}
  void sync() { bar.arrive_and_wait(); }
};
static thread_local Warp *t_warp = nullptr;
static thread_local std::barrier<> *t_block = nullptr;
static thread_local unsigned t_par = 0;        // parity of the warp's exchange buffers
static int g_and_flag[3] = {1, 1, 1};
static thread_local unsigned t_and_par = 0;
static inline void block_sync() { t_block->arrive_and_wait(); }
#else
// ---- fibers ---------------------------------------------------------------------------------------
struct Bar {
  int count = 0, expected = 0;
  unsigned gen = 0;
};
struct Warp {
  Bar bar;
  uint64_t slot[2][32];
  void sync();
};
struct Fiber {
  void *sp = nullptr;
  Warp *warp = nullptr;
  unsigned tid = 0, par = 0, and_par = 0;
  bool done = false;
};
struct Block {  // the block a host thread is running
  std::vector<Fiber> f;
  std::vector<Warp> warps;
  Bar bar;
  int cur = 0, live = 0;
  long idle = 0;  // consecutive switches without progress (deadlock detector)
  void *main_sp = nullptr;
  const void *body = nullptr;
  void (*invoke)(const void *) = nullptr;
  int and_flag[3] = {1, 1, 1};
};
static thread_local Block *t_blk = nullptr;
static thread_local Warp *t_warp = nullptr;
static thread_local unsigned t_par = 0, t_and_par = 0;

extern "C" void gh_emu_switch(void **save_sp, void *load_sp);
#if defined(__x86_64__)
__asm__(".text\n.p2align 4\n.type gh_emu_switch,@function\ngh_emu_switch:\n"
        "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
        "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
        "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
        ".size gh_emu_switch,.-gh_emu_switch\n");
static constexpr int SAVED_WORDS = 6;
#elif defined(__aarch64__)
__asm__(".text\n.p2align 4\n.type gh_emu_switch,%function\ngh_emu_switch:\n"
        "  sub sp, sp, #176\n"
        "  stp x19, x20, [sp, #0]\n  stp x21, x22, [sp, #16]\n  stp x23, x24, [sp, #32]\n"
        "  stp x25, x26, [sp, #48]\n  stp x27, x28, [sp, #64]\n  stp x29, x30, [sp, #80]\n"
        "  stp d8, d9, [sp, #96]\n  stp d10, d11, [sp, #112]\n  stp d12, d13, [sp, #128]\n  stp d14, d15, [sp, #144]\n"
        "  mov x9, sp\n  str x9, [x0]\n  mov sp, x1\n"
        "  ldp x19, x20, [sp, #0]\n  ldp x21, x22, [sp, #16]\n  ldp x23, x24, [sp, #32]\n"
        "  ldp x25, x26, [sp, #48]\n  ldp x27, x28, [sp, #64]\n  ldp x29, x30, [sp, #80]\n"
        "  ldp d8, d9, [sp, #96]\n  ldp d10, d11, [sp, #112]\n  ldp d12, d13, [sp, #128]\n  ldp d14, d15, [sp, #144]\n"
        "  add sp, sp, #176\n  ret\n.size gh_emu_switch,.-gh_emu_switch\n");
#else
#error "emu_shim.h: fibers need x86-64 or aarch64 (or build with -DGH_EMU_THREADS)"
#endif

static inline void resume(Block *b, int k) {  // make fiber k the running one
  Fiber &n = b->f[(size_t)k];
  Fiber &c = b->f[(size_t)b->cur];
  c.par = t_par; c.and_par = t_and_par;
  b->cur = k;
  threadIdx.x = n.tid; t_lane = (int)(n.tid & 31); t_warp = n.warp; t_par = n.par; t_and_par = n.and_par;
  gh_emu_switch(&c.sp, n.sp);
}
static inline void yield() {  // to the next live fiber of the block (round robin)
  Block *b = t_blk;
  if (++b->idle > 64l * (long)b->f.size() + 1024) {
    std::fprintf(stderr, "emu: deadlock in block %u (a barrier some threads never reach)\n", blockIdx.x);
    std::abort();
  }
  const int n = (int)b->f.size();
  int k = b->cur;
  do { k = (k + 1 == n) ? 0 : k + 1; } while (b->f[(size_t)k].done);
  if (k != b->cur) resume(b, k);
}
static inline void bar_sync(Bar &bar) {
  if (++bar.count == bar.expected) { bar.count = 0; bar.gen++; t_blk->idle = 0; return; }
  const unsigned g = bar.gen;
  while (bar.gen == g) yield();
}
inline void Warp::sync() { bar_sync(bar); }
static inline void block_sync() { bar_sync(t_blk->bar); }

static void fiber_main() {
  Block *b = t_blk;
  b->invoke(b->body);
  Fiber &me = b->f[(size_t)b->cur];
  me.done = true;
  b->idle = 0;
  if (--b->live == 0) { gh_emu_switch(&me.sp, b->main_sp); }
  else {
    const int n = (int)b->f.size();
    int k = b->cur;
    do { k = (k + 1 == n) ? 0 : k + 1; } while (b->f[(size_t)k].done);
    resume(b, k);
  }
  std::abort();  // a finished fiber is never resumed
}

static constexpr size_t FIBER_STACK = 256 * 1024;
// fiber stacks are recycled between launches (fresh mappings cost page faults on every launch)
struct StackLease {
  char *p = nullptr;
  size_t size = 0;
  static std::mutex &mu() { static std::mutex m; return m; }
  static std::vector<std::pair<char *, size_t>> &pool() { static std::vector<std::pair<char *, size_t>> v; return v; }
  explicit StackLease(size_t need) {
    {
      std::lock_guard<std::mutex> g(mu());
      auto &v = pool();
      for (size_t k = 0; k < v.size(); k++)
        if (v[k].second >= need) { p = v[k].first; size = v[k].second; v.erase(v.begin() + (long)k); break; }
    }
    if (!p) { p = static_cast<char *>(std::malloc(need)); size = need; }
    if (!p) { std::fprintf(stderr, "emu: out of memory for fiber stacks\n"); std::abort(); }
  }
  ~StackLease() { std::lock_guard<std::mutex> g(mu()); pool().emplace_back(p, size); }
  char *get() const { return p; }
};
// run block `bx` of the grid with `nthreads` fibers on this host thread
template <class F> static void run_block(Block &b, char *stacks, unsigned bx, unsigned nthreads, unsigned grid, const F &body) {
  b.body = &body;
  b.invoke = [](const void *p) { (*static_cast<const F *>(p))(); };
  const unsigned nw = (nthreads + 31) / 32;
  b.f.assign(nthreads, Fiber());
  b.warps.assign(nw, Warp());
  for (unsigned w = 0; w < nw; w++) b.warps[w].bar.expected = (int)((w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32);
  b.bar = Bar();
  b.bar.expected = (int)nthreads;
  b.and_flag[0] = b.and_flag[1] = b.and_flag[2] = 1;
  b.live = (int)nthreads;
  b.idle = 0;
  for (unsigned t = 0; t < nthreads; t++) {
    Fiber &f = b.f[t];
    f.tid = t;
    f.warp = &b.warps[t >> 5];
    uintptr_t top = (uintptr_t)(stacks + (size_t)(t + 1) * FIBER_STACK) & ~(uintptr_t)15;
    void **sp = (void **)top;
#if defined(__x86_64__)
    *--sp = nullptr;                 // return address of fiber_main (never used)
    *--sp = (void *)&fiber_main;     // popped by gh_emu_switch's ret
    for (int k = 0; k < SAVED_WORDS; k++) *--sp = nullptr;
#else
    sp -= 22;                        // 176 bytes: x19..x30, d8..d15
    for (int k = 0; k < 22; k++) sp[k] = nullptr;
    sp[11] = (void *)&fiber_main;    // x30
#endif
    f.sp = sp;
  }
  blockIdx.x = bx; blockDim.x = nthreads; gridDim.x = grid;
  blockIdx.y = g_block_y; gridDim.y = g_grid_y;
  t_blk = &b;
  b.cur = 0;
  Fiber &f0 = b.f[0];
  threadIdx.x = 0; t_lane = 0; t_warp = f0.warp; t_par = 0; t_and_par = 0;
  gh_emu_switch(&b.main_sp, f0.sp);
  t_blk = nullptr;
}
#endif  // GH_EMU_THREADS

#ifndef GH_EMU_THREADS
// persistent host threads (creating them per launch costs more than most launches)
struct Pool {
  struct Worker {
    std::binary_semaphore go{0};
    std::thread th;
  };
  std::mutex launch_m;
  std::binary_semaphore done{0};
  std::atomic<int> pending{0};
  std::function<void()> job;
  std::vector<std::unique_ptr<Worker>> w;
  static Pool &get() { static Pool *p = new Pool(); return *p; }  // never destroyed: the threads outlive main
  size_t size() const { return w.size(); }
  Pool() {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (const char *e = std::getenv("GH_EMU_HOST_THREADS")) nt = std::atoi(e) > 0 ? (unsigned)std::atoi(e) : 1;
    if (nt > 16) nt = 16;
    if (nt <= 1) return;
    for (unsigned k = 0; k < nt; k++) {
      w.emplace_back(new Worker());
      Worker *me = w.back().get();
      me->th = std::thread([this, me] {
        for (;;) {
          me->go.acquire();
          job();
          if (pending.fetch_sub(1, std::memory_order_acq_rel) == 1) done.release();
        }
      });
      me->th.detach();
    }
  }
  // `want` workers run w() concurrently (only as many as there are blocks are woken)
  template <class W> void run(W &fn, unsigned want) {
    std::lock_guard<std::mutex> one(launch_m);
    if (want > w.size()) want = (unsigned)w.size();
    job = [&fn] { fn(); };
    pending.store((int)want, std::memory_order_release);
    for (unsigned k = 0; k < want; k++) w[k]->go.release();
    done.acquire();
  }
};
#endif

template <class F> static void launch(unsigned grid, unsigned block, F body, bool serial = false, unsigned grid_y = 1) {
  if (grid_y > 1) {  // 2-D grid: rows one after another
    for (unsigned y = 0; y < grid_y; y++) {
      g_block_y = y; g_grid_y = grid_y;
      launch(grid, block, body, serial, 1);
    }
    g_block_y = 0; g_grid_y = 1;
    return;
  }
  if (serial) {  // kernels without collectives or barriers: plain loops
    for (unsigned b = 0; b < grid; b++)
      for (unsigned t = 0; t < block; t++) {
        threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
        blockIdx.y = g_block_y; gridDim.y = g_grid_y;
        t_lane = (int)(t & 31);
        body();
      }
    return;
  }
#ifdef GH_EMU_THREADS
  // `block` host threads live for the whole launch and run the blocks one after another (a block
  // barrier between two blocks: the __shared__ statics are reused)
  const unsigned nw = (block + 31) / 32;
  std::vector<std::unique_ptr<Warp>> warps;
  for (unsigned w = 0; w < nw; w++) warps.emplace_back(new Warp((int)((w + 1) * 32 <= block ? 32 : block - w * 32)));
  std::barrier<> blockbar((std::ptrdiff_t)block);
  std::vector<std::thread> th;
  th.reserve(block);
  g_and_flag[0] = g_and_flag[1] = g_and_flag[2] = 1;
  for (unsigned t = 0; t < block; t++)
    th.emplace_back([&, t] {
      threadIdx.x = t; blockDim.x = block; gridDim.x = grid;
      blockIdx.y = g_block_y; gridDim.y = g_grid_y;
      t_lane = (int)(t & 31);
      t_warp = warps[t >> 5].get();
      t_block = &blockbar;
      for (unsigned b = 0; b < grid; b++) {
        blockIdx.x = b;
        body();
        blockbar.arrive_and_wait();
      }
    });
  for (auto &x : th) x.join();
#else
  // the blocks are dealt to a few persistent host threads; each runs one block at a time with fibers
  std::atomic<unsigned> next{0};
  auto worker = [&] {
    StackLease stacks((size_t)block * FIBER_STACK + 64);
    Block b;
    for (;;) {
      const unsigned bx = next.fetch_add(1, std::memory_order_relaxed);
      if (bx >= grid) break;
      run_block(b, stacks.get(), bx, block, grid, body);
    }
  };
  if (grid <= 2 || Pool::get().size() <= 1) { worker(); return; }
  Pool::get().run(worker, grid);
#endif
}
}  // namespace emu

static inline void __syncthreads() { emu::block_sync(); }
static inline int __syncthreads_and(int pred) {
  // three flags in rotation: the flag of barrier-and k-1 is reset at k+1, by which time everybody
  // has passed barrier k and therefore read it
#ifdef GH_EMU_THREADS
  int *flag = emu::g_and_flag;
#else
  int *flag = emu::t_blk->and_flag;
#endif
  const unsigned par = emu::t_and_par % 3u;
  emu::t_and_par++;
  if (!pred) __atomic_store_n(&flag[par], 0, __ATOMIC_RELAXED);
  __atomic_store_n(&flag[(par + 1u) % 3u], 1, __ATOMIC_RELAXED);
  emu::block_sync();
  return __atomic_load_n(&flag[par], __ATOMIC_RELAXED);
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_warp->sync(); }
// one barrier per collective: the lanes write their values into one of two slot rows used
// alternately; a row is overwritten only after the next collective's barrier, which nobody passes
// before everybody has read this one
static inline uint64_t *emu_post(uint64_t v) {
  uint64_t *row = emu::t_warp->slot[emu::t_par & 1u];
  emu::t_par++;
  row[emu::t_lane] = v;
  emu::t_warp->sync();
  return row;
}
template <class T> static inline T emu_exchange(T v, int src) {  // value of lane src (own if out of range)
  const uint64_t *row = emu_post(emu::bits(v));
  return (src >= 0 && src < 32) ? emu::from<T>(row[src]) : v;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_exchange(v, emu::t_lane + d); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_exchange(v, emu::t_lane - d >= 0 ? emu::t_lane - d : -1); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (emu::t_lane ^ m) & 31); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
  const uint64_t *row = emu_post(pred ? 1u : 0u);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(row[l] & 1u) << l;
  return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
static inline unsigned __match_any_sync(unsigned, unsigned v) {
  const uint64_t *row = emu_post(v);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(row[l] == (uint64_t)v) << l;
  return m;
}
static inline int __reduce_min_sync(unsigned, int v) {
  const uint64_t *row = emu_post(emu::bits(v));
  int r = emu::from<int>(row[0]);
  for (int l = 1; l < 32; l++) { const int o = emu::from<int>(row[l]); r = o < r ? o : r; }
  return r;
}
static inline int __reduce_max_sync(unsigned, int v) {
  const uint64_t *row = emu_post(emu::bits(v));
  int r = emu::from<int>(row[0]);
  for (int l = 1; l < 32; l++) { const int o = emu::from<int>(row[l]); r = o > r ? o : r; }
  return r;
}
static inline void __threadfence_block() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline int atomicMax(int *p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
