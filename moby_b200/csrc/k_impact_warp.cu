// Impact phase, warp per env.
#include "sim_kernel_util.cuh"
using namespace b2m;

// ---- impact: islands, Delassus / LCP assembly, solve, impulses for the parked envs of one LCP class ----
// L.ctl != nullptr: the launch runs the Lemke ladder's rungs as tasks (lcp_device.cuh): a warp that has run out of envs
// keeps taking tasks until every warp of the launch has run out of envs and the task list is empty.
// feed_slot >= 0 (the hard-queue launch): once its own queue is empty the launch also takes the envs that the class launches
// running next to it hand on (the straggler queue of the round) -- as they arrive, instead of in a launch of their own after
// every class has finished.  feed_done counts the class launches that have completed (one signal_kernel per class, queued
// behind it on its stream); an entry of the fed queue is valid once it is no longer -1 (the producer bumps the count
// first).  A timeout ends the wait whatever happens; what is left over is run by the straggler launch that follows.
__global__ void signal_kernel(int* ctr) { __threadfence(); atomicAdd(ctr, 1); }
const void* b2m_k_signal() { return (const void*)signal_kernel; }

__global__ void __launch_bounds__(256) impact_warp_kernel(SimParams P, double dt, int round, int slot, int wpb, LadderPool L, int feed_slot, int* feed_done, int feed_expect) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int w = threadIdx.x >> 5;
  EnvMem m;
  env_mem_full(P, m, smem, w, wpb);
  WarpGroup g(nullptr);
  unsigned long long lc[CNT_COUNT], tot[CNT_COUNT];
  for (int k = 0; k < CNT_COUNT; k++) tot[k] = 0;
  unsigned long long envs = 0;
  const int count = q_size(P, round, slot);
  int* head = q_head(P, round, slot);
  LadderCtx C; C.pool = L; C.owner = blockIdx.x * wpb + w; C.wd = m.work; C.wi = m.iwork;
  for (int i = pull_warp(head); i < count; i = pull_warp(head)) {
    if (L.ctl) while (ladder_help_one(g, L, m.work, m.iwork)) {}     // rungs of a running ladder are on some env's critical path: they go before the next env
    for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
    EnvCtx cx; cx.limit = P.pivot_budget > 0; cx.budget = P.pivot_budget;
    if (L.ctl && !cx.limit) cx.ladder = &C;
    if (env_impact(g, P, q_at(P, round, slot, i), m, dt, round, lc, cx)) add_counters(tot, lc);
    envs++;
  }
  if (feed_slot >= 0) {
    int* fhead = q_head(P, round, feed_slot);
    volatile int* fcount = q_count(P, round, feed_slot);
    volatile int* flist = q_list(P, round, feed_slot);
    volatile int* fdone = feed_done;
    long long t_begin; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    for (;;) {
      if (L.ctl && ladder_help_one(g, L, m.work, m.iwork)) continue;
      int e = -2;                                        // -2: nothing to take now, -3: the producers are done and the queue is empty
      if (g.tid == 0) {
        const int finished = *fdone >= feed_expect;      // read BEFORE the count: a count read after it is final
        __threadfence();
        const int hd = *(volatile int*)fhead, cn = *fcount;
        if (hd < cn) {
          if (atomicCAS(fhead, hd, hd + 1) == hd) { while ((e = flist[hd]) < 0) __nanosleep(100); }
        } else if (finished) e = -3;
        else {
          long long t_now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
          if (t_now - t_begin > 4000000ll) e = -3;         // 4 ms: never spin for long on a producer that cannot run
        }
      }
      e = __shfl_sync(0xffffffffu, e, 0);
      if (e == -3) break;
      if (e < 0) { __nanosleep(1000); continue; }
      for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
      EnvCtx cx; cx.limit = false; cx.budget = 0;
      if (L.ctl) cx.ladder = &C;
      if (env_impact(g, P, e, m, dt, round, lc, cx)) add_counters(tot, lc);
      envs++;
    }
  }
  if (g.tid == 0) commit_counters(P, tot, envs);
  if (L.ctl) {                                     // no env left for this warp: serve ladder tasks until the launch has none left
    const int total_warps = gridDim.x * wpb;
    if (g.tid == 0) { __threadfence(); atomicAdd(L.ctl + 2, 1); }
    for (;;) {
      if (ladder_help_one(g, L, m.work, m.iwork)) continue;
      int done = 0;
      if (g.tid == 0) { volatile int* ctl = L.ctl; done = (ctl[2] >= total_warps && ctl[1] >= min(ctl[0], L.cap)) ? 1 : 0; }
      if (__shfl_sync(0xffffffffu, done, 0)) break;
      __nanosleep(500);
    }
  }
}

const void* b2m_k_impact_warp() { return (const void*)impact_warp_kernel; }

// ---- impact, L lanes per env (L = 8): 32 / L envs per warp, working sets in shared memory ----
// For the small LCP classes (one or two contacts: n <= 24).  A thread per env walks these envs' few thousand instructions
// alone, a warp per env leaves most lanes idle (7-15 active lanes per instruction in round 1's captures); eight lanes share
// one env's loops and four envs share a warp's instruction stream.  The sub-groups of a warp are independent (every
// collective carries the group's own lane mask, common.cuh: SubWarpGroup), so a long solve holds up its own eight lanes only.
// Same device functions, same arithmetic per element: bit-identical to the other kernels.
template <int L>
__global__ void __launch_bounds__(256) impact_subwarp_kernel(SimParams P, double dt, int round, int slot, int wpb) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int GPW = 32 / L;
  const int lane = threadIdx.x & 31, gib = (threadIdx.x >> 5) * GPW + lane / L;
  EnvMem m;
  env_mem_full(P, m, smem, gib, wpb * GPW);
  SubWarpGroup<L> g(nullptr);
  unsigned long long lc[CNT_COUNT], tot[CNT_COUNT];
  for (int k = 0; k < CNT_COUNT; k++) tot[k] = 0;
  unsigned long long envs = 0;
  const int count = q_size(P, round, slot);
  int* head = q_head(P, round, slot);
  for (;;) {
    int i = 0;
    if (g.tid == 0) i = atomicAdd(head, 1);
    i = __shfl_sync(g.mask, i, lane & ~(L - 1));
    if (i >= count) break;
    for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
    EnvCtx cx; cx.limit = P.pivot_budget > 0; cx.budget = P.pivot_budget;
    if (env_impact(g, P, q_at(P, round, slot, i), m, dt, round, lc, cx)) add_counters(tot, lc);
    envs++;
  }
  if (g.tid == 0) commit_counters(P, tot, envs);
}
const void* b2m_k_impact_subwarp8() { return (const void*)impact_subwarp_kernel<8>; }
