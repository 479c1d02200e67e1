// Impact phase, one block of NT threads per env (instantiated per NT in k_impact_block{64,128,256}.cu).
#pragma once
#include "sim_kernel_util.cuh"
using namespace b2m;

// L.ctl != nullptr: the launch runs the rungs of the Lemke ladder as tasks (lcp_device.cuh), as the warp-per-env launch does:
// a block between two envs, a block waiting for its own ladder and a block that has run out of envs all take rungs.  For the
// n = 320 LCPs of the box stacks the envs whose ladder fails walk 22 rungs of 1,000 pivots; one after the other on one block
// that chain was the step time.
template <int NT>
__global__ void __launch_bounds__(NT) impact_block_kernel(SimParams P, double dt, int round, int slot, LadderPool L) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red[4 * (NT / 32) + 4];
  __shared__ int next;
  EnvMem m;
  env_mem_full(P, m, smem, 0, 1);
  BlockGroup<NT> g(red);
  unsigned long long lc[CNT_COUNT];
  for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
  const int count = q_size(P, round, slot);
  int* head = q_head(P, round, slot);
  unsigned long long envs = 0;
  LadderCtx C; C.pool = L; C.owner = blockIdx.x; C.wd = m.work; C.wi = m.iwork;
  for (int i = pull_block(head, &next); i < count; i = pull_block(head, &next)) {
    if (L.ctl) while (ladder_help_one(g, L, m.work, m.iwork)) {}
    EnvCtx cx; cx.limit = false; cx.budget = 0;
    if (L.ctl) cx.ladder = &C;
    env_impact(g, P, q_at(P, round, slot, i), m, dt, round, lc, cx);
    envs++;
  }
  if (g.tid == 0) commit_counters(P, lc, envs);
  if (L.ctl) {                                     // no env left for this block: serve ladder tasks until the launch has none left
    if (g.tid == 0) { __threadfence(); atomicAdd(L.ctl + 2, 1); }
    for (;;) {
      if (ladder_help_one(g, L, m.work, m.iwork)) continue;
      int done = 0;
      if (g.tid == 0) { volatile int* ctl = L.ctl; done = (ctl[2] >= (int)gridDim.x && ctl[1] >= min(ctl[0], L.cap)) ? 1 : 0; }
      if (g.bcast(done)) break;
      __nanosleep(500);
    }
  }
}

// Finish phase for scenes whose LCPs are too large for a warp (n in the hundreds): the envs that still have time left in
// their step after the last round run the fused mini-step loop to completion, one block per env, full working set.
template <int NT>
__global__ void __launch_bounds__(NT) finish_block_kernel(SimParams P, double dt, int round) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red[4 * (NT / 32) + 4];
  __shared__ int next;
  EnvMem m;
  env_mem_full(P, m, smem, 0, 1);
  BlockGroup<NT> g(red);
  unsigned long long lc[CNT_COUNT];
  for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
  const int count = *q_count(P, round, B2M_SLOT_CONT);
  const int* list = q_list(P, round, B2M_SLOT_CONT);
  int* head = q_head(P, round, B2M_SLOT_CONT);
  unsigned long long envs = 0;
  for (int i = pull_block(head, &next); i < count; i = pull_block(head, &next)) { env_finish(g, P, list[i], m, dt, lc); envs++; }
  if (g.tid == 0) commit_counters(P, lc, envs);
}

