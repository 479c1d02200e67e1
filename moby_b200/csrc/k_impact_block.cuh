// Impact phase, one block of NT threads per env (instantiated per NT in k_impact_block{64,128,256}.cu).
#pragma once
#include "sim_kernel_util.cuh"
using namespace b2m;

template <int NT>
__global__ void __launch_bounds__(NT) impact_block_kernel(SimParams P, double dt, int round, int slot) {
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
  for (int i = pull_block(head, &next); i < count; i = pull_block(head, &next)) {
    EnvCtx cx; cx.limit = false; cx.budget = 0;
    env_impact(g, P, q_at(P, round, slot, i), m, dt, round, lc, cx);
    envs++;
  }
  if (g.tid == 0) commit_counters(P, lc, envs);
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

