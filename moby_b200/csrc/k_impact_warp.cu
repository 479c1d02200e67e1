// Impact phase, warp per env.
#include "sim_kernel_util.cuh"
using namespace b2m;

// ---- impact: islands, Delassus / LCP assembly, solve, impulses for the parked envs of one LCP class ----
__global__ void __launch_bounds__(256) impact_warp_kernel(SimParams P, double dt, int round, int slot, int wpb) {
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
  for (int i = pull_warp(head); i < count; i = pull_warp(head)) {
    for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
    EnvCtx cx; cx.limit = P.pivot_budget > 0; cx.budget = P.pivot_budget;
    if (env_impact(g, P, q_at(P, round, slot, i), m, dt, round, lc, cx)) add_counters(tot, lc);
    envs++;
  }
  if (g.tid == 0) commit_counters(P, tot, envs);
}

const void* b2m_k_impact_warp() { return (const void*)impact_warp_kernel; }
