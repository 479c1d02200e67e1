// Fused comparison kernel and the finish kernel (full working set, whole mini-step loop on chip).
#include "sim_kernel_util.cuh"
using namespace b2m;

// ---- fused: one env per warp-sized block, the whole n_steps loop on chip (comparison path, B200MOBY_FUSED=1) ----
__global__ void __launch_bounds__(32) step_warp_kernel(SimParams P, double dt, int n_steps, size_t env_d) {
  extern __shared__ __align__(16) unsigned char smem[];
  EnvMem m;
  env_mem_full(P, m, smem, 0, 1);
  WarpGroup g(nullptr);
  unsigned long long lc[CNT_COUNT];
  for (int e = blockIdx.x; e < P.n_envs; e += gridDim.x) {
    for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
    EnvCtx cx; cx.limit = false; cx.budget = 0;
    env_run(g, P, e, m, dt, n_steps, lc, cx);
    if (g.tid == 0) commit_counters(P, lc);
    g.sync();
  }
}

// ---- finish: envs that still have time left in their step after the last round; fused loop, full working set ----
__global__ void __launch_bounds__(32) finish_kernel(SimParams P, double dt, int round) {
  extern __shared__ __align__(16) unsigned char smem[];
  EnvMem m;
  env_mem_full(P, m, smem, 0, 1);
  WarpGroup g(nullptr);
  unsigned long long lc[CNT_COUNT];
  for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
  const int count = *q_count(P, round, B2M_SLOT_CONT);
  const int* list = q_list(P, round, B2M_SLOT_CONT);
  int* head = q_head(P, round, B2M_SLOT_CONT);
  unsigned long long envs = 0;
  for (int i = pull_warp(head); i < count; i = pull_warp(head)) { env_finish(g, P, list[i], m, dt, lc); envs++; }
  if (g.tid == 0) commit_counters(P, lc, envs);
}

const void* b2m_k_step_warp() { return (const void*)step_warp_kernel; }
const void* b2m_k_finish() { return (const void*)finish_kernel; }
