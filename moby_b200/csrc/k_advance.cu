// Advance phase of the stepped path.
#include "sim_kernel_util.cuh"
using namespace b2m;

// ---- advance: positions (conservative advancement), forward dynamics, narrowphase; warp per env, small working set ----
__global__ void __launch_bounds__(256) advance_kernel(SimParams P, double dt, int round, int wpb) {
  extern __shared__ __align__(16) unsigned char smem[];
  const EnvDims D = env_dims(P);
  const size_t sd = (env_small_doubles(D) + 1) & ~(size_t)1, si = (env_small_ints(D) + 3) & ~(size_t)3;
  const int w = threadIdx.x >> 5;
  EnvMem m;
  env_carve_small(m, (double*)smem + (size_t)w * sd, (int*)((double*)smem + (size_t)wpb * sd) + (size_t)w * si, D);
  WarpGroup g(nullptr);
  unsigned long long lc[CNT_COUNT];
  for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
  const int count = (round == 0) ? P.n_envs : *q_count(P, round - 1, B2M_SLOT_CONT);
  const int* list = (round == 0) ? nullptr : q_list(P, round - 1, B2M_SLOT_CONT);
  int* head = q_head(P, round, B2M_SLOTS);
  unsigned long long envs = 0;
  for (int i = pull_warp(head); i < count; i = pull_warp(head)) {
    const int e = list ? list[i] : i;
    env_advance(g, P, e, m, dt, round, lc);
    envs++;
  }
  if (g.tid == 0) commit_counters(P, lc, envs);
}

const void* b2m_k_advance() { return (const void*)advance_kernel; }
