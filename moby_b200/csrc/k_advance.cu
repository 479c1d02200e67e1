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

// ---- advance, one THREAD per env ----
// The advance phase is short, branchy, mostly scalar work per env (a handful of bodies and pairs): with a warp per env
// most lanes idle and every env streams the whole instruction sequence on its own (ncu: no_instruction is the top
// stall, 4.4 active threads per instruction).  Here a warp steps 32 envs in lock step instead: the same device
// functions instantiated for the one-thread group, the small working set in the thread's local memory (interleaved by
// the hardware, so equal indices of neighbouring envs coalesce), and state loads / stores that are unit-stride across
// the warp because the state is SoA over envs.  Same arithmetic per env as the warp kernel: results are bit-identical.
template <int ND, int NI>
__global__ void __launch_bounds__(128) advance_thread_kernel(SimParams P, double dt, int round) {
  double wd[ND];
  int wi[NI];
  EnvMem m;
  env_carve_small(m, wd, wi, env_dims(P));
  SerialGroup g(nullptr);
  unsigned long long lc[CNT_COUNT];
  for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
  const int count = (round == 0) ? P.n_envs : *q_count(P, round - 1, B2M_SLOT_CONT);
  const int* list = (round == 0) ? nullptr : q_list(P, round - 1, B2M_SLOT_CONT);
  unsigned long long envs = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const int e = list ? list[i] : i;
    env_advance(g, P, e, m, dt, round, lc);
    envs++;
  }
  // one set of atomics per warp
  for (int k = 0; k < CNT_COUNT; k++) {
    unsigned long long v = lc[k];
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o); v = (k == CNT_MAX_N) ? (u > v ? u : v) : v + u; }
    lc[k] = v;
  }
  for (int o = 16; o > 0; o >>= 1) envs += __shfl_xor_sync(0xffffffffu, envs, o);
  if ((threadIdx.x & 31) == 0) commit_counters(P, lc, envs);
}

const void* b2m_k_advance_thread(int cls) {
  switch (cls) {
    case 0: return (const void*)advance_thread_kernel<256, 64>;
    case 1: return (const void*)advance_thread_kernel<1024, 256>;
    default: return (const void*)advance_thread_kernel<4096, 1024>;
  }
}
