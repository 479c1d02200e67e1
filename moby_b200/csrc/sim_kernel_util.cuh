// Device-side helpers shared by the step kernels (one kernel per translation unit so they compile in parallel).
#pragma once
#include "../../include/b200moby.h"
#include "sim_device.cuh"
#include "sim_launch.h"

namespace b2m {

__device__ inline void commit_counters(const SimParams& P, const unsigned long long* lc, unsigned long long envs = 0) {
  if (P.kstat) {
    if (envs) atomicAdd(P.kstat + 3 * P.kslot, envs);
    if (lc[CNT_PIVOT_FLOPS] + lc[CNT_ASM_FLOPS]) atomicAdd(P.kstat + 3 * P.kslot + 1, lc[CNT_PIVOT_FLOPS] + lc[CNT_ASM_FLOPS]);
    if (lc[CNT_LCP_SOLVES]) atomicAdd(P.kstat + 3 * P.kslot + 2, lc[CNT_LCP_SOLVES]);
  }
  for (int k = 0; k < CNT_COUNT; k++) {
    if (k == CNT_MAX_N) { if (lc[k]) atomicMax(P.counters + k, lc[k]); }
    else if (lc[k]) atomicAdd(P.counters + k, lc[k]);
  }
}
__device__ __forceinline__ void add_counters(unsigned long long* tot, const unsigned long long* lc) {
  for (int k = 0; k < CNT_COUNT; k++) { if (k == CNT_MAX_N) { if (lc[k] > tot[k]) tot[k] = lc[k]; } else tot[k] += lc[k]; }
}

// Full working set of thread group `slot` of this block (`slots` groups per block): shared memory, or the group's slice of
// the global scratch when the launch was planned with one (P.gscratch).
__device__ __forceinline__ void env_mem_full(const SimParams& P, EnvMem& m, unsigned char* smem, int slot, int slots) {
  const EnvDims D = env_dims(P);
  const size_t ed = (env_doubles(D) + 1) & ~(size_t)1, ei = (env_ints(D) + 3) & ~(size_t)3;
  if (P.gscratch) {
    double* base = P.gscratch + ((size_t)blockIdx.x * slots + slot) * P.gstride;
    env_carve(m, base, (int*)(base + ed), D);
  } else {
    env_carve(m, (double*)smem + (size_t)slot * ed, (int*)((double*)smem + (size_t)slots * ed) + (size_t)slot * ei, D);
  }
}

// next queue index for a warp (lane 0 pulls, broadcast) / a block (thread 0 pulls, broadcast through shared memory)
__device__ __forceinline__ int pull_warp(int* head) {
  int i = 0;
  if ((threadIdx.x & 31) == 0) i = atomicAdd(head, 1);
  return __shfl_sync(0xffffffffu, i, 0);
}
__device__ __forceinline__ int pull_block(int* head, int* slot) {
  __syncthreads();
  if (threadIdx.x == 0) *slot = atomicAdd(head, 1);
  __syncthreads();
  return *slot;
}


}  // namespace b2m
