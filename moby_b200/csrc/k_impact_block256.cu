#include "k_impact_block.cuh"
template __global__ void impact_block_kernel<256>(SimParams, double, int, int, LadderPool);
const void* b2m_k_impact_block256() { return (const void*)impact_block_kernel<256>; }
template __global__ void finish_block_kernel<256>(SimParams, double, int);
const void* b2m_k_finish_block256() { return (const void*)finish_block_kernel<256>; }
