#include "k_impact_block.cuh"
template __global__ void impact_block_kernel<128>(SimParams, double, int, int, LadderPool);
const void* b2m_k_impact_block128() { return (const void*)impact_block_kernel<128>; }
