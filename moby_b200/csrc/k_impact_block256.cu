#include "k_impact_block.cuh"
template __global__ void impact_block_kernel<256>(SimParams, double, int, int);
const void* b2m_k_impact_block256() { return (const void*)impact_block_kernel<256>; }
