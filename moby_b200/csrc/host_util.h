// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include "../../include/b200moby.h"

b200moby_status b2m_fail(b200moby_status code, const char* fmt, ...);
bool b2m_have_device();

#define B2M_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return b2m_fail(B200MOBY_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
