// Measures the FP64 DFMA peak (and, for reference, shared-memory bandwidth) of the current GPU: the roofline
// denominators MEASURED_PEAKS.json lacks for this FP64 path.  Prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void smem_kernel(double* out, int iters) {
  extern __shared__ double s[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
  __syncthreads();
  double acc = 0;
  int idx = threadIdx.x;
  for (int i = 0; i < iters; i++) { acc += s[idx]; idx = (idx + blockDim.x) & 4095; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 16, blocks = sms * 8, threads = 256;
  double best = 0, best_s = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
    cudaEventRecord(e0); smem_kernel<<<blocks, threads, 32768>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double gbs = 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e9;
    if (rep > 0 && gbs > best_s) best_s = gbs;
  }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.3f, \"smem_read_gbs\": %.1f, \"how\": \"8 independent DFMA chains/thread, %d blocks x %d threads, best of 5 (CUDA events)\"}\n",
         prop.name, sms, best, best_s, blocks, threads);
  return 0;
}
