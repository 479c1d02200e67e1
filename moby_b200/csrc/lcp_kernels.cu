// Batched LCP kernels + their C ABI (include/b200moby.h): replaces LCP::lcp_lemke / lcp_fast and the
// regularised wrappers (Moby include/Moby/LCP.h:21-27) for batches of independent dense problems.
//
// Mapping: one warp per LCP while the solver's working set fits the SM's shared memory (n <= ~160: the
// Lemke tableau is n (n+2) doubles), otherwise one 256-thread block per LCP with the working set in a
// per-block global scratch that stays L2-resident.  Grids are persistent (148 SMs x resident blocks) and
// stride over the batch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/b200moby.h"
#include "host_util.h"
#include "lcp_device.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

using namespace b2m;

namespace {

enum Mode { MODE_LEMKE = 0, MODE_FAST = 1, MODE_LEMKE_REG = 2, MODE_FAST_REG = 3 };

struct LcpArgs {
  int batch, n, mode;
  const double* M; const double* q; double* z;
  double piv_tol, zero_tol;
  int warm, min_exp, step_exp, max_exp;
  int* status; int* pivots; int* log; int log_cap;
  double* scratch_d; int* scratch_i;       // block-per-LCP only
  size_t scratch_d_stride, scratch_i_stride;
};

__host__ __device__ inline size_t work_doubles(int mode, int n) {
  const size_t a = lemke_work_doubles(n), b = fast_work_doubles(n);
  if (mode == MODE_LEMKE || mode == MODE_LEMKE_REG) return a;
  return b;
}
__host__ __device__ inline size_t work_ints(int mode, int n) {
  return (mode == MODE_LEMKE || mode == MODE_LEMKE_REG) ? lemke_work_ints(n) : fast_work_ints(n);
}
// per-problem staging: M (n*n), q (n), z (n) for the lcp_fast family (random gathers into M); Lemke reads M once.
__host__ __device__ inline size_t stage_doubles(int mode, int n) {
  return (mode == MODE_FAST || mode == MODE_FAST_REG) ? (size_t)n * n + 2 * (size_t)n : 2 * (size_t)n;
}

template <class G>
__device__ void solve_one(const G& g, const LcpArgs& a, int b, double* sm_d, int* sm_i) {
  const int n = a.n;
  const double* Mg = a.M + (size_t)b * n * n;
  const double* qg = a.q + (size_t)b * n;
  double* zg = a.z + (size_t)b * n;
  int st, piv = 0, nlog = 0;
  int* logp = a.log ? a.log + (size_t)b * a.log_cap : nullptr;
  if (a.mode == MODE_FAST || a.mode == MODE_FAST_REG) {
    double* Ms = sm_d; double* qs = Ms + (size_t)n * n; double* zs = qs + n; double* wd = zs + n;
    for (int e = g.tid; e < n * n; e += G::size) Ms[e] = Mg[e];
    for (int i = g.tid; i < n; i += G::size) { qs[i] = qg[i]; zs[i] = a.warm ? zg[i] : 0.0; }
    g.sync();
    if (a.mode == MODE_FAST) st = lcp_fast_solve(g, n, Ms, n, qs, 0.0, a.zero_tol, a.warm != 0, zs, wd, sm_i, &piv, logp, a.log_cap, &nlog);
    else st = lcp_fast_regularized(g, n, Ms, n, qs, a.zero_tol, a.warm != 0, a.min_exp, a.step_exp, a.max_exp, zs, wd, sm_i, &piv, nullptr);
    g.sync();
    const bool ok = (st == LCP_OK || st == LCP_TRIVIAL || st >= LCP_REGULARIZED);
    if (ok) for (int i = g.tid; i < n; i += G::size) zg[i] = zs[i];      // on failure z is left as given (LCP.cpp:118-126,192-195)
  } else {
    double* qs = sm_d; double* zs = qs + n; double* wd = zs + n;
    for (int i = g.tid; i < n; i += G::size) qs[i] = qg[i];
    g.sync();
    if (a.mode == MODE_LEMKE) st = lemke_solve(g, n, Mg, n, qs, 0.0, a.piv_tol, a.zero_tol, zs, wd, sm_i, &piv, logp, a.log_cap, &nlog);
    else st = lcp_lemke_regularized(g, n, Mg, n, qs, a.piv_tol, a.zero_tol, a.min_exp, a.step_exp, a.max_exp, zs, wd, sm_i, &piv, nullptr);
    g.sync();
    for (int i = g.tid; i < n; i += G::size) zg[i] = zs[i];
  }
  if (g.tid == 0) {
    a.status[b] = st;
    if (a.pivots) a.pivots[b] = piv;
    if (logp && nlog < a.log_cap) logp[nlog] = -1;     // terminator
  }
  g.sync();
}

// one warp per LCP, working set in shared memory
__global__ void __launch_bounds__(256) lcp_warp_kernel(LcpArgs a, int warps_per_block, size_t warp_d, size_t warp_i) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int w = threadIdx.x >> 5;
  double* sm_d = (double*)smem + (size_t)w * warp_d;
  int* sm_i = (int*)((double*)smem + (size_t)warps_per_block * warp_d) + (size_t)w * warp_i;
  WarpGroup g(nullptr);
  const int stride = gridDim.x * warps_per_block;
  for (int b = blockIdx.x * warps_per_block + w; b < a.batch; b += stride) {
    // the next problem of this warp into L2 while this one pivots (its M and q are read twice from global memory: norm, tableau)
    if (b + stride < a.batch) {
      const char* nm = (const char*)(a.M + (size_t)(b + stride) * a.n * a.n);
      const size_t bytes = (size_t)a.n * a.n * sizeof(double);
      for (size_t o = (size_t)g.tid * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(nm + o));
      if (g.tid == 0) asm volatile("prefetch.global.L2 [%0];" :: "l"(a.q + (size_t)(b + stride) * a.n));
    }
    solve_one(g, a, b, sm_d, sm_i);
  }
}

// L lanes per LCP (n <= L): 32 / L problems per warp, each sub-group with its own slice of shared memory
template <int L>
__global__ void __launch_bounds__(256) lcp_subwarp_kernel(LcpArgs a, int warps_per_block, size_t grp_d, size_t grp_i) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int GPW = 32 / L;                                  // groups per warp
  const int gib = (threadIdx.x >> 5) * GPW + ((threadIdx.x & 31) / L), groups_per_block = warps_per_block * GPW;
  double* sm_d = (double*)smem + (size_t)gib * grp_d;
  int* sm_i = (int*)((double*)smem + (size_t)groups_per_block * grp_d) + (size_t)gib * grp_i;
  SubWarpGroup<L> g(nullptr);
  for (int b = blockIdx.x * groups_per_block + gib; b < a.batch; b += gridDim.x * groups_per_block) solve_one(g, a, b, sm_d, sm_i);
}

// one block per LCP, working set in global scratch (L2-resident), reductions through shared memory
__global__ void __launch_bounds__(256) lcp_block_kernel(LcpArgs a) {
  __shared__ double red[4 * 8 + 4];
  BlockGroup<256> g(red);
  double* wd = a.scratch_d + (size_t)blockIdx.x * a.scratch_d_stride;
  int* wi = a.scratch_i + (size_t)blockIdx.x * a.scratch_i_stride;
  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) solve_one(g, a, b, wd, wi);
}

// ---- Lemke for n in the hundreds: the tableau in the distributed shared memory of an 8-CTA cluster ----------------------
// (SURVEY 7.4 option (i).)  At n = 320 the tableau is 0.82 MB: a block that keeps it in global scratch moves 1.64 MB through L2
// per pivot (29 us alone, 110 us when every SM does it).  Here CTA k of a cluster owns the columns [k W, (k + 1) W) of the
// column-major tableau in its own shared memory (W = ceil((n + 2) / 8): 105 KB at n = 320) and updates only those.  Per pivot
// every CTA copies the entering column and the x column from their owners through DSMEM, then runs the ratio test and the basis
// bookkeeping REDUNDANTLY on its private copies -- all CTAs reach the same pivot row without a cross-CTA reduction -- scales its
// own slice of the pivot row and updates its own columns.  Two cluster barriers per pivot: after the remote reads (the owner
// is about to overwrite the entering column) and after the update.  Same arithmetic per entry as lemke_solve: bit-identical
// results, pivots and pivot log.
#define B2M_CL 8
__global__ void __launch_bounds__(256) lcp_cluster_kernel(LcpArgs a, int W) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red[4 * 8 + 4];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int n = a.n, t = 2 * n, tid = threadIdx.x;
  BlockGroup<256> g(red);
  double* Tl = (double*)smem;                       // own columns, column-major, n rows each
  double* dvec = Tl + (size_t)W * n;
  double* xl = dvec + n;                            // private copy of the x column
  double* rl = xl + n;                              // own slice of the scaled pivot row
  int* where = (int*)(rl + W);
  int* bas = where + 2 * n + 1;
  const int c0 = rank * W, c1 = min(c0 + W, n + 2), nown = max(c1 - c0, 0);
  const int xo = (n + 1) / W, xc = (n + 1) - xo * W;            // owner and local index of the x column
  const int n_clusters = gridDim.x / B2M_CL;
  const int MAXITER = min(1000, 50 * n);
  for (int b = blockIdx.x / B2M_CL; b < a.batch; b += n_clusters) {
    const double* M = a.M + (size_t)b * n * n;
    const double* q = a.q + (size_t)b * n;
    double* z = a.z + (size_t)b * n;
    int* logp = a.log ? a.log + (size_t)b * a.log_cap : nullptr;
    int nlog = 0, piv = 0, status = LCP_OK;
    const double nrm = norm_inf_with(g, n, M, n, 0.0, -1.0);
    const double zero_tol = (a.zero_tol > 0.0) ? a.zero_tol : B2M_EPS * nrm * n;
    const double PIV_TOL = (a.piv_tol > 0.0) ? a.piv_tol : B2M_EPS * n * fmax(1.0, nrm);
    double mq = B2M_INF;
    for (int i = tid; i < n; i += 256) mq = fmin(mq, q[i]);
    mq = g.min(mq);
    if (rank == 0) for (int i = tid; i < n; i += 256) z[i] = 0.0;
    bool done = false;
    if (mq > -zero_tol) { status = LCP_TRIVIAL; done = true; }
    else if (!(mq < 0.0)) { status = LCP_OK; done = true; }
    if (!done) {
      for (int e = tid; e < nown * n; e += 256) {
        const int cl = e / n, i = e - cl * n, c = c0 + cl;
        double v;
        if (c < n) { const double mv = M[(size_t)c * n + i]; v = -((c == i) ? mv + 0.0 : mv); }      // as m_at with lambda = 0 (the diagonal takes the addition)
        else if (c == n) v = (q[i] < 0.0) ? -1.0 : 0.0;
        else v = q[i];
        Tl[e] = v;
      }
      for (int i = tid; i < n; i += 256) { xl[i] = q[i]; where[i] = i; where[n + i] = -(i + 1); bas[i] = n + i; }
      if (tid == 0) where[t] = n;
      __syncthreads();
      int r; { double key = B2M_INF; int idx = 0x7fffffff;
        for (int i = tid; i < n; i += 256) { const double x = xl[i]; if (x < key) { key = x; idx = i; } }
        g.min_key_idx(key, idx); r = idx; }
      int s = n, entering = t;
      bool first = true;
      cluster.sync();                                             // every CTA's columns are in place
      for (;;) {
        { const int so = s / W;
          const double* src = cluster.map_shared_rank(Tl, so) + (size_t)(s - so * W) * n;
          const double* xs = cluster.map_shared_rank(Tl, xo) + (size_t)xc * n;
          for (int i = tid; i < n; i += 256) { dvec[i] = src[i]; xl[i] = xs[i]; } }
        cluster.sync();                                           // remote reads done: the owners may overwrite these columns
        if (!first) {
          double theta = B2M_INF;
          for (int i = tid; i < n; i += 256) { const double d = dvec[i]; if (d > PIV_TOL) theta = fmin(theta, (xl[i] + zero_tol) / d); }
          theta = g.min(theta);
          if (theta == B2M_INF) { status = LCP_RAY; break; }
          int lo = 0x7fffffff;
          const int trow = -(where[t] + 1);
          for (int i = tid; i < n; i += 256) {
            const double d = dvec[i];
            if (d > PIV_TOL && xl[i] / d <= theta) { const int key = (i == trow) ? -1 : i; if (key < lo) lo = key; }
          }
          lo = g.min(lo);
          if (lo == 0x7fffffff) { status = LCP_EMPTY_RATIO; break; }
          r = (lo < 0) ? trow : lo;
        }
        const int leaving = bas[r];
        const double p = dvec[r];
        __syncthreads();
        for (int cl = tid; cl < nown; cl += 256) { const int c = c0 + cl; rl[cl] = (c == s) ? 1.0 / p : Tl[(size_t)cl * n + r] / p; }
        if (s >= c0 && s < c1) for (int i = tid; i < n; i += 256) Tl[(size_t)(s - c0) * n + i] = 0.0;
        if (tid == 0) {
          if (rank == 0 && logp && nlog < a.log_cap) logp[nlog] = leaving;
          where[entering] = -(r + 1); where[leaving] = s; bas[r] = entering;
        }
        nlog++;
        __syncthreads();
        {                                                         // a warp per column, lanes down the rows: no index arithmetic per entry, the pivot-row entry is read once per column
          const int lane = tid & 31;
          for (int cl = tid >> 5; cl < nown; cl += 8) {
            const double rv = rl[cl];
            double* col = Tl + (size_t)cl * n;
            int i = lane;
            for (; i + 96 < n; i += 128) {
              const double a0 = col[i], a1 = col[i + 32], a2 = col[i + 64], a3 = col[i + 96];
              const double d0 = dvec[i], d1 = dvec[i + 32], d2 = dvec[i + 64], d3 = dvec[i + 96];
              col[i] = (i == r) ? rv : fma(-d0, rv, a0);
              col[i + 32] = (i + 32 == r) ? rv : fma(-d1, rv, a1);
              col[i + 64] = (i + 64 == r) ? rv : fma(-d2, rv, a2);
              col[i + 96] = (i + 96 == r) ? rv : fma(-d3, rv, a3);
            }
            for (; i < n; i += 32) col[i] = (i == r) ? rv : fma(-dvec[i], rv, col[i]);
          }
        }
        cluster.sync();                                           // the update is visible to the next pivot's remote reads
        if (!first) piv++;
        first = false;
        if (leaving == t) break;
        if (piv >= MAXITER) { status = LCP_MAXITER; break; }
        entering = (leaving < n) ? n + leaving : leaving - n;
        s = where[entering];
      }
      if (status == LCP_OK && rank == xo) {
        __syncthreads();
        for (int i = tid; i < n; i += 256) { const int bb = bas[i]; if (bb < n) z[bb] = Tl[(size_t)xc * n + i]; }
      }
    }
    if (rank == 0 && tid == 0) {
      a.status[b] = status;
      if (a.pivots) a.pivots[b] = piv;
      if (logp && nlog < a.log_cap) logp[nlog] = -1;
    }
    cluster.sync();                                               // nobody starts the next problem while a neighbour still reads this one
  }
}

// self-test of b2m_divn (common.cuh): q = the lock-step division, qref = the compiler's `/`, four pairs per thread
__global__ void div_selftest_kernel(int n, const double* x, const double* y, double* q, double* qref) {
  const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i + 3 >= n) return;
  const double xs[4] = {x[i], x[i + 1], x[i + 2], x[i + 3]}, ys[4] = {y[i], y[i + 1], y[i + 2], y[i + 3]};
  double qs[4];
  b2m_divn<4>(xs, ys, qs);
  for (int k = 0; k < 4; k++) { q[i + k] = qs[k]; qref[i + k] = xs[k] / ys[k]; }
}

size_t cluster_smem(int n) {
  const size_t W = (n + 2 + B2M_CL - 1) / B2M_CL;
  return (W * n + 2 * (size_t)n + W) * sizeof(double) + ((size_t)3 * n + 2) * sizeof(int);
}
bool cluster_fits(int n) { return n >= 16 && cluster_smem(n) <= 227 * 1024 - 2048; }

b200moby_status launch(LcpArgs a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (a.batch < 0 || a.n < 0 || (a.batch > 0 && a.n > 0 && (!a.M || !a.q || !a.z || !a.status))) return b2m_fail(B200MOBY_ERR_INVALID, "null pointer or negative size");
  if (a.batch == 0 || a.n == 0) return B200MOBY_OK;
  int dev = 0, sms = 0;
  B2M_CUDA(cudaGetDevice(&dev));
  B2M_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t wd = stage_doubles(a.mode, a.n) + work_doubles(a.mode, a.n);
  const size_t wi = (work_ints(a.mode, a.n) + 3) & ~(size_t)3;
  const size_t per_warp = wd * sizeof(double) + wi * sizeof(int);
  const size_t MAXS = 227 * 1024;
  const char* sub_env = getenv("B200MOBY_LCP_SUBWARP_NMAX");
  const int sub_nmax = sub_env ? atoi(sub_env) : 8;
  if (a.n <= sub_nmax && a.n <= 16 && a.log == nullptr) {        // several problems per warp (the logged variant keeps the warp loop the parity tests pin)
    const int L = a.n <= 8 ? 8 : 16, gpw = 32 / L, wpb = 8;
    const size_t shmem = per_warp * gpw * wpb;                  // per_warp is the working set of ONE problem
    const void* k = L == 8 ? (const void*)lcp_subwarp_kernel<8> : (const void*)lcp_subwarp_kernel<16>;
    B2M_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAXS));
    int per_sm = 1;
    B2M_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, wpb * 32, shmem));
    if (per_sm < 1) per_sm = 1;
    const int need = (a.batch + wpb * gpw - 1) / (wpb * gpw);
    const int grid = std::min(need, sms * per_sm);
    int wpb_ = wpb; size_t wd_ = wd, wi_ = wi;
    void* args[] = {&a, &wpb_, &wd_, &wi_};
    B2M_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(wpb * 32), args, shmem, stream));
    B2M_CUDA(cudaGetLastError());
    return B200MOBY_OK;
  }
  if (per_warp <= MAXS) {
    B2M_CUDA(cudaFuncSetAttribute(lcp_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAXS));
    // warps per block that put the most warps on an SM (shared memory and registers both limit the blocks: at n = 40 one block
    // of 8 warps fits, but two of 7 do); ties: the wider block
    int wpb = 1, per_sm = 1, best = 0;
    for (int w = (int)std::min<size_t>(8, MAXS / per_warp); w >= 1; w--) {
      int blocks = 0;
      B2M_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, lcp_warp_kernel, w * 32, per_warp * w));
      if (blocks * w > best) { best = blocks * w; wpb = w; per_sm = blocks; }
    }
    // do not launch blocks wider than the batch needs
    while (wpb > 1 && (size_t)(wpb - 1) * sms >= (size_t)a.batch) wpb--;
    const size_t shmem = per_warp * wpb;
    B2M_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lcp_warp_kernel, wpb * 32, shmem));
    if (per_sm < 1) per_sm = 1;
    const int need = (a.batch + wpb - 1) / wpb;
    const int grid = std::min(need, sms * per_sm);
    lcp_warp_kernel<<<grid, wpb * 32, shmem, stream>>>(a, wpb, wd, wi);
    B2M_CUDA(cudaGetLastError());
  } else if (a.mode == MODE_LEMKE && cluster_fits(a.n) && !(getenv("B200MOBY_LCP_CLUSTER") && atoi(getenv("B200MOBY_LCP_CLUSTER")) == 0)) {
    // tableau in the distributed shared memory of an 8-CTA cluster
    int W = (a.n + 2 + B2M_CL - 1) / B2M_CL;
    const size_t shmem = cluster_smem(a.n);
    B2M_CUDA(cudaFuncSetAttribute(lcp_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = B2M_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = shmem; cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(B2M_CL);
    int max_clusters = 0;
    B2M_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, lcp_cluster_kernel, &cfg));
    if (max_clusters < 1) max_clusters = 1;
    cfg.gridDim = dim3(B2M_CL * std::min(a.batch, max_clusters));
    B2M_CUDA(cudaLaunchKernelEx(&cfg, lcp_cluster_kernel, a, W));
    B2M_CUDA(cudaGetLastError());
  } else {
    const int grid = std::min(a.batch, sms * 2);
    a.scratch_d_stride = (wd + 1) & ~(size_t)1;
    a.scratch_i_stride = wi;
    B2M_CUDA(cudaMallocAsync((void**)&a.scratch_d, a.scratch_d_stride * grid * sizeof(double), stream));
    B2M_CUDA(cudaMallocAsync((void**)&a.scratch_i, a.scratch_i_stride * grid * sizeof(int), stream));
    lcp_block_kernel<<<grid, 256, 0, stream>>>(a);
    B2M_CUDA(cudaGetLastError());
    B2M_CUDA(cudaFreeAsync(a.scratch_d, stream));
    B2M_CUDA(cudaFreeAsync(a.scratch_i, stream));
  }
  return B200MOBY_OK;
}

}  // namespace

extern "C" {

b200moby_status b200moby_lcp_lemke_batched(int batch, int n, const double* M, const double* q, double* z, double piv_tol,
                                           double zero_tol, int* status, int* pivots, int* log, int log_cap, void* stream) {
  if (!b2m_have_device()) return b2m_fail(B200MOBY_ERR_NO_DEVICE, "no CUDA device: the hot path has no CPU fallback");
  LcpArgs a; memset(&a, 0, sizeof(a));
  a.batch = batch; a.n = n; a.mode = MODE_LEMKE; a.M = M; a.q = q; a.z = z; a.piv_tol = piv_tol; a.zero_tol = zero_tol;
  a.status = status; a.pivots = pivots; a.log = log; a.log_cap = log_cap;
  return launch(a, stream);
}

b200moby_status b200moby_lcp_fast_batched(int batch, int n, const double* M, const double* q, double* z, int warm,
                                          double zero_tol, int* status, int* pivots, int* log, int log_cap, void* stream) {
  if (!b2m_have_device()) return b2m_fail(B200MOBY_ERR_NO_DEVICE, "no CUDA device: the hot path has no CPU fallback");
  LcpArgs a; memset(&a, 0, sizeof(a));
  a.batch = batch; a.n = n; a.mode = MODE_FAST; a.M = M; a.q = q; a.z = z; a.zero_tol = zero_tol; a.warm = warm;
  a.status = status; a.pivots = pivots; a.log = log; a.log_cap = log_cap;
  return launch(a, stream);
}

b200moby_status b200moby_lcp_lemke_regularized_batched(int batch, int n, const double* M, const double* q, double* z,
                                                       int min_exp, int step_exp, int max_exp, double piv_tol,
                                                       double zero_tol, int* status, int* pivots, void* stream) {
  if (!b2m_have_device()) return b2m_fail(B200MOBY_ERR_NO_DEVICE, "no CUDA device: the hot path has no CPU fallback");
  if (step_exp <= 0) return b2m_fail(B200MOBY_ERR_INVALID, "step_exp must be positive");
  LcpArgs a; memset(&a, 0, sizeof(a));
  a.batch = batch; a.n = n; a.mode = MODE_LEMKE_REG; a.M = M; a.q = q; a.z = z; a.piv_tol = piv_tol; a.zero_tol = zero_tol;
  a.min_exp = min_exp; a.step_exp = step_exp; a.max_exp = max_exp; a.status = status; a.pivots = pivots;
  return launch(a, stream);
}

b200moby_status b200moby_lcp_fast_regularized_batched(int batch, int n, const double* M, const double* q, double* z, int warm,
                                                      int min_exp, int step_exp, int max_exp, double zero_tol,
                                                      int* status, int* pivots, void* stream) {
  if (!b2m_have_device()) return b2m_fail(B200MOBY_ERR_NO_DEVICE, "no CUDA device: the hot path has no CPU fallback");
  if (step_exp <= 0) return b2m_fail(B200MOBY_ERR_INVALID, "step_exp must be positive");
  LcpArgs a; memset(&a, 0, sizeof(a));
  a.batch = batch; a.n = n; a.mode = MODE_FAST_REG; a.M = M; a.q = q; a.z = z; a.zero_tol = zero_tol; a.warm = warm;
  a.min_exp = min_exp; a.step_exp = step_exp; a.max_exp = max_exp; a.status = status; a.pivots = pivots;
  return launch(a, stream);
}

// device buffers + stream of one host-form call, released on every exit path
struct HostFormScratch {
  cudaStream_t s = nullptr;
  void* p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  ~HostFormScratch() {
    if (!s) { for (void* q : p) if (q) cudaFree(q); return; }
    for (void* q : p) if (q) cudaFreeAsync(q, s);
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
  }
};

b200moby_status b200moby_selftest_div(int n, const double* x_dev, const double* y_dev, double* q_dev, double* qref_dev, void* stream) {
  if (!b2m_have_device()) return b2m_fail(B200MOBY_ERR_NO_DEVICE, "no CUDA device: the hot path has no CPU fallback");
  if (n < 0 || (n & 3) || !x_dev || !y_dev || !q_dev || !qref_dev) return b2m_fail(B200MOBY_ERR_INVALID, "n must be a multiple of 4 and the pointers non-null");
  if (n == 0) return B200MOBY_OK;
  div_selftest_kernel<<<(n / 4 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, x_dev, y_dev, q_dev, qref_dev);
  B2M_CUDA(cudaGetLastError());
  return B200MOBY_OK;
}

static b200moby_status host_form(int mode, int batch, int n, const double* M, const double* q, double* z, int warm,
                                 double piv_tol, double zero_tol, int min_exp, int step_exp, int max_exp, int* status,
                                 int* pivots, int device) {
  if (!b2m_have_device()) return b2m_fail(B200MOBY_ERR_NO_DEVICE, "no CUDA device: the hot path has no CPU fallback");
  if (mode < MODE_LEMKE || mode > MODE_FAST_REG) return b2m_fail(B200MOBY_ERR_INVALID, "mode must be 0..3");
  if (batch <= 0 || n <= 0) return B200MOBY_OK;
  if (!M || !q || !z) return b2m_fail(B200MOBY_ERR_INVALID, "null pointer");
  B2M_CUDA(cudaSetDevice(device));
  HostFormScratch sc;
  const size_t nM = (size_t)batch * n * n, nv = (size_t)batch * n;
  B2M_CUDA(cudaStreamCreate(&sc.s));
  cudaStream_t s = sc.s;
  B2M_CUDA(cudaMallocAsync(&sc.p[0], nM * sizeof(double), s));
  B2M_CUDA(cudaMallocAsync(&sc.p[1], nv * sizeof(double), s));
  B2M_CUDA(cudaMallocAsync(&sc.p[2], nv * sizeof(double), s));
  B2M_CUDA(cudaMallocAsync(&sc.p[3], batch * sizeof(int), s));
  B2M_CUDA(cudaMallocAsync(&sc.p[4], batch * sizeof(int), s));
  double *dM = (double*)sc.p[0], *dq = (double*)sc.p[1], *dz = (double*)sc.p[2]; int *dst = (int*)sc.p[3], *dpv = (int*)sc.p[4];
  B2M_CUDA(cudaMemcpyAsync(dM, M, nM * sizeof(double), cudaMemcpyHostToDevice, s));
  B2M_CUDA(cudaMemcpyAsync(dq, q, nv * sizeof(double), cudaMemcpyHostToDevice, s));
  // z is the warm start of the lcp_fast family and is left as given when such a solve fails (LCP.cpp:118-126,192-195):
  // the device copy always starts from the caller's z so that a failed cold solve hands back what it was given
  B2M_CUDA(cudaMemcpyAsync(dz, z, nv * sizeof(double), cudaMemcpyHostToDevice, s));
  b200moby_status r;
  switch (mode) {
    case MODE_LEMKE: r = b200moby_lcp_lemke_batched(batch, n, dM, dq, dz, piv_tol, zero_tol, dst, dpv, nullptr, 0, s); break;
    case MODE_FAST: r = b200moby_lcp_fast_batched(batch, n, dM, dq, dz, warm, zero_tol, dst, dpv, nullptr, 0, s); break;
    case MODE_LEMKE_REG: r = b200moby_lcp_lemke_regularized_batched(batch, n, dM, dq, dz, min_exp, step_exp, max_exp, piv_tol, zero_tol, dst, dpv, s); break;
    default: r = b200moby_lcp_fast_regularized_batched(batch, n, dM, dq, dz, warm, min_exp, step_exp, max_exp, zero_tol, dst, dpv, s); break;
  }
  if (r != B200MOBY_OK) return r;
  B2M_CUDA(cudaMemcpyAsync(z, dz, nv * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (status) B2M_CUDA(cudaMemcpyAsync(status, dst, batch * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (pivots) B2M_CUDA(cudaMemcpyAsync(pivots, dpv, batch * sizeof(int), cudaMemcpyDeviceToHost, s));
  B2M_CUDA(cudaStreamSynchronize(s));
  return B200MOBY_OK;
}

b200moby_status b200moby_lcp_solve_host(int mode, int batch, int n, const double* M, const double* q, double* z, int warm_start,
                                        double piv_tol, double zero_tol, int min_exp, int step_exp, int max_exp, int* status,
                                        int* pivots, int device) {
  return host_form(mode, batch, n, M, q, z, warm_start, piv_tol, zero_tol, min_exp, step_exp, max_exp, status, pivots, device);
}

b200moby_status b200moby_lcp_lemke_host(int batch, int n, const double* M, const double* q, double* z, double piv_tol,
                                        double zero_tol, int* status, int* pivots, int device) {
  return host_form(MODE_LEMKE, batch, n, M, q, z, 0, piv_tol, zero_tol, 0, 1, 0, status, pivots, device);
}
b200moby_status b200moby_lcp_fast_host(int batch, int n, const double* M, const double* q, double* z, int warm,
                                       double zero_tol, int* status, int* pivots, int device) {
  return host_form(MODE_FAST, batch, n, M, q, z, warm, -1.0, zero_tol, 0, 1, 0, status, pivots, device);
}

}  // extern "C"
