// Simulator handle, the fused env-step kernel and the stage kernels behind the C ABI (include/b200moby.h).
// b200moby_step replaces TimeSteppingSimulator::step (Moby src/TimeSteppingSimulator.cpp:52-111) for a batch of
// independent envs: one warp per env, the whole mini-step loop (narrowphase, conservative advancement, forward
// dynamics, assembly, LCP solve, impulses) runs out of shared memory; HBM sees the state and the warm start only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "friction_table.h"
#include "host_util.h"
#include "sim_device.cuh"

using namespace b2m;

struct b200moby_sim {
  int device = 0;
  int n_envs = 0, nb = 0, cmax = 0, nmax = 0, npmax = 0;
  SimParams P;
  std::vector<void*> allocs;
  size_t env_d = 0, env_i = 0;
  int wpb = 1, grid = 1, sms = 148;
  size_t shmem = 0;
  bool taps = false;
};

namespace {

__device__ void commit_counters(const SimParams& P, const unsigned long long* lc) {
  for (int k = 0; k < CNT_COUNT; k++) {
    if (k == CNT_MAX_N) atomicMax(P.counters + k, lc[k]);
    else if (lc[k]) atomicAdd(P.counters + k, lc[k]);
  }
}

// One env per warp-sized block.  Envs whose solver work exceeds the pivot budget are queued for step_block_kernel.
__global__ void __launch_bounds__(32) step_warp_kernel(SimParams P, double dt, int n_steps, size_t env_d) {
  extern __shared__ __align__(16) unsigned char smem[];
  EnvMem m;
  env_carve(m, (double*)smem, (int*)((double*)smem + env_d), P.nb, P.cmax, P.nmax, P.npmax);
  WarpGroup g(nullptr);
  unsigned long long lc[CNT_COUNT];
  for (int e = blockIdx.x; e < P.n_envs; e += gridDim.x) {
    for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
    EnvCtx cx; cx.limit = P.pivot_budget > 0; cx.budget = P.pivot_budget;
    const bool done = env_run(g, P, e, m, dt, n_steps, lc, cx);
    if (g.tid == 0) {
      if (done) commit_counters(P, lc);
      else P.defer_list[atomicAdd(P.defer_count, 1)] = e;
    }
    g.sync();
  }
}

// The deferred envs again, from their untouched stored state, with a whole 128-thread block per env: the same code and
// arithmetic (reductions are order-independent), four times the lanes on every pivot.
#define B2M_BLOCK_THREADS 128
__global__ void __launch_bounds__(B2M_BLOCK_THREADS) step_block_kernel(SimParams P, double dt, int n_steps, size_t env_d) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red[4 * (B2M_BLOCK_THREADS / 32) + 4];
  EnvMem m;
  env_carve(m, (double*)smem, (int*)((double*)smem + env_d), P.nb, P.cmax, P.nmax, P.npmax);
  BlockGroup<B2M_BLOCK_THREADS> g(red);
  unsigned long long lc[CNT_COUNT];
  const int count = *P.defer_count;
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    const int e = P.defer_list[i];
    for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
    EnvCtx cx; cx.limit = false; cx.budget = 0;
    env_run(g, P, e, m, dt, n_steps, lc, cx);
    if (g.tid == 0) commit_counters(P, lc);
    g.sync();
  }
}

// stage kernels: same device functions, one warp per env, results written out instead of carried on
enum { STAGE_FWD_DYN = 0, STAGE_CONTACTS = 1, STAGE_DELASSUS = 2 };
struct StageOut {
  double dt;
  int cap; int* count; double* point; double* normal; double* tan1; double* tan2; int* pair; double* dist;
  int nmax_out; double* MM; double* qq; int* n;
};

__global__ void __launch_bounds__(256) stage_warp_kernel(SimParams P, int stage, StageOut o, int wpb, size_t env_d, size_t env_i) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int w = threadIdx.x >> 5;
  double* sd = (double*)smem + (size_t)w * env_d;
  int* si = (int*)((double*)smem + (size_t)wpb * env_d) + (size_t)w * env_i;
  EnvMem m;
  env_carve(m, sd, si, P.nb, P.cmax, P.nmax, P.npmax);
  WarpGroup g(nullptr);
  unsigned long long lc[CNT_COUNT];
  for (int k = 0; k < CNT_COUNT; k++) lc[k] = 0;
  const int ne = P.n_envs;
  for (int e = blockIdx.x * wpb + w; e < ne; e += gridDim.x * wpb) {
    env_load(g, P, e, m);
    if (stage == STAGE_FWD_DYN) {
      fwd_dyn_integrate_velocity(g, P, m, o.dt);
      for (int k = g.tid; k < 3 * P.nb; k += 32) { const int b = k / 3, c = k - 3 * b; if (m.ben[b]) { P.v[((size_t)b * 6 + c) * ne + e] = m.bvl[k]; P.v[((size_t)b * 6 + 3 + c) * ne + e] = m.bva[k]; } }
    } else {
      calc_pairwise_distances(g, m);
      find_unilateral_constraints(g, P, e, m, lc);
      const int ncon = m.scal[S_NCON];
      if (stage == STAGE_CONTACTS) {
        if (g.tid == 0) o.count[e] = ncon;
        for (int c = g.tid; c < ncon && c < o.cap; c += 32) {
          for (int k = 0; k < 3; k++) {
            o.point[((size_t)c * 3 + k) * ne + e] = m.cp[3 * c + k]; o.normal[((size_t)c * 3 + k) * ne + e] = m.cnrm[3 * c + k];
            o.tan1[((size_t)c * 3 + k) * ne + e] = m.ct1[3 * c + k]; o.tan2[((size_t)c * 3 + k) * ne + e] = m.ct2[3 * c + k];
          }
          o.pair[(size_t)c * ne + e] = m.cb1[c] * P.nb + m.cb2[c];
          o.dist[(size_t)c * ne + e] = m.cdist[c];
        }
      } else {
        int n = 0;
        if (ncon > 0) {
          if (g.tid == 0) {            // all contacts of the env as one island, generation order (matches the checker's helper)
            for (int c = 0; c < ncon; c++) m.icon[c] = c;
            int gc = 0;
            for (int b = 0; b < P.nb; b++) {
              bool in = false;
              for (int c = 0; c < ncon; c++) if (m.cb1[c] == b || m.cb2[c] == b) in = true;
              if (in && m.ben[b]) { m.gcoff[b] = gc; gc += 6; } else m.gcoff[b] = -1;
            }
            m.scal[S_NC] = ncon; m.scal[S_NGC] = gc;
          }
          g.sync();
          compute_problem_data(g, P, m);
          n = (P.model == 1) ? build_ap_lcp(g, P, m) : build_qp_lcp(g, P, m);
          if (n <= P.nmax && n <= o.nmax_out) {
            for (int t = g.tid; t < n * n; t += 32) o.MM[(size_t)e * o.nmax_out * o.nmax_out + t] = m.MM[t];
            for (int i = g.tid; i < n; i += 32) o.qq[(size_t)e * o.nmax_out + i] = m.qq[i];
          }
        }
        if (g.tid == 0) o.n[e] = n;
      }
    }
    g.sync();
  }
}

// set_generalized_coordinates_euler normalises the quaternion it is given (Ravelin RigidBodyd); done once here so
// the kernels can trust the stored state.
__global__ void normalize_quat_kernel(double* q, int nb, int ne) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb * ne) return;
  const int b = t / ne, e = t - b * ne;
  double* p = q + ((size_t)b * 7 + 3) * ne + e;
  const double x = p[0], y = p[ne], z = p[2 * (size_t)ne], w = p[3 * (size_t)ne];
  const double nrm = sqrt(x * x + y * y + z * z + w * w);
  p[0] = x / nrm; p[ne] = y / nrm; p[2 * (size_t)ne] = z / nrm; p[3 * (size_t)ne] = w / nrm;
}

template <class T>
b200moby_status dev_copy(b200moby_sim* h, const T* src, size_t count, const T** dst) {
  T* d = nullptr;
  B2M_CUDA(cudaMalloc((void**)&d, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(d);
  if (count) B2M_CUDA(cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *dst = d;
  return B200MOBY_OK;
}
template <class T>
b200moby_status dev_zero(b200moby_sim* h, size_t count, T** dst) {
  T* d = nullptr;
  B2M_CUDA(cudaMalloc((void**)&d, std::max<size_t>(count, 1) * sizeof(T)));
  h->allocs.push_back(d);
  B2M_CUDA(cudaMemset(d, 0, std::max<size_t>(count, 1) * sizeof(T)));
  *dst = d;
  return B200MOBY_OK;
}

b200moby_status plan_launch(b200moby_sim* h, const void* kernel) {
  int sms = 0;
  B2M_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  h->env_d = (env_doubles(h->nb, h->cmax, h->nmax, h->npmax) + 1) & ~(size_t)1;
  h->env_i = (env_ints(h->nb, h->cmax, h->nmax, h->npmax) + 3) & ~(size_t)3;
  const size_t per_warp = h->env_d * sizeof(double) + h->env_i * sizeof(int);
  const size_t MAXS = 227 * 1024;
  if (per_warp > MAXS)
    return b2m_fail(B200MOBY_ERR_UNSUPPORTED, "env working set (%zu bytes, LCP n <= %d) exceeds one SM's shared memory; the block-per-env path is not built yet", per_warp, h->nmax);
  // One env per 32-thread block: envs differ wildly in work (conservative-advancement sub-steps, solver retries), so
  // the hardware block scheduler doing the load balancing beats any static env->warp assignment; shared memory,
  // not threads, limits residency (227 KB / per-env working set blocks per SM).
  h->wpb = 1;
  h->shmem = per_warp;
  B2M_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAXS));
  B2M_CUDA(cudaFuncSetAttribute(step_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAXS));
  h->grid = h->n_envs;
  h->sms = sms;
  return B200MOBY_OK;
}

}  // namespace

extern "C" {

b200moby_status b200moby_create(const b200moby_scene_desc* d, int device, b200moby_handle* out) {
  if (!d || !out) return b2m_fail(B200MOBY_ERR_INVALID, "null descriptor");
  if (!b2m_have_device()) return b2m_fail(B200MOBY_ERR_NO_DEVICE, "no CUDA device: the hot path has no CPU fallback");
  if (d->n_envs <= 0 || d->n_bodies <= 0 || d->n_bodies > B200MOBY_MAX_BODIES) return b2m_fail(B200MOBY_ERR_INVALID, "n_envs > 0 and 0 < n_bodies <= %d required", B200MOBY_MAX_BODIES);
  if (d->stabilization_max_iterations != 0) return b2m_fail(B200MOBY_ERR_UNSUPPORTED, "constraint stabilization is not on the accelerated path: set constraint-stabilization-max-iterations=0");
  B2M_CUDA(cudaSetDevice(device));
  const int ne = d->n_envs, nb = d->n_bodies;
  // validate and size
  int cmax = 0, nmax = 0, npmax = 0;
  {
    std::vector<int> sh(nb), en(nb), nk(nb * nb);
    for (int e = 0; e < ne; e++) {
      for (int b = 0; b < nb; b++) { sh[b] = d->shape[(size_t)b * ne + e]; en[b] = d->enabled[(size_t)b * ne + e]; }
      for (int i = 0; i < nb; i++) for (int j = i + 1; j < nb; j++) {
        const int k = d->NK[((size_t)i * nb + j) * ne + e];
        if (k != 0 && (k < 4 || k > B2M_NKMAX || (k & 1))) return b2m_fail(B200MOBY_ERR_INVALID, "friction-cone-edges must be even and in [4,%d] (ContactParameters.cpp:129-136); got %d", B2M_NKMAX, k);
        nk[i * nb + j] = k;
      }
      int c, n, np;
      b2m_env_bounds(nb, sh.data(), en.data(), nk.data(), d->impact_model, c, n, np);
      for (int i = 0; i < nb; i++) for (int j = i + 1; j < nb; j++)
        if (nk[i * nb + j] && sh[i] == 2 && sh[j] == 2 && (en[i] || en[j])) return b2m_fail(B200MOBY_ERR_UNSUPPORTED, "box-box narrowphase is not on the accelerated path yet; disable the pair or use spheres");
      cmax = std::max(cmax, c); nmax = std::max(nmax, n); npmax = std::max(npmax, np);
    }
  }
  b200moby_sim* h = new b200moby_sim;
  h->device = device; h->n_envs = ne; h->nb = nb; h->cmax = std::max(cmax, 1); h->nmax = std::max(nmax, 1); h->npmax = std::max(npmax, 1);
  SimParams& P = h->P;
  memset(&P, 0, sizeof(P));
  P.n_envs = ne; P.nb = nb; P.cmax = h->cmax; P.nmax = h->nmax; P.npmax = h->npmax; P.model = d->impact_model;
  b200moby_status st;
#define TRY(x) if ((st = (x)) != B200MOBY_OK) { b200moby_destroy(h); return st; }
  TRY(dev_copy(h, d->shape, (size_t)nb * ne, &P.shape));
  TRY(dev_copy(h, d->enabled, (size_t)nb * ne, &P.enabled));
  TRY(dev_copy(h, d->mass, (size_t)nb * ne, &P.mass));
  TRY(dev_copy(h, d->dims, (size_t)nb * 3 * ne, &P.dims));
  TRY(dev_copy(h, d->inertia, (size_t)nb * 3 * ne, &P.inertia));
  TRY(dev_copy(h, d->mu_coulomb, (size_t)nb * nb * ne, &P.mu_c));
  TRY(dev_copy(h, d->mu_viscous, (size_t)nb * nb * ne, &P.mu_v));
  TRY(dev_copy(h, d->epsilon, (size_t)nb * nb * ne, &P.eps));
  TRY(dev_copy(h, d->compliance, (size_t)nb * nb * ne, &P.compliance));
  TRY(dev_copy(h, d->NK, (size_t)nb * nb * ne, &P.NK));
  std::vector<double> tab = b2m_friction_table();
  TRY(dev_copy(h, tab.data(), tab.size(), &P.fr_tab));
  P.gx = d->gravity[0]; P.gy = d->gravity[1]; P.gz = d->gravity[2];
  P.contact_dist_thresh = d->contact_dist_thresh; P.min_step_size = d->min_step_size;
  if (d->min_step_size_env) TRY(dev_copy(h, d->min_step_size_env, (size_t)ne, &P.min_step_env));
  TRY(dev_zero(h, (size_t)nb * 7 * ne, &P.q));
  TRY(dev_zero(h, (size_t)nb * 6 * ne, &P.v));
  TRY(dev_zero(h, (size_t)ne, &P.time));
  TRY(dev_zero(h, (size_t)h->nmax * ne, &P.zlast));
  TRY(dev_zero(h, (size_t)ne, &P.zlast_n));
  TRY(dev_zero(h, (size_t)CNT_COUNT, &P.counters));
  TRY(dev_zero(h, (size_t)ne, &P.defer_list));
  TRY(dev_zero(h, (size_t)1, &P.defer_count));
  P.pivot_budget = 96;
  if (const char* s = getenv("B200MOBY_PIVOT_BUDGET")) P.pivot_budget = atoi(s);
  TRY(plan_launch(h, (const void*)step_warp_kernel));
#undef TRY
  *out = h;
  return B200MOBY_OK;
}

b200moby_status b200moby_destroy(b200moby_handle h) {
  if (!h) return B200MOBY_OK;
  cudaSetDevice(h->device);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return B200MOBY_OK;
}

b200moby_status b200moby_set_state(b200moby_handle h, const double* q, const double* v) {
  if (!h || !q || !v) return b2m_fail(B200MOBY_ERR_INVALID, "null argument");
  B2M_CUDA(cudaSetDevice(h->device));
  B2M_CUDA(cudaMemcpy(h->P.q, q, sizeof(double) * h->nb * 7 * h->n_envs, cudaMemcpyHostToDevice));
  B2M_CUDA(cudaMemcpy(h->P.v, v, sizeof(double) * h->nb * 6 * h->n_envs, cudaMemcpyHostToDevice));
  normalize_quat_kernel<<<(h->nb * h->n_envs + 255) / 256, 256>>>(h->P.q, h->nb, h->n_envs);
  B2M_CUDA(cudaGetLastError());
  return B200MOBY_OK;
}
b200moby_status b200moby_get_state(b200moby_handle h, double* q, double* v) {
  if (!h) return b2m_fail(B200MOBY_ERR_INVALID, "null handle");
  B2M_CUDA(cudaSetDevice(h->device));
  if (q) B2M_CUDA(cudaMemcpy(q, h->P.q, sizeof(double) * h->nb * 7 * h->n_envs, cudaMemcpyDeviceToHost));
  if (v) B2M_CUDA(cudaMemcpy(v, h->P.v, sizeof(double) * h->nb * 6 * h->n_envs, cudaMemcpyDeviceToHost));
  return B200MOBY_OK;
}
b200moby_status b200moby_set_state_dev(b200moby_handle h, const double* q, const double* v, void* stream) {
  if (!h || !q || !v) return b2m_fail(B200MOBY_ERR_INVALID, "null argument");
  B2M_CUDA(cudaMemcpyAsync(h->P.q, q, sizeof(double) * h->nb * 7 * h->n_envs, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  B2M_CUDA(cudaMemcpyAsync(h->P.v, v, sizeof(double) * h->nb * 6 * h->n_envs, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  normalize_quat_kernel<<<(h->nb * h->n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->P.q, h->nb, h->n_envs);
  B2M_CUDA(cudaGetLastError());
  return B200MOBY_OK;
}
b200moby_status b200moby_get_state_dev(b200moby_handle h, double* q, double* v, void* stream) {
  if (!h) return b2m_fail(B200MOBY_ERR_INVALID, "null handle");
  if (q) B2M_CUDA(cudaMemcpyAsync(q, h->P.q, sizeof(double) * h->nb * 7 * h->n_envs, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (v) B2M_CUDA(cudaMemcpyAsync(v, h->P.v, sizeof(double) * h->nb * 6 * h->n_envs, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return B200MOBY_OK;
}

b200moby_status b200moby_step(b200moby_handle h, double dt, int n_steps, void* stream) {
  if (!h || !(dt > 0.0) || n_steps < 0) return b2m_fail(B200MOBY_ERR_INVALID, "bad step arguments");
  if (n_steps == 0) return B200MOBY_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->P.pivot_budget > 0) B2M_CUDA(cudaMemsetAsync(h->P.defer_count, 0, sizeof(int), s));
  step_warp_kernel<<<h->grid, 32, h->shmem, s>>>(h->P, dt, n_steps, h->env_d);
  B2M_CUDA(cudaGetLastError());
  if (h->P.pivot_budget > 0) {
    step_block_kernel<<<h->sms * 2, B2M_BLOCK_THREADS, h->shmem, s>>>(h->P, dt, n_steps, h->env_d);
    B2M_CUDA(cudaGetLastError());
  }
  return B200MOBY_OK;
}

b200moby_status b200moby_set_pivot_budget(b200moby_handle h, int budget) {
  if (!h) return b2m_fail(B200MOBY_ERR_INVALID, "null handle");
  h->P.pivot_budget = budget;
  return B200MOBY_OK;
}

b200moby_status b200moby_get_counters(b200moby_handle h, b200moby_counters* out) {
  if (!h || !out) return b2m_fail(B200MOBY_ERR_INVALID, "null argument");
  unsigned long long c[CNT_COUNT];
  B2M_CUDA(cudaSetDevice(h->device));
  B2M_CUDA(cudaMemcpy(c, h->P.counters, sizeof(c), cudaMemcpyDeviceToHost));
  out->env_steps = c[CNT_ENV_STEPS]; out->mini_steps = c[CNT_MINI_STEPS]; out->lcp_solves = c[CNT_LCP_SOLVES];
  out->lcp_fast_calls = c[CNT_FAST_CALLS]; out->lemke_calls = c[CNT_LEMKE_CALLS]; out->pivots = c[CNT_PIVOTS];
  out->lcp_failures = c[CNT_LCP_FAIL] + c[CNT_OVERFLOW]; out->impact_tol_events = c[CNT_IMPACT_TOL]; out->contacts = c[CNT_CONTACTS];
  out->max_lcp_n = c[CNT_MAX_N]; out->pivot_flops = c[CNT_PIVOT_FLOPS]; out->assembly_flops = c[CNT_ASM_FLOPS]; out->ca_iterations = c[CNT_CA_ITERS];
  return B200MOBY_OK;
}
b200moby_status b200moby_reset_counters(b200moby_handle h) {
  if (!h) return b2m_fail(B200MOBY_ERR_INVALID, "null handle");
  B2M_CUDA(cudaSetDevice(h->device));
  B2M_CUDA(cudaMemset(h->P.counters, 0, sizeof(unsigned long long) * CNT_COUNT));
  return B200MOBY_OK;
}
b200moby_status b200moby_get_time(b200moby_handle h, double* t) {
  if (!h || !t) return b2m_fail(B200MOBY_ERR_INVALID, "null argument");
  B2M_CUDA(cudaSetDevice(h->device));
  B2M_CUDA(cudaMemcpy(t, h->P.time, sizeof(double) * h->n_envs, cudaMemcpyDeviceToHost));
  return B200MOBY_OK;
}

// Debug tap: after this call every impact solve records its LCP; returns the last one per env.
// n: [env]; z: [env][zcap] (host).  First call only arms the tap (returns zeros).
b200moby_status b200moby_get_last_lcp(b200moby_handle h, int* n, double* z, int zcap) {
  if (!h) return b2m_fail(B200MOBY_ERR_INVALID, "null handle");
  B2M_CUDA(cudaSetDevice(h->device));
  if (!h->taps) {
    b200moby_status st;
    if ((st = dev_zero(h, (size_t)h->n_envs * h->nmax * h->nmax, &h->P.tap_MM)) != B200MOBY_OK) return st;
    if ((st = dev_zero(h, (size_t)h->n_envs * h->nmax, &h->P.tap_qq)) != B200MOBY_OK) return st;
    if ((st = dev_zero(h, (size_t)h->n_envs * h->nmax, &h->P.tap_z)) != B200MOBY_OK) return st;
    if ((st = dev_zero(h, (size_t)h->n_envs, &h->P.tap_n)) != B200MOBY_OK) return st;
    h->taps = true;
  }
  if (n) B2M_CUDA(cudaMemcpy(n, h->P.tap_n, sizeof(int) * h->n_envs, cudaMemcpyDeviceToHost));
  if (z) {
    std::vector<double> tmp((size_t)h->n_envs * h->nmax);
    B2M_CUDA(cudaMemcpy(tmp.data(), h->P.tap_z, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
    for (int e = 0; e < h->n_envs; e++) for (int i = 0; i < zcap; i++) z[(size_t)e * zcap + i] = (i < h->nmax) ? tmp[(size_t)e * h->nmax + i] : 0.0;
  }
  return B200MOBY_OK;
}

static b200moby_status run_stage(b200moby_handle h, int stage, const double* q, const double* v, StageOut o, void* stream) {
  if (!h || !q || !v) return b2m_fail(B200MOBY_ERR_INVALID, "null argument");
  static bool attr_set = false;
  if (!attr_set) { B2M_CUDA(cudaFuncSetAttribute(stage_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_set = true; }
  SimParams P = h->P;
  P.q = const_cast<double*>(q); P.v = const_cast<double*>(v);
  stage_warp_kernel<<<h->grid, h->wpb * 32, h->shmem, (cudaStream_t)stream>>>(P, stage, o, h->wpb, h->env_d, h->env_i);
  B2M_CUDA(cudaGetLastError());
  return B200MOBY_OK;
}

b200moby_status b200moby_fwd_dyn_batched(b200moby_handle h, const double* q, double* v, double dt, void* stream) {
  StageOut o; memset(&o, 0, sizeof(o)); o.dt = dt;
  return run_stage(h, STAGE_FWD_DYN, q, v, o, stream);
}
b200moby_status b200moby_find_contacts_batched(b200moby_handle h, const double* q, const double* v, int cap, int* count, double* point,
                                               double* normal, double* tan1, double* tan2, int* pair, double* dist, void* stream) {
  if (!count || !point || !normal || !tan1 || !tan2 || !pair || !dist || cap <= 0) return b2m_fail(B200MOBY_ERR_INVALID, "null output");
  StageOut o; memset(&o, 0, sizeof(o));
  o.cap = cap; o.count = count; o.point = point; o.normal = normal; o.tan1 = tan1; o.tan2 = tan2; o.pair = pair; o.dist = dist;
  return run_stage(h, STAGE_CONTACTS, q, v, o, stream);
}
b200moby_status b200moby_delassus_batched(b200moby_handle h, const double* q, const double* v, int nmax, double* MM, double* qq,
                                          int* n, void* stream) {
  if (!MM || !qq || !n || nmax <= 0) return b2m_fail(B200MOBY_ERR_INVALID, "null output");
  StageOut o; memset(&o, 0, sizeof(o));
  o.nmax_out = nmax; o.MM = MM; o.qq = qq; o.n = n;
  return run_stage(h, STAGE_DELASSUS, q, v, o, stream);
}

}  // extern "C"
