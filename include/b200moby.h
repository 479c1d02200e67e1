/*
 * b200moby.h -- C ABI of the B200-native time-stepping contact hot path.
 *
 * This is the drop-in boundary for the one path this repository accelerates
 * (SURVEY.md section 8): narrowphase -> forward dynamics + semi-implicit Euler
 * -> Delassus / LCP assembly -> lcp_fast / Lemke pivoting -> impulse
 * application, run as a batch of independent simulation instances ("envs").
 *
 * Moby itself has no C ABI (its seams are C++ virtuals); each
 * entry point below names the reference interface it replaces (file:line under
 * the Moby tree).  Conventions:
 *   - plain pointers and sizes only; no C++/torch types; no exceptions cross
 *     the boundary; every call returns a b200moby_status;
 *   - one handle per GPU, not thread-safe per handle;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - "_dev" pointers are device pointers on the handle's GPU, everything
 *     else is host memory;
 *   - all arithmetic is IEEE FP64, indices are int32;
 *   - per-env arrays are structure-of-arrays across envs: element k of env e
 *     lives at [k * n_envs + e]  (written [k][env] below);
 *   - dense LCP matrices are column-major, one contiguous n*n block per
 *     problem: M[b*n*n + c*n + r].
 */
#ifndef B200MOBY_H
#define B200MOBY_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200MOBY_ABI_VERSION 3

typedef enum {
  B200MOBY_OK = 0,
  B200MOBY_ERR_INVALID = 1,      /* bad argument */
  B200MOBY_ERR_CUDA = 2,         /* CUDA runtime error (see b200moby_last_error) */
  B200MOBY_ERR_UNSUPPORTED = 3,  /* scene needs a feature outside the hot path */
  B200MOBY_ERR_NO_DEVICE = 4     /* no sm_100 device: there is no CPU fallback */
} b200moby_status;

/* Per-LCP solver status words (written by the kernels; never abort a batch).
 * They mirror the bool / exception cascade of LCP.cpp:212-487. */
enum {
  B200MOBY_LCP_OK = 0,            /* solved, no regularisation */
  B200MOBY_LCP_TRIVIAL = 1,       /* q >= -zero_tol: z = 0 (LCP.cpp:578-584, :89-95) */
  B200MOBY_LCP_RAY = 2,           /* Lemke ray termination (LCP.cpp:892-903) */
  B200MOBY_LCP_MAXITER = 3,       /* iteration cap (LCP.cpp:548, :107) */
  B200MOBY_LCP_SINGULAR = 4,      /* exact zero pivot in the basis solve (LCP.cpp:122, :840) */
  B200MOBY_LCP_EMPTY_RATIO = 5,   /* "zero tolerance too low" (LCP.cpp:946-958) */
  B200MOBY_LCP_UNVERIFIED = 6,    /* regularised wrapper: no lambda passed the checks */
  B200MOBY_LCP_REGULARIZED = 16   /* OK after regularisation: 16 + attempt index (0-based) */
};

/* Shapes the narrowphase handles. */
enum { B200MOBY_SHAPE_NONE = 0, B200MOBY_SHAPE_SPHERE = 1, B200MOBY_SHAPE_BOX = 2, B200MOBY_SHAPE_PLANE = 3,
       /* The rimless wheel of example/rimless-wheel: a body whose CollisionGeometry has no primitive and whose distance,
        * contacts and conservative-advancement step come from the collision-detection plugin (coldet-plugin.cpp:86-137,
        * 214-334; extern "C" factory() :340-346).  Built in as a shape: dims = (R, W, N_SPOKES) (params.h:4-6), spokes in the
        * body's x-z plane, checked against planes only. */
       B200MOBY_SHAPE_WHEEL = 4,
       /* example/contact-constrained-pendulum: a pin joint emulated by six frictionless contacts, again the work of a collision-detection
        * plugin (contact-constrained-pendulum-coldet-plugin.cpp:60-110): PIN on the moving body (dims = the anchor point in its frame),
        * PINWORLD on the fixed one (anchor = its origin). */
       B200MOBY_SHAPE_PIN = 5, B200MOBY_SHAPE_PINWORLD = 6 };

/* Impact models (ImpactConstraintHandler.cpp:122-146). */
enum {
  B200MOBY_MODEL_QP = 0,   /* default build: QP-as-LCP, ImpactConstraintHandlerQP.cpp:94-263 */
  B200MOBY_MODEL_AP = 1    /* -DUSE_AP_MODEL: Anitescu-Potra LCP, ImpactConstraintHandlerLCP.cpp:94-370 */
};

#define B200MOBY_MAX_BODIES 16
#define B200MOBY_MAX_LINKS 16

/* Joint types of a reduced-coordinate articulated body (RevoluteJoint.cpp, PrismaticJoint.cpp): one DoF each. */
enum { B200MOBY_JOINT_REVOLUTE = 1, B200MOBY_JOINT_PRISMATIC = 2 };

/* Forward-dynamics algorithm of the articulated body (RCArticulatedBody.cpp:178-201 fdyn-algorithm=). */
enum {
  B200MOBY_FDYN_FSAB = 0,  /* Featherstone articulated-body algorithm, O(N) */
  B200MOBY_FDYN_CRB = 1    /* composite rigid body + Cholesky, O(N^3): what SDFReader.cpp:931-935 wires for the UR10 */
};

/*
 * One fixed-base RCArticulatedBody per env (include/Moby/RCArticulatedBody.h:43; dynamics in Ravelin's
 * RCArticulatedBodyd).  Its links are bodies [first_body, first_body + n_links) of the scene: link 0 is the base
 * (a disabled body, welded to the world at the pose b200moby_set_state gives it), link i > 0 hangs off
 * parent[i] < i by a one-DoF joint whose coordinate is jq[i-1].  Link body frames sit at the link's centre of mass
 * with axes along the principal inertia axes (shape / dims / mass / inertia / contact parameters come from the body
 * arrays of the scene descriptor, so they may differ per env); the kinematic tree below is shared by all envs.
 * Joint limits and floating bases are outside this round's scope.
 * All arrays are host pointers of n_links entries (entry 0 unused), copied at create time.
 */
typedef struct {
  int n_links;
  int first_body;
  const int*    parent;      /* [link] */
  const int*    joint_type;  /* [link] B200MOBY_JOINT_* */
  const double* joint_axis;  /* [link][3] unit axis in the outboard link's frame */
  const double* loc_parent;  /* [link][3] joint location in the inboard link's (COM) frame */
  const double* loc_child;   /* [link][3] joint location in the outboard link's (COM) frame */
  const double* rel_quat;    /* [link][4] outboard orientation relative to inboard at q = 0 (x y z w) */
  int fdyn_algorithm;        /* B200MOBY_FDYN_* used by the stepped path */
  /* Built-in joint-space controller standing in for the ControlledBody callback (ControlledBody.h:37-40,
   * Simulator.cpp:339-348) with the law of example/ur10/controller.cpp:46-96:
   *   tau_k = kp_k (amp_k sin(freq_k t) - q_k) + kv_k (amp_k cos(freq_k t) - qd_k),  t = Simulator::current_time.
   * [dof] arrays, NULL kp = no controller.  b200moby_set_joint_forces adds a per-env feed-forward term. */
  const double* ctrl_kp; const double* ctrl_kv; const double* ctrl_amp; const double* ctrl_freq;
} b200moby_rc_desc;

/*
 * Batch scene descriptor: `n_bodies` rigid bodies per env (static ones
 * included), one collision geometry per body located at the body frame
 * (primitive pose = identity), `n_envs` independent copies whose parameters may
 * differ per env.  Replaces the object graph XMLReader::read builds
 * (XMLReader.cpp:60-132) for the scenes in BASELINE.json.
 *
 * All arrays are host pointers, copied at create time.
 */
typedef struct {
  int n_envs;
  int n_bodies;
  /* [body][env] */
  const int*    shape;      /* B200MOBY_SHAPE_* */
  const int*    enabled;    /* RigidBody enabled= (0: static) */
  const double* mass;       /* kg; ignored for disabled bodies */
  /* [body][3][env] */
  const double* dims;       /* box: xlen,ylen,zlen; sphere: radius,-,-; plane: -,-,- (plane is y=0 of the body frame, BoxPrimitive.cpp:358, PlanePrimitive.cpp:477); wheel: R,W,N_SPOKES (<= 16) */
  const double* inertia;    /* principal body-frame inertia (InertiaFromPrimitive) */
  /* contact parameters, [body_i*n_bodies + body_j][env] for i<j (ContactParameters.cpp:97-136) */
  const double* mu_coulomb;  /* an island whose contacts all have mu_coulomb >= 100 takes the no-slip model (ImpactConstraintHandler.cpp:122-135,1009-1417) */
  const double* mu_viscous;
  const double* epsilon;
  const double* compliance;
  const int*    NK;         /* friction-cone-edges (>=4, even); 0 = pair disabled (<DisabledPair>) */
  double gravity[3];            /* GravityForce accel= (GravityForce.cpp:32-68) */
  double contact_dist_thresh;   /* ConstraintSimulator.cpp:56, default 1e-6 */
  double min_step_size;         /* TimeSteppingSimulator.cpp:48, default sqrt(eps) */
  const double* min_step_size_env; /* optional [env] override (XML min-step-size, TimeSteppingSimulator.cpp:470-472); NULL = scalar above */
  int    impact_model;          /* B200MOBY_MODEL_* */
  int    stabilization_max_iterations; /* constraint-stabilization-max-iterations (ConstraintStabilization.cpp:53-59): 0 = off (as
                                          example/ur10/ur10.xml:12), < 0 = the reference's default (no limit: until no pair is closer than
                                          sqrt(eps); capped at 100 iterations per step here and counted), > 0 = that many */
  const b200moby_rc_desc* rc;   /* optional articulated body (NULL: free bodies only) */
  /* Working-set bounds per env.  0 = the worst case over all body pairs (every pair in contact at once), which is
   * what small scenes use; many-body scenes (a 10-box stack has 55 pairs but ~40 simultaneous contacts) give the bounds
   * they need.  An env that exceeds them in some step skips that impact solve and is counted in lcp_failures. */
  int max_contacts;
  int max_lcp_n;
} b200moby_scene_desc;

typedef struct b200moby_sim* b200moby_handle;

/* Device-side counters, summed over envs (SURVEY.md section 5 "Metrics"). */
typedef struct {
  long long env_steps;        /* TimeSteppingSimulator::step calls x envs */
  long long mini_steps;       /* do_mini_step calls (TimeSteppingSimulator.cpp:114) */
  long long lcp_solves;       /* impact problems handed to the LCP solver */
  long long lcp_fast_calls;   /* lcp_fast invocations, regularised retries included */
  long long lemke_calls;      /* lcp_lemke invocations, regularised retries included */
  long long pivots;           /* total pivots / iterations */
  long long lcp_failures;     /* LCPSolverException equivalents (ImpactConstraintHandlerQP.cpp:224) */
  long long impact_tol_events;/* ImpactToleranceException equivalents (ImpactConstraintHandler.cpp:153-167) */
  long long contacts;         /* contact constraints generated */
  long long max_lcp_n;        /* largest LCP dimension seen */
  long long pivot_flops;      /* sum over solves of pivots * 2 n (n+1): the algorithmic solver flops of SURVEY.md 8(d) */
  long long ca_iterations;    /* position sub-steps of the conservative-advancement loop (TimeSteppingSimulator.cpp:133-168) */
  long long assembly_flops;   /* F_delassus + F_apply per island solve + F_fd + F_narrow per mini-step (SURVEY.md 8(d) formulas) */
  long long stab_iterations;  /* iterations of ConstraintStabilization::stabilize's loop (ConstraintStabilization.cpp:197-244) */
  long long stab_lcp_solves;  /* frictionless position LCPs solved by determine_dq (ConstraintStabilization.cpp:932-970); their lcp_fast /
                                 lcp_lemke calls and pivots are included in the counters above, failures in lcp_failures */
  long long stab_line_search_failures; /* update_q gave up (t < sqrt(eps), :1186-1187) or the 100-iteration cap was hit */
} b200moby_counters;

const char* b200moby_last_error(void);
int b200moby_abi_version(void);
/* Number of visible sm_100 devices (0 => every compute call returns NO_DEVICE). */
int b200moby_device_count(void);

/* ---- simulator: replaces TimeSteppingSimulator::step (TimeSteppingSimulator.cpp:52-111) ---- */
b200moby_status b200moby_create(const b200moby_scene_desc* desc, int device, b200moby_handle* out);
b200moby_status b200moby_destroy(b200moby_handle h);
/* q: [body][7][env] = x y z qx qy qz qw (Euler coordinates, regress.cpp:78-95);
 * v: [body][6][env] = linear, angular velocity at the COM in a global-aligned frame. Host buffers. */
b200moby_status b200moby_set_state(b200moby_handle h, const double* q, const double* v);
b200moby_status b200moby_get_state(b200moby_handle h, double* q, double* v);
/* Same with device buffers (no host round trip). */
b200moby_status b200moby_set_state_dev(b200moby_handle h, const double* q_dev, const double* v_dev, void* stream);
b200moby_status b200moby_get_state_dev(b200moby_handle h, double* q_dev, double* v_dev, void* stream);
/* Joint state of the articulated body, [dof][env] (RCArticulatedBodyd::get/set_generalized_coordinates_euler /
 * _velocity for a fixed base: the joint q and qd).  Setting it refreshes the link poses and velocities that
 * b200moby_get_state reports.  INVALID when the scene has no articulated body. */
b200moby_status b200moby_set_joint_state(b200moby_handle h, const double* jq, const double* jqd);
b200moby_status b200moby_get_joint_state(b200moby_handle h, double* jq, double* jqd);
b200moby_status b200moby_set_joint_state_dev(b200moby_handle h, const double* jq_dev, const double* jqd_dev, void* stream);
b200moby_status b200moby_get_joint_state_dev(b200moby_handle h, double* jq_dev, double* jqd_dev, void* stream);
/* Generalized joint forces added every mini-step until changed, [dof][env] host buffer (what a ControlledBody
 * controller callback returns, Simulator.cpp:339-348); NULL clears them. */
b200moby_status b200moby_set_joint_forces(b200moby_handle h, const double* tau);
/* n_steps x step(dt) for every env; asynchronous on `stream`. */
b200moby_status b200moby_step(b200moby_handle h, double dt, int n_steps, void* stream);
/* Scheduling knob (results do not depend on it): an env whose LCP pivots within one b200moby_step call exceed
 * `budget` leaves the warp-per-env kernel untouched and is re-run by the block-per-env kernel (same arithmetic, 4x the
 * lanes per pivot), so one hard solve cannot hold a whole step.  <= 0 disables; default 96 (env B200MOBY_PIVOT_BUDGET). */
b200moby_status b200moby_set_pivot_budget(b200moby_handle h, int budget);
b200moby_status b200moby_get_counters(b200moby_handle h, b200moby_counters* out);
b200moby_status b200moby_reset_counters(b200moby_handle h);
/* Kernel launches issued through this handle so far (what bench.py reports as gpu_launches). */
b200moby_status b200moby_get_launch_count(b200moby_handle h, long long* out);
/* Simulated time per env, [env] host buffer (Simulator::current_time). */
b200moby_status b200moby_get_time(b200moby_handle h, double* t);
/* Debug tap: LCP of the last impact solve of each env. MM_dev: [env][nmax*nmax] column-major with
 * leading dimension n[env]; any pointer may be NULL. */
b200moby_status b200moby_get_last_lcp(b200moby_handle h, int* n, double* z, int zcap);

/* Per-kernel profile of b200moby_step: which kernels ran, for how long (CUDA events on each launch's own stream; the
 * impact classes of one round overlap, so the durations can add up to more than the step), how many envs each
 * processed and the algorithmic flops it did (pivots * 2n(n+1) + assembly, the formulas of SURVEY.md 8d).
 * enable != 0 turns the event bracketing on for subsequent steps; reset != 0 clears the accumulators after reading.
 * Synchronises the device.  out may be NULL. */
#define B200MOBY_MAX_KERNELS 20
typedef struct {
  char name[48];
  double ms;            /* summed launch durations */
  long long launches;
  long long envs;       /* envs processed */
  long long flops;      /* algorithmic FP64 flops */
  long long lcp_solves;
  int lcp_nmax;         /* LCP class bound of an impact kernel (0 otherwise) */
  int threads_per_env;
} b200moby_kernel_stat;
typedef struct {
  int n_kernels;
  b200moby_kernel_stat k[B200MOBY_MAX_KERNELS];
} b200moby_kernel_profile;
b200moby_status b200moby_get_kernel_profile(b200moby_handle h, int enable, int reset, b200moby_kernel_profile* out);

/* Debug tap: SM cycles, pivots, executed solver iterations and LCP dimension of each env's most recent impact phase,
 * followed by the cycles of nine phases (load, contacts, islands, problem data, LCP build, lcp_fast, Lemke, apply,
 * store); prof: host buffer [13][env]; reading clears it.  The first call arms the tap. */
b200moby_status b200moby_get_impact_profile(b200moby_handle h, long long* prof);

/* Debug tap: per-env solver statistics accumulated since the previous call -- LCPSolverException equivalents
 * (ImpactConstraintHandlerQP.cpp:224), LCP::lcp_lemke calls, LCP::lcp_fast calls, impact problems solved, pivots (LCP::pivots summed); stat: host
 * buffer [5][env]; reading clears it.  The first call arms the tap.  Used by the parity tests to match the checker env
 * by env. */
b200moby_status b200moby_get_env_stats(b200moby_handle h, int* stat);

/* ---- batched solvers: replace LCP::lcp_lemke / lcp_fast and wrappers (LCP.h:21-27) ----
 * M_dev [batch][n*n] column-major, q_dev [batch][n], z_dev [batch][n] (in: warm start for lcp_fast, out: solution),
 * status_dev [batch], pivots_dev [batch] (may be NULL), pivot_log_dev [batch][log_cap] (may be NULL):
 *   Lemke logs the leaving variable id per pivot (z_i: i, w_i: n+i, artificial: 2n), lcp_fast logs (moved index | 0x40000000 if moved to basic);
 *   a -1 terminates the list when it is shorter than log_cap.
 * piv_tol / zero_tol <= 0 select the reference defaults (LCP.cpp:570-571, :761, :57-58). */
b200moby_status b200moby_lcp_lemke_batched(int batch, int n, const double* M_dev, const double* q_dev, double* z_dev,
                                           double piv_tol, double zero_tol, int* status_dev, int* pivots_dev,
                                           int* pivot_log_dev, int log_cap, void* stream);
b200moby_status b200moby_lcp_fast_batched(int batch, int n, const double* M_dev, const double* q_dev, double* z_dev,
                                          int warm_start, double zero_tol, int* status_dev, int* pivots_dev,
                                          int* pivot_log_dev, int log_cap, void* stream);
/* Regularised wrappers (LCP.cpp:212-350, :353-487): lambda = 10^rf for rf = min_exp; rf < max_exp; rf += step_exp. */
b200moby_status b200moby_lcp_lemke_regularized_batched(int batch, int n, const double* M_dev, const double* q_dev,
                                                       double* z_dev, int min_exp, int step_exp, int max_exp,
                                                       double piv_tol, double zero_tol, int* status_dev,
                                                       int* pivots_dev, void* stream);
b200moby_status b200moby_lcp_fast_regularized_batched(int batch, int n, const double* M_dev, const double* q_dev,
                                                      double* z_dev, int warm_start, int min_exp, int step_exp,
                                                      int max_exp, double zero_tol, int* status_dev, int* pivots_dev,
                                                      void* stream);
/* Host-buffer convenience forms: H2D, solve, D2H, synchronise (the call a Moby LCP object would make). */
b200moby_status b200moby_lcp_lemke_host(int batch, int n, const double* M, const double* q, double* z, double piv_tol,
                                        double zero_tol, int* status, int* pivots, int device);
b200moby_status b200moby_lcp_fast_host(int batch, int n, const double* M, const double* q, double* z, int warm_start,
                                       double zero_tol, int* status, int* pivots, int device);

/* Self-test hook: the kernels form the quotients of Lemke's ratio test several at a time with their own IEEE division
 * (moby_b200/csrc/common.cuh b2m_divn); q_dev receives that division of x_dev[i] / y_dev[i], qref_dev the compiler's.  n must
 * be a multiple of 4.  The two must be bit-identical (tests/test_gpu_lcp.py). */
b200moby_status b200moby_selftest_div(int n, const double* x_dev, const double* y_dev, double* q_dev, double* qref_dev, void* stream);

/* All four solvers behind one host-buffer entry point: mode 0 lcp_lemke, 1 lcp_fast, 2 lcp_lemke_regularized,
 * 3 lcp_fast_regularized (min_exp/step_exp/max_exp used by modes 2 and 3 only).  This is what the C++ facade's
 * Moby::LCP methods call (include/b200moby.hpp). */
b200moby_status b200moby_lcp_solve_host(int mode, int batch, int n, const double* M, const double* q, double* z, int warm_start,
                                        double piv_tol, double zero_tol, int min_exp, int step_exp, int max_exp, int* status,
                                        int* pivots, int device);

/* ---- stage kernels, exposed for parity tests and for callers that keep Moby's own step loop ---- */
/* Forward dynamics + velocity half of semi-implicit Euler for free rigid bodies
 * (Simulator.cpp:319-350,482-602; TimeSteppingSimulator.cpp:181-192). q_dev [body][7][env], v_dev [body][6][env] in/out. */
b200moby_status b200moby_fwd_dyn_batched(b200moby_handle h, const double* q_dev, double* v_dev, double dt, void* stream);
/* Forward dynamics of the articulated body, one thread per env with the spatial recursions in registers
 * (Simulator.cpp:544-553 -> RCArticulatedBodyd::calc_fwd_dyn; algorithm: B200MOBY_FDYN_*).
 * jq_dev, jqd_dev, tau_dev (may be NULL), qdd_dev: [dof][env]. */
b200moby_status b200moby_rc_fwd_dyn_batched(b200moby_handle h, int algorithm, const double* jq_dev, const double* jqd_dev,
                                            const double* tau_dev, double* qdd_dev, void* stream);
/* Joint-space inertia H(q) (RCArticulatedBodyd::get_generalized_inertia), H_dev [dof*dof][env] column-major. */
b200moby_status b200moby_rc_inertia_batched(b200moby_handle h, const double* jq_dev, double* H_dev, void* stream);
/* Narrowphase for every body pair (CCD.inl:3-82 and leaves). Outputs, per env, up to `cap` contacts:
 * count_dev [env]; point/normal/tan1/tan2 [cap][3][env]; pair_dev [cap][env] = body1*n_bodies+body2; dist_dev [cap][env]. */
b200moby_status b200moby_find_contacts_batched(b200moby_handle h, const double* q_dev, const double* v_dev, int cap, int* count_dev,
                                               double* point_dev, double* normal_dev, double* tan1_dev,
                                               double* tan2_dev, int* pair_dev, double* dist_dev, void* stream);
/* The same for the simulator's current state with HOST output buffers (synchronous; what ConstraintSimulator::get_rigid_constraints
 * and the constraint callbacks of ConstraintSimulator.h:51-68 need on the host; a slow path by design). */
b200moby_status b200moby_find_contacts_host(b200moby_handle h, int cap, int* count, double* point, double* normal, double* tan1,
                                            double* tan2, int* pair, double* dist);
/* Delassus / LCP assembly for the contacts found at (q,v): writes MM [env][nmax*nmax] (column-major, ld = n[env]),
 * qq [env][nmax], n_dev [env] (ImpactConstraintHandler.cpp:1898-2166 + ImpactConstraintHandlerQP.cpp:271-497
 * or ImpactConstraintHandlerLCP.cpp:94-310). Only the first island of each env is assembled. */
b200moby_status b200moby_delassus_batched(b200moby_handle h, const double* q_dev, const double* v_dev, int nmax,
                                          double* MM_dev, double* qq_dev, int* n_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MOBY_H */
