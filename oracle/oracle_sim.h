// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into the product library.
//
// Single-threaded CPU restatement of Moby's time-stepping contact path for free rigid bodies with
// sphere / box / plane geometry.  One Sim == one Moby
// TimeSteppingSimulator.  Follows, statement by statement where the source is in the reference tree:
//   TimeSteppingSimulator.cpp:52-222,272-331,433-455   step / do_mini_step / CA loop
//   ConstraintSimulator.cpp:298-355,450-537            constraint handling entry, distances, contacts
//   CCD.cpp:122-460,585-607 ; CCD.inl:3-82,805-886,1165-1259   conservative advancement, narrowphase
//   UnilateralConstraint.cpp:695-747,940-1225,1387-1446       constraint velocity, islands, tangents
//   ImpactConstraintHandler.cpp:96-168,298-626,1590-2166      model dispatch, problem data, restitution
//   ImpactConstraintHandlerQP.cpp:94-497 ; ImpactConstraintHandlerLCP.cpp:36-370   QP-as-LCP and A-P LCP
// and restates from published semantics what lives in the absent, un-pinned Ravelin dependency
// (rigid-body Newton-Euler, quaternion kinematics, spatial transforms, determine_orthonormal_basis):
// PARITY UNPINNED for those parts; they are pinned only loosely by regress/sitting-box.dat and
// regress/sphere-stack.dat (6 significant digits).
//
// Rules adopted where the reference is history-, allocation- or rand()-dependent (SURVEY.md 8a'):
//   H1 zlast is zero-filled on size change; H2 lowest-index tie rule (glibc-rand optional);
//   H3 restitution re-applies friction (literal); H4 bodies by scene index, pairs lexicographic,
//   contacts in generation order; H6 non-logging create_contact; H8 collinearity scan tests points 0,1,2;
//   H10 the no-slip path's trailing update_from_stacked(_epd, _z) with the QP handler's stale _z is skipped;
//   H7 stabilization runs while min dist < +NEAR_ZERO (as coded, ConstraintStabilization.cpp:58-59,197);
//   H12 a stabilization LCP that neither lcp_fast nor lcp_lemke_regularized solves contributes dq = 0 (the reference
//       ignores lcp_lemke_regularized's return value and uses whatever z it left, :961-962); the stabilization loop is
//       also capped at 100 iterations per step (the reference's default is unbounded): both are counted.
#pragma once
#include <vector>
#include "oracle_lcp.h"
#include "oracle_rc.h"

namespace oracle {

struct V3 {
  double x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return (&x)[i]; }
  double operator[](int i) const { return (&x)[i]; }
};

enum Shape { SHAPE_NONE = 0, SHAPE_SPHERE = 1, SHAPE_BOX = 2, SHAPE_PLANE = 3,
             SHAPE_WHEEL = 4,
             // example/contact-constrained-pendulum: a pin joint emulated by six frictionless contacts (its collision-detection plugin):
             // SHAPE_PIN on the moving body (dims = the anchor point in its frame), SHAPE_PINWORLD on the fixed one (anchor = its origin)
             SHAPE_PIN = 5, SHAPE_PINWORLD = 6 };   // rimless wheel of example/rimless-wheel/coldet-plugin.cpp: dims = (R, W, N_SPOKES), params.h:4-6; collides with planes only
enum Model { MODEL_QP = 0, MODEL_AP = 1 };

struct Body {
  int shape = SHAPE_NONE;
  bool enabled = true;
  double mass = 1.0;
  double dims[3] = {0, 0, 0};
  double J[3] = {1, 1, 1};       // principal inertia, body frame
  V3 x;                          // COM position
  double quat[4] = {0, 0, 0, 1}; // x y z w
  V3 vl, va;                     // linear / angular velocity (COM, global-aligned frame)
  V3 fext, text;                 // external force / torque added by a controller each mini-step
  double R[9];                   // rotation matrix, row-major, refreshed by Sim::update_pose
};

struct ContactParams {
  double mu_c = 0, mu_v = 0, eps = 0, compliance = 0;
  int NK = 4;                    // 0 => pair disabled
};

struct Contact {
  V3 p, n, t1, t2;
  int b1, b2;                    // contact_geom1 / contact_geom2 bodies; normal points from b2 toward b1
  double dist;
  ContactParams cp;
};

struct PairDist {
  int a, b;
  double dist;
  V3 pa, pb;                     // closest points, global frame
};

struct Counters {
  long long env_steps = 0, mini_steps = 0, lcp_solves = 0, lcp_fast_calls = 0, lemke_calls = 0, pivots = 0,
            lcp_failures = 0, impact_tol_events = 0, contacts = 0, max_lcp_n = 0, pivot_flops = 0, ca_iterations = 0,
            stab_iterations = 0, stab_lcp_solves = 0, stab_line_search_failures = 0;
};

struct Sim {
  std::vector<Body> bodies;
  std::vector<ContactParams> cparams;   // [i*nb + j], i<j
  V3 gravity;
  double contact_dist_thresh = 1e-6;    // ConstraintSimulator.cpp:56
  double min_step_size;                 // TimeSteppingSimulator.cpp:48
  int model = MODEL_QP;
  double current_time = 0.0;
  LCP lcp;
  Vec zlast;                            // ImpactConstraintHandler::_zlast
  Vec vlast;                            // ImpactConstraintHandler::_v, the no-slip LCP's solution / warm start (:1239)
  Counters cnt;
  // ConstraintStabilization (ConstraintStabilization.cpp:53-66): max_iterations 0 = off, < 0 = the reference's default
  // (UINT_MAX, i.e. until no pair is closer than eps); eps = +NEAR_ZERO as coded (rule H7)
  int stab_max_iterations = 0;
  double stab_eps;
  LCP stab_lcp;                         // ConstraintStabilization::_lcp
  bool mini_failed = false;             // an LCP of the current mini-step stayed unsolved (LCPSolverException in the reference)
  // taps for parity tests: LCP of the most recent impact solve
  int last_n = 0;
  Vec last_MM, last_qq, last_z;
  std::vector<Contact> last_contacts;

  // Optional fixed-base reduced-coordinate articulated body (RCArticulatedBody, oracle_rc.h): its links are bodies
  // [rc_first, rc_first + rc.n_links) of `bodies`, link 0 (the base) a disabled body.  All moving links form one super
  // body whose generalized coordinates are the joint positions (ImpactConstraintHandler.cpp:1817-1895,1905-1916).
  bool has_rc = false;
  RCModel rc;
  int rc_first = 0;
  int rc_fdyn = 0;                      // 0: fsab (ABA), 1: crb (RCArticulatedBody.cpp:178-201)
  Vec jq, jqd, jtau;                    // joint positions, velocities, feed-forward generalized forces
  bool has_ctrl = false;                // joint-space PD law of example/ur10/controller.cpp:46-96
  Vec ctrl_kp, ctrl_kv, ctrl_amp, ctrl_freq;
  bool is_link(int b) const { return has_rc && b > rc_first && b < rc_first + rc.n_links; }
  int super_of(int b) const { return is_link(b) ? rc_first + 1 : b; }
  void rc_update_links();               // update_link_poses / update_link_velocities

  Sim();
  void init(int nb);
  void update_pose(Body& b);
  double step(double dt);                       // TimeSteppingSimulator::step
  double do_mini_step(double dt);               // TimeSteppingSimulator::do_mini_step
  // pieces, public for unit tests
  void broad_phase(std::vector<std::pair<int, int> >& pairs) const;
  void calc_pairwise_distances(const std::vector<std::pair<int, int> >& pairs, std::vector<PairDist>& out) const;
  void find_contacts(int a, int b, double TOL, std::vector<Contact>& out) const;
  void find_unilateral_constraints(const std::vector<PairDist>& pd, std::vector<Contact>& out) const;
  double calc_CA_Euler_step(const PairDist& pdi) const;
  double calc_constraint_vel(const Contact& c) const;
  void calc_fwd_dyn_and_integrate_velocity(double h);
  void process_constraints(std::vector<Contact>& contacts);
  void stabilize();                             // ConstraintStabilization::stabilize, called at the end of step()
  // assemble the LCP (MM,qq) of one island without solving (for the assembly parity test)
  void assemble_island_lcp(const std::vector<Contact*>& cons, const std::vector<int>& island_bodies, int& n, Vec& MM, Vec& qq);
};

}  // namespace oracle
