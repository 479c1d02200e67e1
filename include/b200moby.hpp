// b200moby.hpp -- C++ host facade over the C ABI (include/b200moby.h) with Moby's class and member names for the
// time-stepping contact hot path, so a program or plugin written against Moby's TimeSteppingSimulator ports by
// recompiling against this header and linking libb200moby.so.
//
// What is mirrored (Moby tree, file:line):
//   Moby::TimeSteppingSimulator::step / current_time / min_step_size        include/Moby/TimeSteppingSimulator.h:27-52, Simulator.h:50,80
//   Moby::ConstraintSimulator members contact_dist_thresh, contact_params,
//     cstab.max_iterations, post_mini_step_callback_fn                       include/Moby/ConstraintSimulator.h:40-100
//   Moby::Simulator::add_dynamic_body, post_step_callback_fn                 include/Moby/Simulator.h:44-80
//   Moby::RigidBody (set_pose / get_pose / set_enabled / set_inertia /
//     velocity accessors / geometries / get_recurrent_forces)                include/Moby/RigidBody.h:43-80 (+ Ravelin::RigidBodyd)
//   Moby::BoxPrimitive / SpherePrimitive / PlanePrimitive (mass properties)  src/BoxPrimitive.cpp, SpherePrimitive.cpp, PlanePrimitive.cpp
//   Moby::ContactParameters                                                  include/Moby/ContactParameters.h:15-50
//   Moby::GravityForce                                                       src/GravityForce.cpp:32-68
//   Moby::LCP::lcp_lemke / lcp_fast / *_regularized                          include/Moby/LCP.h:21-27
// Ravelin's value types are replaced by the minimal stand-ins in namespace Ravelin below (define
// B200MOBY_NO_RAVELIN_STANDINS when the real Ravelin headers are on the include path).
//
// Beyond the reference: a simulator can be replicated into a batch of independent instances ("envs") that step
// together on the GPU -- replicate(n) -- with per-env state accessors; one env behaves exactly like the reference's
// single simulator.  Everything here is host-side glue: all arithmetic happens in the CUDA kernels behind the C ABI
// and every call fails loudly (std::runtime_error) when no sm_100 device is present.  There is no CPU fallback.
#ifndef B200MOBY_HPP
#define B200MOBY_HPP

#include <cmath>
#include <cstring>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b200moby.h"

#ifndef B200MOBY_NO_RAVELIN_STANDINS
namespace Ravelin {
struct Origin3d {
  double v[3];
  Origin3d(double x = 0, double y = 0, double z = 0) : v{x, y, z} {}
  double& operator[](unsigned i) { return v[i]; }
  double operator[](unsigned i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};
typedef Origin3d Vector3d;
struct Quatd {
  double x, y, z, w;
  Quatd(double x_ = 0, double y_ = 0, double z_ = 0, double w_ = 1) : x(x_), y(y_), z(z_), w(w_) {}
};
struct Pose3d {
  Quatd q;
  Origin3d x;
  Pose3d() {}
  Pose3d(const Quatd& q_, const Origin3d& x_) : q(q_), x(x_) {}
};
// spatial velocity [linear; angular] at the body's COM, global-aligned frame
struct SVelocityd {
  Vector3d linear, angular;
  const Vector3d& get_linear() const { return linear; }
  const Vector3d& get_angular() const { return angular; }
  void set_linear(const Vector3d& l) { linear = l; }
  void set_angular(const Vector3d& a) { angular = a; }
};
struct SpatialRBInertiad {
  double m = 1.0;
  double J[3] = {1, 1, 1};     // principal moments, body frame
};
typedef std::vector<double> VectorNd;
// dense column-major matrix
class MatrixNd {
 public:
  MatrixNd() : r_(0), c_(0) {}
  MatrixNd(unsigned r, unsigned c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
  void resize(unsigned r, unsigned c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0); }
  unsigned rows() const { return r_; }
  unsigned columns() const { return c_; }
  double& operator()(unsigned i, unsigned j) { return d_[(size_t)j * r_ + i]; }
  double operator()(unsigned i, unsigned j) const { return d_[(size_t)j * r_ + i]; }
  const double* data() const { return d_.data(); }
  double* data() { return d_.data(); }
 private:
  unsigned r_, c_;
  std::vector<double> d_;
};
}  // namespace Ravelin
#endif

namespace Moby {

inline void b200_check(b200moby_status st, const char* what) {
  if (st != B200MOBY_OK) throw std::runtime_error(std::string(what) + ": " + b200moby_last_error());
}

class LCPSolverException : public std::runtime_error {
 public:
  LCPSolverException() : std::runtime_error("LCP solver failed") {}
};

// ---------------------------------------------------------------------------------------------- LCP (LCP.h:21-27)
class LCP {
 public:
  LCP(int device = 0) : pivots(0), device_(device) {}
  bool lcp_lemke(const Ravelin::MatrixNd& M, const Ravelin::VectorNd& q, Ravelin::VectorNd& z, double piv_tol = -1.0, double zero_tol = -1.0) {
    return single(0, M, q, z, false, piv_tol, zero_tol, 0, 1, 0);
  }
  bool lcp_fast(const Ravelin::MatrixNd& M, const Ravelin::VectorNd& q, Ravelin::VectorNd& z, double zero_tol = -1.0) {
    return single(1, M, q, z, z.size() == q.size(), -1.0, zero_tol, 0, 1, 0);          // z doubles as the warm start (LCP.cpp:65)
  }
  bool lcp_lemke_regularized(const Ravelin::MatrixNd& M, const Ravelin::VectorNd& q, Ravelin::VectorNd& z, int min_exp = -20,
                             unsigned step_exp = 1, int max_exp = 1, double piv_tol = -1.0, double zero_tol = -1.0) {
    return single(2, M, q, z, false, piv_tol, zero_tol, min_exp, (int)step_exp, max_exp);
  }
  bool lcp_fast_regularized(const Ravelin::MatrixNd& M, const Ravelin::VectorNd& q, Ravelin::VectorNd& z, int min_exp = -20,
                            unsigned step_exp = 4, int max_exp = 20, double /*piv_tol*/ = -1.0, double zero_tol = -1.0) {
    return single(3, M, q, z, z.size() == q.size(), -1.0, zero_tol, min_exp, (int)step_exp, max_exp);
  }
  // Batched form (beyond the reference): `batch` problems of dimension n, M column-major per problem, host buffers.
  // status[b] holds the B200MOBY_LCP_* word of each problem.
  void lcp_lemke_batch(int batch, int n, const double* M, const double* q, double* z, int* status, int* pivots_out = nullptr) {
    b200_check(b200moby_lcp_lemke_host(batch, n, M, q, z, -1.0, -1.0, status, pivots_out, device_), "b200moby_lcp_lemke_host");
  }
  void lcp_fast_batch(int batch, int n, const double* M, const double* q, double* z, bool warm, int* status, int* pivots_out = nullptr) {
    b200_check(b200moby_lcp_fast_host(batch, n, M, q, z, warm ? 1 : 0, -1.0, status, pivots_out, device_), "b200moby_lcp_fast_host");
  }
  unsigned pivots;

 private:
  bool single(int mode, const Ravelin::MatrixNd& M, const Ravelin::VectorNd& q, Ravelin::VectorNd& z, bool warm, double piv_tol,
              double zero_tol, int min_exp, int step_exp, int max_exp) {
    const int n = (int)q.size();
    if ((int)M.rows() != n || (int)M.columns() != n) throw std::invalid_argument("LCP: M must be n x n");
    if (!warm) z.assign(n, 0.0);
    if (n == 0) return true;
    int status = 0, piv = 0;
    b200_check(b200moby_lcp_solve_host(mode, 1, n, M.data(), q.data(), z.data(), warm ? 1 : 0, piv_tol, zero_tol, min_exp, step_exp,
                                       max_exp, &status, &piv, device_), "b200moby_lcp_solve_host");
    pivots = (unsigned)piv;
    return status == B200MOBY_LCP_OK || status == B200MOBY_LCP_TRIVIAL || status >= B200MOBY_LCP_REGULARIZED;
  }
  int device_;
};

// ---------------------------------------------------------------------------------------------- scene objects
class Base {
 public:
  virtual ~Base() {}
  std::string id;
};
typedef std::shared_ptr<Base> BasePtr;

class Primitive : public Base {
 public:
  int shape = B200MOBY_SHAPE_NONE;
  double dims[3] = {0, 0, 0};
  void set_mass(double m) { mass_ = m; density_ = -1.0; }
  void set_density(double d) { density_ = d; mass_ = -1.0; }
  virtual double volume() const { return 0.0; }
  double get_mass() const { return mass_ >= 0.0 ? mass_ : (density_ >= 0.0 ? density_ * volume() : 0.0); }
  virtual Ravelin::SpatialRBInertiad get_inertia() const { Ravelin::SpatialRBInertiad J; J.m = get_mass(); return J; }
 protected:
  double mass_ = -1.0, density_ = -1.0;
};
typedef std::shared_ptr<Primitive> PrimitivePtr;

class BoxPrimitive : public Primitive {           // BoxPrimitive::calc_mass_properties
 public:
  BoxPrimitive(double xlen = 1, double ylen = 1, double zlen = 1) { shape = B200MOBY_SHAPE_BOX; dims[0] = xlen; dims[1] = ylen; dims[2] = zlen; }
  double volume() const override { return dims[0] * dims[1] * dims[2]; }
  Ravelin::SpatialRBInertiad get_inertia() const override {
    Ravelin::SpatialRBInertiad J; J.m = get_mass();
    J.J[0] = J.m * (dims[1] * dims[1] + dims[2] * dims[2]) / 12.0;
    J.J[1] = J.m * (dims[0] * dims[0] + dims[2] * dims[2]) / 12.0;
    J.J[2] = J.m * (dims[0] * dims[0] + dims[1] * dims[1]) / 12.0;
    return J;
  }
};
class SpherePrimitive : public Primitive {        // SpherePrimitive::calc_mass_properties
 public:
  explicit SpherePrimitive(double radius = 1) { shape = B200MOBY_SHAPE_SPHERE; dims[0] = radius; }
  double volume() const override { return (4.0 / 3.0) * M_PI * dims[0] * dims[0] * dims[0]; }
  Ravelin::SpatialRBInertiad get_inertia() const override {
    Ravelin::SpatialRBInertiad J; J.m = get_mass();
    J.J[0] = J.J[1] = J.J[2] = 0.4 * J.m * dims[0] * dims[0];
    return J;
  }
};
class PlanePrimitive : public Primitive {         // half-space y <= 0 of the body frame (PlanePrimitive.cpp:342-411)
 public:
  PlanePrimitive() { shape = B200MOBY_SHAPE_PLANE; }
};
// The rimless wheel of example/rimless-wheel: in the reference a CollisionGeometry WITHOUT a primitive whose distance and
// contacts come from the collision-detection plugin (coldet-plugin.cpp:86-137,222-310; params.h:4-6); here a shape of its own.
// example/contact-constrained-pendulum: the pin joint its collision-detection plugin emulates with six frictionless contacts
// (contact-constrained-pendulum-coldet-plugin.cpp:60-110).  PinPrimitive goes on the moving body (anchor point in its frame),
// PinWorldPrimitive on the fixed one (anchor = its origin).
class PinPrimitive : public Primitive {
 public:
  PinPrimitive(double ax = 0.0, double ay = 1.0, double az = 0.0) { shape = B200MOBY_SHAPE_PIN; dims[0] = ax; dims[1] = ay; dims[2] = az; }
};
class PinWorldPrimitive : public Primitive {
 public:
  PinWorldPrimitive() { shape = B200MOBY_SHAPE_PINWORLD; }
};
class RimlessWheelPrimitive : public Primitive {
 public:
  RimlessWheelPrimitive(double R = 1.0, double W = 0.0, unsigned n_spokes = 6) { shape = B200MOBY_SHAPE_WHEEL; dims[0] = R; dims[1] = W; dims[2] = (double)n_spokes; }
};

class CollisionGeometry : public Base {
 public:
  void set_geometry(PrimitivePtr p) { primitive_ = p; }
  PrimitivePtr get_geometry() const { return primitive_; }
 private:
  PrimitivePtr primitive_;
};
typedef std::shared_ptr<CollisionGeometry> CollisionGeometryPtr;

class RecurrentForce : public Base {};
typedef std::shared_ptr<RecurrentForce> RecurrentForcePtr;
class GravityForce : public RecurrentForce {
 public:
  Ravelin::Vector3d gravity;
};

class TimeSteppingSimulator;

// ControlledBody (ControlledBody.h:37-40): the controller callback and its argument
class ControlledBody : public Base, public std::enable_shared_from_this<ControlledBody> {
 public:
  Ravelin::VectorNd& (*controller)(std::shared_ptr<ControlledBody>, Ravelin::VectorNd&, double, void*) = nullptr;
  void* controller_arg = nullptr;
};
typedef std::shared_ptr<ControlledBody> ControlledBodyPtr;

class RigidBody : public ControlledBody {
 public:
  std::string body_id;
  std::list<CollisionGeometryPtr> geometries;
  void set_enabled(bool e) { enabled_ = e; }
  bool is_enabled() const { return enabled_; }
  void set_inertia(const Ravelin::SpatialRBInertiad& J) { J_ = J; }
  const Ravelin::SpatialRBInertiad& get_inertia() const { return J_; }
  std::list<RecurrentForcePtr>& get_recurrent_forces() { return forces_; }
  // pose / velocity of env 0 (of env e with the second argument); reads go to the device once the simulator runs
  void set_pose(const Ravelin::Pose3d& p, int env = -1);
  Ravelin::Pose3d get_pose(int env = 0) const;
  void set_velocity(const Ravelin::SVelocityd& v, int env = -1);
  Ravelin::SVelocityd get_velocity(int env = 0) const;

 private:
  friend class TimeSteppingSimulator;
  friend class RCArticulatedBody;
  bool enabled_ = true;
  Ravelin::SpatialRBInertiad J_;
  std::list<RecurrentForcePtr> forces_;
  Ravelin::Pose3d pose0_;
  Ravelin::SVelocityd vel0_;
  TimeSteppingSimulator* sim_ = nullptr;
  int index_ = -1;
};
typedef std::shared_ptr<RigidBody> RigidBodyPtr;

// ---------------------------------------------------------------------------------------------- joints, articulated body
struct DynamicBodyd { enum GeneralizedCoordinateType { eEuler, eSpatial }; };
// Joint.h / Ravelin::Jointd: one-DoF joints of a fixed-base tree.  set_location(point, inboard, outboard) and set_axis(axis)
// take GLOBAL-frame values at the links' construction poses, as example/sims-in-code/pendulum.cpp:85-103 uses them
// (set_location before set_axis).
class Joint : public Base {
 public:
  std::string joint_id;
  Ravelin::VectorNd q = Ravelin::VectorNd(1, 0.0), qd = Ravelin::VectorNd(1, 0.0);    // position / velocity of the CURRENT env (see RCArticulatedBody)
  virtual int type() const = 0;
  unsigned num_dof() const { return 1; }
  unsigned get_coord_index() const { return (unsigned)coord_; }
  void set_location(const Ravelin::Vector3d& p, RigidBodyPtr inboard, RigidBodyPtr outboard) { loc_ = p; in_ = inboard; out_ = outboard; }
  void set_axis(const Ravelin::Vector3d& a) { axis_ = a; }
  RigidBodyPtr get_inboard_link() const { return in_; }
  RigidBodyPtr get_outboard_link() const { return out_; }

 private:
  friend class RCArticulatedBody;
  friend class TimeSteppingSimulator;
  Ravelin::Vector3d loc_, axis_ = Ravelin::Vector3d(0, 0, 1);
  RigidBodyPtr in_, out_;
  int coord_ = -1;
};
class RevoluteJoint : public Joint { public: int type() const override { return B200MOBY_JOINT_REVOLUTE; } };
class PrismaticJoint : public Joint { public: int type() const override { return B200MOBY_JOINT_PRISMATIC; } };
typedef std::shared_ptr<Joint> JointPtr;

// RCArticulatedBody (RCArticulatedBody.h:43; dynamics in Ravelin's RCArticulatedBodyd): fixed-base tree of one-DoF joints.
// Generalized coordinates are the joint positions in joint order (get_coord_index); with several envs the accessors
// refer to the env selected by TimeSteppingSimulator (env 0 outside controller callbacks).
class RCArticulatedBody : public ControlledBody {
 public:
  enum ForwardDynamicsAlgorithmType { eFeatherstone, eCRB };
  ForwardDynamicsAlgorithmType algorithm_type = eCRB;                      // RCArticulatedBody.cpp:60 default
  void set_links_and_joints(const std::vector<RigidBodyPtr>& links, const std::vector<JointPtr>& joints);
  const std::vector<RigidBodyPtr>& get_links() const { return links_; }
  const std::vector<JointPtr>& get_joints() const { return joints_; }
  void set_floating_base(bool f) { if (f) throw std::runtime_error("floating bases are outside the accelerated path (SURVEY.md 8f #4)"); }
  bool is_floating_base() const { return false; }
  std::list<RecurrentForcePtr>& get_recurrent_forces() { return forces_; }
  unsigned num_generalized_coordinates(DynamicBodyd::GeneralizedCoordinateType) const { return (unsigned)joints_.size(); }
  void get_generalized_coordinates_euler(Ravelin::VectorNd& q);
  void get_generalized_velocity(DynamicBodyd::GeneralizedCoordinateType, Ravelin::VectorNd& qd);
  void set_generalized_coordinates_euler(const Ravelin::VectorNd& q);
  void set_generalized_velocity(DynamicBodyd::GeneralizedCoordinateType, const Ravelin::VectorNd& qd);

 private:
  friend class TimeSteppingSimulator;
  std::vector<RigidBodyPtr> links_;        // link 0 = the base; parent before child
  std::vector<JointPtr> joints_;           // joint k = inboard joint of link k + 1
  std::vector<int> parent_;
  std::list<RecurrentForcePtr> forces_;
  TimeSteppingSimulator* sim_ = nullptr;
  int first_body_ = -1;
};
typedef std::shared_ptr<RCArticulatedBody> RCArticulatedBodyPtr;

// UnilateralConstraint (UnilateralConstraint.h): the contact fields the constraint callbacks and get_rigid_constraints() expose
struct UnilateralConstraint {
  enum UnilateralConstraintType { eNone, eContact, eLimit };
  UnilateralConstraintType constraint_type = eContact;
  Ravelin::Vector3d contact_point, contact_normal, contact_tan1, contact_tan2;
  CollisionGeometryPtr contact_geom1, contact_geom2;
  double signed_violation = 0.0;
};

class ContactParameters : public Base {           // defaults: ContactParameters.cpp:21-28
 public:
  ContactParameters() {}
  ContactParameters(BasePtr o1, BasePtr o2) : objects(o1, o2) {}
  std::pair<BasePtr, BasePtr> objects;
  double epsilon = 0.0, mu_coulomb = 0.0, mu_viscous = 0.0, compliance = 0.0;
  unsigned NK = 4;
};

struct ConstraintStabilization {                   // ConstraintStabilization.h:27-37: the members the path honours
  unsigned max_iterations = 0xffffffffu;           // ConstraintStabilization.cpp:56: no limit by default; 0 switches stabilization off (ur10.xml:12)
  double eps = 1.4901161193847656e-08;             // :59 (+sqrt(eps), rule H7); other values are refused at compile()
};

// ---------------------------------------------------------------------------------------------- the simulator
class TimeSteppingSimulator : public Base {
 public:
  TimeSteppingSimulator() {}
  ~TimeSteppingSimulator() { if (h_) b200moby_destroy(h_); }
  TimeSteppingSimulator(const TimeSteppingSimulator&) = delete;
  TimeSteppingSimulator& operator=(const TimeSteppingSimulator&) = delete;

  // --- Moby's members ---
  double current_time = 0.0;                                     // Simulator::current_time (env 0)
  double min_step_size = 1.4901161193847656e-08;                 // TimeSteppingSimulator.cpp:48
  double contact_dist_thresh = 1e-6;                             // ConstraintSimulator.cpp:56
  ConstraintStabilization cstab;
  std::map<std::pair<BasePtr, BasePtr>, std::shared_ptr<ContactParameters> > contact_params;
  void (*post_step_callback_fn)(TimeSteppingSimulator*) = nullptr;           // Simulator.h:80
  // ConstraintSimulator.h:51-68.  Both are called once per step() with the contacts of the state the step ended in (found
  // again on the device and copied to the host: a slow path); the list is informational -- edits do not feed back into the
  // solve, which has already run on the device.
  void (*constraint_callback_fn)(std::vector<UnilateralConstraint>&, std::shared_ptr<void>) = nullptr;
  void (*constraint_post_callback_fn)(const std::vector<UnilateralConstraint>&, std::shared_ptr<void>) = nullptr;
  std::shared_ptr<void> constraint_callback_data, constraint_post_callback_data;
  void (*post_mini_step_callback_fn)(TimeSteppingSimulator*) = nullptr;      // ConstraintSimulator.h:55; see step()
  int impact_model = B200MOBY_MODEL_QP;                          // default build; B200MOBY_MODEL_AP == -DUSE_AP_MODEL
  int device = 0;

  void add_dynamic_body(ControlledBodyPtr body) {
    if (h_) throw std::logic_error("add_dynamic_body after the first step");
    dyn_bodies_.push_back(body);
    if (RigidBodyPtr rb = std::dynamic_pointer_cast<RigidBody>(body)) { rb->sim_ = this; rb->index_ = (int)bodies_.size(); bodies_.push_back(rb); return; }
    RCArticulatedBodyPtr ab = std::dynamic_pointer_cast<RCArticulatedBody>(body);
    if (!ab) throw std::invalid_argument("add_dynamic_body: RigidBody or RCArticulatedBody");
    if (rc_) throw std::runtime_error("one RCArticulatedBody per simulator on the accelerated path");
    if (ab->links_.empty()) throw std::logic_error("set_links_and_joints first");
    rc_ = ab; ab->sim_ = this; ab->first_body_ = (int)bodies_.size();
    for (const RigidBodyPtr& l : ab->links_) { l->sim_ = this; l->index_ = (int)bodies_.size(); bodies_.push_back(l); }
  }
  const std::vector<ControlledBodyPtr>& get_dynamic_bodies() const { return dyn_bodies_; }
  // ConstraintSimulator::get_rigid_constraints: the contacts of env `env` at the current state
  std::vector<UnilateralConstraint> get_rigid_constraints(int env = 0) {
    if (!h_) compile();
    const int cap = 64, ne = n_envs_, nb = (int)bodies_.size();
    std::vector<int> count(ne), pair((size_t)cap * ne);
    std::vector<double> pt((size_t)cap * 3 * ne), nr(pt.size()), t1(pt.size()), t2(pt.size()), dist((size_t)cap * ne);
    b200_check(b200moby_find_contacts_host(h_, cap, count.data(), pt.data(), nr.data(), t1.data(), t2.data(), pair.data(), dist.data()), "b200moby_find_contacts_host");
    std::vector<UnilateralConstraint> out;
    for (int i = 0; i < count[env] && i < cap; i++) {
      UnilateralConstraint c;
      for (int k = 0; k < 3; k++) {
        const size_t o = ((size_t)i * 3 + k) * ne + env;
        c.contact_point[k] = pt[o]; c.contact_normal[k] = nr[o]; c.contact_tan1[k] = t1[o]; c.contact_tan2[k] = t2[o];
      }
      const int p = pair[(size_t)i * ne + env], b1 = p / nb, b2 = p % nb;
      if (!bodies_[b1]->geometries.empty()) c.contact_geom1 = bodies_[b1]->geometries.front();
      if (!bodies_[b2]->geometries.empty()) c.contact_geom2 = bodies_[b2]->geometries.front();
      c.signed_violation = dist[(size_t)i * ne + env];
      out.push_back(c);
    }
    return out;
  }
  void add_contact_parameters(std::shared_ptr<ContactParameters> cp) { contact_params[cp->objects] = cp; }
  std::vector<std::pair<BasePtr, BasePtr> > disabled_pairs;       // <DisabledPair> (ConstraintSimulator.cpp:585-611): never checked

  // --- batch extension: n independent copies of the scene; per-env perturbations through RigidBody::set_pose(p, env) ---
  void replicate(int n_envs) {
    if (h_) throw std::logic_error("replicate after the first step");
    if (n_envs < 1) throw std::invalid_argument("replicate: n_envs >= 1");
    n_envs_ = n_envs;
  }
  int num_envs() const { return n_envs_; }

  // TimeSteppingSimulator::step (TimeSteppingSimulator.cpp:52-111): every env advances by dt; returns dt.
  // post_mini_step_callback_fn is invoked once per step() (after the device finished the step's mini-steps): per
  // mini-step host callbacks would serialise the batch, so the reference's per-mini-step granularity is not kept.
  double step(double dt) {
    if (!h_) compile();
    run_controllers(dt);
    b200_check(b200moby_step(h_, dt, 1, nullptr), "b200moby_step");
    dirty_ = true; jdirty_ = true;
    if (constraint_callback_fn || constraint_post_callback_fn) {
      std::vector<UnilateralConstraint> c = get_rigid_constraints(0);
      if (constraint_callback_fn) constraint_callback_fn(c, constraint_callback_data);
      if (constraint_post_callback_fn) constraint_post_callback_fn(c, constraint_post_callback_data);
    }
    if (post_mini_step_callback_fn) post_mini_step_callback_fn(this);
    current_time += dt;                                           // every env advances by exactly dt per step()
    if (post_step_callback_fn) post_step_callback_fn(this);
    return dt;
  }
  // n steps without returning to the host in between (no callbacks)
  void step_n(double dt, int n) {
    if (!h_) compile();
    if (rc_ && rc_->controller) throw std::logic_error("step_n: a controller callback needs the host every step; use step()");
    b200_check(b200moby_step(h_, dt, n, nullptr), "b200moby_step");
    dirty_ = true; jdirty_ = true;
    current_time += dt * n;
  }
  b200moby_counters counters() {
    if (!h_) compile();
    b200moby_counters c;
    b200_check(b200moby_get_counters(h_, &c), "b200moby_get_counters");
    return c;
  }
  b200moby_handle handle() { if (!h_) compile(); return h_; }

  // host copies of the state, SoA [body][7|6][env] (regress.cpp:78-95 row order per body: x y z qx qy qz qw)
  const std::vector<double>& q() { sync_host(); return q_; }
  const std::vector<double>& v() { sync_host(); return v_; }

 private:
  friend class RigidBody;
  void ensure_host() {
    const size_t nb = bodies_.size(), ne = (size_t)n_envs_;
    if (q_.size() == nb * 7 * ne) return;
    q_.assign(nb * 7 * ne, 0.0); v_.assign(nb * 6 * ne, 0.0);
    for (size_t b = 0; b < nb; b++)
      for (size_t e = 0; e < ne; e++) write_body(b, e, bodies_[b]->pose0_, bodies_[b]->vel0_);
  }
  void write_body(size_t b, size_t e, const Ravelin::Pose3d& p, const Ravelin::SVelocityd& vel) {
    const size_t ne = (size_t)n_envs_;
    double* q = &q_[(b * 7) * ne + e];
    q[0] = p.x[0]; q[ne] = p.x[1]; q[2 * ne] = p.x[2]; q[3 * ne] = p.q.x; q[4 * ne] = p.q.y; q[5 * ne] = p.q.z; q[6 * ne] = p.q.w;
    double* v = &v_[(b * 6) * ne + e];
    for (int k = 0; k < 3; k++) { v[k * ne] = vel.linear[k]; v[(3 + k) * ne] = vel.angular[k]; }
  }
  void sync_host() {
    ensure_host();
    if (h_ && dirty_) { b200_check(b200moby_get_state(h_, q_.data(), v_.data()), "b200moby_get_state"); dirty_ = false; }
  }
  void push_state() {
    if (h_) { b200_check(b200moby_set_state(h_, q_.data(), v_.data()), "b200moby_set_state"); dirty_ = true; }
  }
  // joint state of the articulated body, host copy [dof][env]
  friend class RCArticulatedBody;
  int ndof() const { return rc_ ? (int)rc_->joints_.size() : 0; }
  void sync_joints() {
    if (!rc_) return;
    const size_t n = (size_t)ndof() * n_envs_;
    if (jq_.size() != n) { jq_.assign(n, 0.0); jqd_.assign(n, 0.0); for (int k = 0; k < ndof(); k++) for (int e = 0; e < n_envs_; e++) { jq_[(size_t)k * n_envs_ + e] = rc_->joints_[k]->q[0]; jqd_[(size_t)k * n_envs_ + e] = rc_->joints_[k]->qd[0]; } }
    if (h_ && jdirty_) { b200_check(b200moby_get_joint_state(h_, jq_.data(), jqd_.data()), "b200moby_get_joint_state"); jdirty_ = false; }
    for (int k = 0; k < ndof(); k++) {
      const double qk = jq_[(size_t)k * n_envs_ + cur_env_], qdk = jqd_[(size_t)k * n_envs_ + cur_env_];
      rc_->joints_[k]->q[0] = qk + ctrl_dt_ * qdk; rc_->joints_[k]->qd[0] = qdk;
    }
  }
  void push_joints() {
    if (h_) { b200_check(b200moby_set_joint_state(h_, jq_.data(), jqd_.data()), "b200moby_set_joint_state"); dirty_ = true; }
  }
  // Simulator.cpp:339-348: the controller of every body, at current_time, before the forward dynamics.  One host call per
  // env (the accessors of the body refer to that env during the call); its generalized forces go to the device.  The
  // reference calls it inside the mini-step AFTER the position half of the semi-implicit Euler step
  // (TimeSteppingSimulator.cpp:155-192): it sees q(t) + h qd(t) and qd(t).  The callback here runs before the device step,
  // so the joint positions it reads are advanced by dt * qd on the host -- exact whenever the step is one mini-step.
  void run_controllers(double dt) {
    for (const RigidBodyPtr& b : bodies_) if (b->controller) throw std::runtime_error("controllers on free rigid bodies are outside the accelerated path");
    if (!rc_ || !rc_->controller) return;
    const int nd = ndof();
    std::vector<double> tau((size_t)nd * n_envs_, 0.0);
    Ravelin::VectorNd u;
    ctrl_dt_ = dt;
    for (int e = 0; e < n_envs_; e++) {
      cur_env_ = e;
      sync_joints();
      u.assign(nd, 0.0);
      Ravelin::VectorNd& r = rc_->controller(rc_, u, current_time, rc_->controller_arg);
      for (int k = 0; k < nd && k < (int)r.size(); k++) tau[(size_t)k * n_envs_ + e] = r[k];
    }
    cur_env_ = 0; ctrl_dt_ = 0.0;
    sync_joints();
    b200_check(b200moby_set_joint_forces(h_, tau.data()), "b200moby_set_joint_forces");
  }
  // object graph -> b200moby_scene_desc (what XMLReader::read + the simulator's containers hold in the reference)
  void compile() {
    if (cstab.eps != 1.4901161193847656e-08) throw std::runtime_error("cstab.eps: only the default sqrt(eps) (ConstraintStabilization.cpp:59) is on the accelerated path");
    const int nb = (int)bodies_.size(), ne = n_envs_;
    if (nb == 0) throw std::logic_error("no bodies");
    ensure_host();
    std::vector<int> shape((size_t)nb * ne), enabled((size_t)nb * ne), NK((size_t)nb * nb * ne, 0);
    std::vector<double> mass((size_t)nb * ne), dims((size_t)nb * 3 * ne), inertia((size_t)nb * 3 * ne);
    std::vector<double> mu_c((size_t)nb * nb * ne, 0.0), mu_v(mu_c), eps(mu_c), comp(mu_c);
    b200moby_scene_desc d; memset(&d, 0, sizeof(d));
    d.gravity[0] = d.gravity[1] = d.gravity[2] = 0.0;
    if (rc_) for (const RecurrentForcePtr& f : rc_->forces_)
      if (const GravityForce* g = dynamic_cast<const GravityForce*>(f.get())) for (int k = 0; k < 3; k++) d.gravity[k] = g->gravity[k];
    for (int b = 0; b < nb; b++) {
      const RigidBody& rb = *bodies_[b];
      PrimitivePtr prim;
      if (rb.geometries.size() > 1) throw std::runtime_error("one collision geometry per body on the accelerated path");
      if (!rb.geometries.empty()) prim = rb.geometries.front()->get_geometry();
      for (int e = 0; e < ne; e++) {
        shape[(size_t)b * ne + e] = prim ? prim->shape : B200MOBY_SHAPE_NONE;
        enabled[(size_t)b * ne + e] = rb.enabled_ ? 1 : 0;
        mass[(size_t)b * ne + e] = rb.J_.m;
        for (int k = 0; k < 3; k++) { dims[((size_t)b * 3 + k) * ne + e] = prim ? prim->dims[k] : 0.0; inertia[((size_t)b * 3 + k) * ne + e] = rb.J_.J[k]; }
      }
      for (const RecurrentForcePtr& f : rb.forces_)
        if (const GravityForce* g = dynamic_cast<const GravityForce*>(f.get())) for (int k = 0; k < 3; k++) d.gravity[k] = g->gravity[k];
    }
    // every pair is checked with default parameters (ContactParameters.cpp:21-28) unless contact_params overrides it
    for (int i = 0; i < nb; i++)
      for (int j = i + 1; j < nb; j++) {
        ContactParameters cp;
        for (const auto& kv : contact_params) {
          const Base* a = kv.first.first.get(); const Base* b = kv.first.second.get();
          if ((a == bodies_[i].get() && b == bodies_[j].get()) || (a == bodies_[j].get() && b == bodies_[i].get())) cp = *kv.second;
        }
        if (cp.NK < 4) cp.NK = 4;                                   // ContactParameters.cpp:132-136
        for (const auto& dp : disabled_pairs)
          if ((dp.first.get() == bodies_[i].get() && dp.second.get() == bodies_[j].get()) || (dp.first.get() == bodies_[j].get() && dp.second.get() == bodies_[i].get())) cp.NK = 0;
        for (int e = 0; e < ne; e++) {
          const size_t o = ((size_t)i * nb + j) * ne + e;
          mu_c[o] = cp.mu_coulomb; mu_v[o] = cp.mu_viscous; eps[o] = cp.epsilon; comp[o] = cp.compliance; NK[o] = (int)cp.NK;
        }
      }
    d.n_envs = ne; d.n_bodies = nb;
    d.shape = shape.data(); d.enabled = enabled.data(); d.mass = mass.data(); d.dims = dims.data(); d.inertia = inertia.data();
    d.mu_coulomb = mu_c.data(); d.mu_viscous = mu_v.data(); d.epsilon = eps.data(); d.compliance = comp.data(); d.NK = NK.data();
    d.contact_dist_thresh = contact_dist_thresh; d.min_step_size = min_step_size; d.min_step_size_env = nullptr;
    d.impact_model = impact_model;
    d.stabilization_max_iterations = cstab.max_iterations > 0x7fffffffu ? -1 : (int)cstab.max_iterations;   // UINT_MAX (the default) = no limit
    // the articulated body: tree and joint frames from the links' construction poses and the joints' global location / axis
    b200moby_rc_desc rd; memset(&rd, 0, sizeof(rd));
    std::vector<int> parent, jtype; std::vector<double> axis, locp, locc, relq;
    if (rc_) {
      const int nl = (int)rc_->links_.size();
      parent.assign(nl, 0); jtype.assign(nl, B200MOBY_JOINT_REVOLUTE); axis.assign(3 * nl, 0.0); locp.assign(3 * nl, 0.0); locc.assign(3 * nl, 0.0); relq.assign(4 * nl, 0.0);
      relq[3] = 1.0; axis[2] = 1.0;
      for (int i = 1; i < nl; i++) {
        const Joint& J = *rc_->joints_[i - 1];
        const Ravelin::Pose3d& Pi = rc_->links_[rc_->parent_[i]]->pose0_; const Ravelin::Pose3d& Po = rc_->links_[i]->pose0_;
        parent[i] = rc_->parent_[i]; jtype[i] = J.type();
        double Ri[9], Ro[9]; quat_to_R(Pi.q, Ri); quat_to_R(Po.q, Ro);
        double an = std::sqrt(J.axis_[0] * J.axis_[0] + J.axis_[1] * J.axis_[1] + J.axis_[2] * J.axis_[2]);
        if (!(an > 0.0)) throw std::runtime_error("joint '" + J.id + "': zero axis");
        for (int r = 0; r < 3; r++) {
          double a = 0, lp = 0, lc = 0;
          for (int k = 0; k < 3; k++) { a += Ro[k * 3 + r] * J.axis_[k] / an; lp += Ri[k * 3 + r] * (J.loc_[k] - Pi.x[k]); lc += Ro[k * 3 + r] * (J.loc_[k] - Po.x[k]); }
          axis[3 * i + r] = a; locp[3 * i + r] = lp; locc[3 * i + r] = lc;                 // R^T (.)
        }
        // conj(q_in) * q_out
        const Ravelin::Quatd a(-Pi.q.x, -Pi.q.y, -Pi.q.z, Pi.q.w), b = Po.q;
        relq[4 * i] = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y; relq[4 * i + 1] = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
        relq[4 * i + 2] = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w; relq[4 * i + 3] = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
      }
      rd.n_links = nl; rd.first_body = rc_->first_body_;
      rd.parent = parent.data(); rd.joint_type = jtype.data(); rd.joint_axis = axis.data(); rd.loc_parent = locp.data(); rd.loc_child = locc.data(); rd.rel_quat = relq.data();
      rd.fdyn_algorithm = rc_->algorithm_type == RCArticulatedBody::eCRB ? B200MOBY_FDYN_CRB : B200MOBY_FDYN_FSAB;
      d.rc = &rd;
      for (int e = 0; e < ne; e++) enabled[(size_t)rc_->first_body_ * ne + e] = 0;           // the base is welded to the world
    }
    b200_check(b200moby_create(&d, device, &h_), "b200moby_create");
    push_state();
    if (rc_) { jdirty_ = false; sync_joints(); push_joints(); }
  }
  static void quat_to_R(const Ravelin::Quatd& q, double* R) {      // row-major
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), x = q.x / n, y = q.y / n, z = q.z / n, w = q.w / n;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
  }

  std::vector<RigidBodyPtr> bodies_;            // every single body in scene order (the articulated body's links included)
  std::vector<ControlledBodyPtr> dyn_bodies_;   // what add_dynamic_body received
  RCArticulatedBodyPtr rc_;
  std::vector<double> jq_, jqd_;
  bool jdirty_ = false;
  int cur_env_ = 0;
  double ctrl_dt_ = 0.0;
  int n_envs_ = 1;
  b200moby_handle h_ = nullptr;
  std::vector<double> q_, v_;
  bool dirty_ = false;
};

inline void RigidBody::set_pose(const Ravelin::Pose3d& p, int env) {
  if (!sim_) { pose0_ = p; return; }
  if (env < 0 && !sim_->h_ && sim_->q_.empty()) { pose0_ = p; return; }
  sim_->sync_host();
  Ravelin::SVelocityd v = get_velocity(env < 0 ? 0 : env);
  if (env < 0) { for (int e = 0; e < sim_->n_envs_; e++) sim_->write_body(index_, e, p, get_velocity(e)); pose0_ = p; }
  else sim_->write_body(index_, env, p, v);
  sim_->push_state();
}
inline Ravelin::Pose3d RigidBody::get_pose(int env) const {
  if (!sim_) return pose0_;
  sim_->sync_host();
  const size_t ne = (size_t)sim_->n_envs_;
  const double* q = &sim_->q_[((size_t)index_ * 7) * ne + env];
  return Ravelin::Pose3d(Ravelin::Quatd(q[3 * ne], q[4 * ne], q[5 * ne], q[6 * ne]), Ravelin::Origin3d(q[0], q[ne], q[2 * ne]));
}
inline void RigidBody::set_velocity(const Ravelin::SVelocityd& vel, int env) {
  if (!sim_) { vel0_ = vel; return; }
  if (env < 0 && !sim_->h_ && sim_->q_.empty()) { vel0_ = vel; return; }
  sim_->sync_host();
  if (env < 0) { for (int e = 0; e < sim_->n_envs_; e++) sim_->write_body(index_, e, get_pose(e), vel); vel0_ = vel; }
  else sim_->write_body(index_, env, get_pose(env), vel);
  sim_->push_state();
}
inline Ravelin::SVelocityd RigidBody::get_velocity(int env) const {
  if (!sim_) return vel0_;
  sim_->sync_host();
  const size_t ne = (size_t)sim_->n_envs_;
  const double* v = &sim_->v_[((size_t)index_ * 6) * ne + env];
  Ravelin::SVelocityd out;
  for (int k = 0; k < 3; k++) { out.linear[k] = v[k * ne]; out.angular[k] = v[(3 + k) * ne]; }
  return out;
}

inline void RCArticulatedBody::set_links_and_joints(const std::vector<RigidBodyPtr>& links, const std::vector<JointPtr>& joints) {
  if (links.size() != joints.size() + 1) throw std::invalid_argument("set_links_and_joints: a tree has one joint per non-base link");
  // base = the link that is no joint's outboard; then breadth-first so that parents precede children
  RigidBodyPtr base;
  for (const RigidBodyPtr& l : links) { bool out = false; for (const JointPtr& j : joints) out = out || j->out_ == l; if (!out) { if (base) throw std::invalid_argument("two roots"); base = l; } }
  if (!base) throw std::invalid_argument("no base link (kinematic loop?)");
  links_.assign(1, base); joints_.clear(); parent_.assign(1, 0);
  for (size_t i = 0; i < links_.size(); i++)
    for (const JointPtr& j : joints)
      if (j->in_ == links_[i]) { j->coord_ = (int)joints_.size(); joints_.push_back(j); links_.push_back(j->out_); parent_.push_back((int)i); }
  if (links_.size() != links.size()) throw std::invalid_argument("set_links_and_joints: links not connected to the base");
  for (const JointPtr& j : joints_) if (j->joint_id.empty()) j->joint_id = j->id;
  for (const RigidBodyPtr& l : links_) if (l->body_id.empty()) l->body_id = l->id;
}
inline void RCArticulatedBody::get_generalized_coordinates_euler(Ravelin::VectorNd& q) {
  if (sim_) sim_->sync_joints();
  q.resize(joints_.size());
  for (size_t k = 0; k < joints_.size(); k++) q[k] = joints_[k]->q[0];
}
inline void RCArticulatedBody::get_generalized_velocity(DynamicBodyd::GeneralizedCoordinateType, Ravelin::VectorNd& qd) {
  if (sim_) sim_->sync_joints();
  qd.resize(joints_.size());
  for (size_t k = 0; k < joints_.size(); k++) qd[k] = joints_[k]->qd[0];
}
inline void RCArticulatedBody::set_generalized_coordinates_euler(const Ravelin::VectorNd& q) {
  if (q.size() != joints_.size()) throw std::invalid_argument("set_generalized_coordinates_euler: one value per joint (fixed base)");
  if (sim_) sim_->sync_joints();
  for (size_t k = 0; k < joints_.size(); k++) joints_[k]->q[0] = q[k];
  if (sim_ && !sim_->jq_.empty()) {                      // before the first step: every env starts from it; afterwards: the current env
    for (size_t k = 0; k < joints_.size(); k++)
      for (int e = 0; e < sim_->n_envs_; e++) if (!sim_->h_ || e == sim_->cur_env_) sim_->jq_[k * sim_->n_envs_ + e] = q[k];
    sim_->push_joints();
  }
}
inline void RCArticulatedBody::set_generalized_velocity(DynamicBodyd::GeneralizedCoordinateType, const Ravelin::VectorNd& qd) {
  if (qd.size() != joints_.size()) throw std::invalid_argument("set_generalized_velocity: one value per joint (fixed base)");
  if (sim_) sim_->sync_joints();
  for (size_t k = 0; k < joints_.size(); k++) joints_[k]->qd[0] = qd[k];
  if (sim_ && !sim_->jqd_.empty()) {
    for (size_t k = 0; k < joints_.size(); k++)
      for (int e = 0; e < sim_->n_envs_; e++) if (!sim_->h_ || e == sim_->cur_env_) sim_->jqd_[k * sim_->n_envs_ + e] = qd[k];
    sim_->push_joints();
  }
}

typedef TimeSteppingSimulator Simulator;      // example/sims-in-code/pendulum.cpp steps a plain Simulator: the same path without collision geometry

}  // namespace Moby

#ifndef B200MOBY_NO_RAVELIN_STANDINS
namespace Ravelin {
typedef Moby::Joint Jointd;
typedef Moby::RigidBody RigidBodyd;
typedef Moby::DynamicBodyd DynamicBodyd;
}
#endif

#endif  // B200MOBY_HPP
