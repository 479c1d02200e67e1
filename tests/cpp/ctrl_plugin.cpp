// A controller plugin in the shape of example/ur10/controller.cpp, compiled against the facade instead of <Moby/...>:
// extern "C" init(separator, read_map, time) finds the simulator and the robot in the read map (:98-113), installs a
// constraint callback (:114) and the controller function pointer (:116), and sets the joints' starting velocities through
// get_joints() / get_coord_index() / set_generalized_velocity (:118-142).  The controller itself is the file's PD law on
// sinusoidal joint targets (:46-96), looked up by joint id.
#include <cmath>
#include <cstdio>
#include <map>
#include <b200moby.hpp>

using std::shared_ptr;
using namespace Ravelin;
using namespace Moby;

Moby::RCArticulatedBodyPtr robot;
std::shared_ptr<TimeSteppingSimulator> sim;
int g_callback_calls = 0, g_min_contacts = 1 << 30;

void check_constraint_num(std::vector<UnilateralConstraint>& constraints, std::shared_ptr<void>) {
  std::vector<UnilateralConstraint> c = sim->get_rigid_constraints();
  g_callback_calls++;
  if ((int)c.size() < g_min_contacts) g_min_contacts = (int)c.size();
  if (c.size() != constraints.size()) fprintf(stderr, "constraint lists differ\n");
}

VectorNd& controller(shared_ptr<ControlledBody>, VectorNd& u, double t, void*) {
  VectorNd q, qd;
  robot->get_generalized_coordinates_euler(q);
  robot->get_generalized_velocity(DynamicBodyd::eEuler, qd);
  const std::vector<shared_ptr<Jointd> >& joints = robot->get_joints();
  std::map<std::string, unsigned> mapping;
  for (unsigned i = 0; i < joints.size(); i++) mapping[joints[i]->joint_id] = joints[i]->get_coord_index();
  const double sh_q_des = std::sin(t * 1.0) * 0.5, sh_qd_des = std::cos(t * 1.0) * 0.5;
  const double el_q_des = std::sin(t * 2.0) * 0.3, el_qd_des = std::cos(t * 2.0) * 0.3;
  const double SH_KP = 30.0, SH_KV = 6.0, EL_KP = 10.0, EL_KV = 2.0;
  u.assign(robot->num_generalized_coordinates(DynamicBodyd::eSpatial), 0.0);
  u[mapping["shoulder"]] = SH_KP * (sh_q_des - q[mapping["shoulder"]]) + SH_KV * (sh_qd_des - qd[mapping["shoulder"]]);
  u[mapping["elbow"]] = EL_KP * (el_q_des - q[mapping["elbow"]]) + EL_KV * (el_qd_des - qd[mapping["elbow"]]);
  return u;
}

extern "C" {
void init(void*, const std::map<std::string, Moby::BasePtr>& read_map, double) {
  for (std::map<std::string, Moby::BasePtr>::const_iterator i = read_map.begin(); i != read_map.end(); i++) {
    if (!sim) sim = std::dynamic_pointer_cast<TimeSteppingSimulator>(i->second);
    if (i->first == "arm2") robot = std::dynamic_pointer_cast<RCArticulatedBody>(i->second);
  }
  if (!sim || !robot) { fprintf(stderr, "plugin: simulator or robot not in the read map\n"); return; }
  sim->constraint_callback_fn = &check_constraint_num;
  robot->controller = &controller;
  const std::vector<shared_ptr<Jointd> >& joints = robot->get_joints();
  std::map<std::string, double> qd_init;
  qd_init["shoulder"] = std::cos(0) * 0.5;
  qd_init["elbow"] = std::cos(0) * 0.3;
  VectorNd qd;
  robot->get_generalized_velocity(DynamicBodyd::eEuler, qd);
  for (unsigned i = 0; i < joints.size(); i++) qd[joints[i]->get_coord_index()] = qd_init[joints[i]->joint_id];
  robot->set_generalized_velocity(DynamicBodyd::eEuler, qd);
}
int plugin_callback_calls() { return g_callback_calls; }
int plugin_min_contacts() { return g_min_contacts; }
}
