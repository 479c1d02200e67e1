// example/sims-in-code/pendulum.cpp against the facade: the same construction calls (links, a RevoluteJoint placed with
// set_location / set_axis in the global frame, set_links_and_joints, set_floating_base(false), add_dynamic_body,
// sim->step(0.001)) with the two things the accelerated path does not have taken out: the OSG viewer and the
// CylinderPrimitive (the arm's inertia is set directly instead: a r = 0.025, h = 1, m = 1 cylinder along y).
// usage: pendulum <steps> <print_every> [n_envs]; prints "t arm_x arm_y arm_z qx qy qz qw  q qd" of env 0.
#include <cstdio>
#include <cstdlib>
#include <b200moby.hpp>

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 1000, every = argc > 2 ? atoi(argv[2]) : 100, ne = argc > 3 ? atoi(argv[3]) : 1;
  try {
    std::shared_ptr<Moby::Simulator> sim(new Moby::Simulator());
    std::shared_ptr<Moby::GravityForce> g(new Moby::GravityForce());
    g->gravity = Ravelin::Vector3d(0, 0, -9.8);

    Moby::RCArticulatedBodyPtr pendulum(new Moby::RCArticulatedBody());
    pendulum->id = "pendulum";
    pendulum->algorithm_type = Moby::RCArticulatedBody::eCRB;
    std::vector<Moby::RigidBodyPtr> links;
    std::vector<Moby::JointPtr> joints;

    Moby::RigidBodyPtr base(new Moby::RigidBody());
    {
      Moby::PrimitivePtr box(new Moby::BoxPrimitive(0.1, 0.1, 0.1));
      box->set_mass(1);
      base->id = "base";
      base->set_inertia(box->get_inertia());
      base->set_enabled(false);
      base->set_pose(Ravelin::Pose3d(Ravelin::Quatd(0, 0, 0, 1), Ravelin::Origin3d(0, 0, 0)));
      links.push_back(base);
    }
    Moby::RigidBodyPtr arm(new Moby::RigidBody());
    {
      Ravelin::SpatialRBInertiad J;                   // CylinderPrimitive(0.025, 1), mass 1, axis y: m (3 r^2 + h^2) / 12, m r^2 / 2
      J.m = 1.0; J.J[0] = J.J[2] = (3 * 0.025 * 0.025 + 1.0) / 12.0; J.J[1] = 0.5 * 0.025 * 0.025;
      arm->id = "arm";
      arm->set_inertia(J);
      arm->set_enabled(true);
      arm->get_recurrent_forces().push_back(g);
      arm->set_pose(Ravelin::Pose3d(Ravelin::Quatd(0, 0, 0, 1), Ravelin::Origin3d(0, -0.5, 0)));
      links.push_back(arm);
    }
    std::shared_ptr<Moby::RevoluteJoint> pivot(new Moby::RevoluteJoint());
    {
      Ravelin::Pose3d pose = base->get_pose();
      Ravelin::Vector3d position(pose.x.x(), pose.x.y(), pose.x.z());
      Ravelin::Vector3d axis(1, 0, 0);
      pivot->id = "pivot";
      pivot->set_location(position, base, arm);
      pivot->set_axis(axis);
      joints.push_back(pivot);
    }
    pendulum->set_links_and_joints(links, joints);
    pendulum->get_recurrent_forces().push_back(g);
    pendulum->set_floating_base(false);
    sim->add_dynamic_body(pendulum);
    sim->replicate(ne);

    for (int k = 1; k <= steps; k++) {
      sim->step(0.001);
      if (k % every == 0) {
        Ravelin::Pose3d pose = arm->get_pose();
        Ravelin::VectorNd q, qd;
        pendulum->get_generalized_coordinates_euler(q);
        pendulum->get_generalized_velocity(Moby::DynamicBodyd::eEuler, qd);
        printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", sim->current_time, pose.x[0], pose.x[1], pose.x[2], pose.q.x, pose.q.y,
               pose.q.z, pose.q.w, q[0], qd[0]);
      }
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
  return 0;
}
