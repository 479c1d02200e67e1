// Host program written against Moby's names (cf. Moby example/sims-in-code/block.cpp and
// example/simple-contact/simplest.xml): a unit box resting on the plane y = 0, stepped with
// TimeSteppingSimulator::step, printing rows in the moby-regress format (programs/regress.cpp:78-95).
// usage: sitting_box <steps> <print_every> [n_envs]
#include <cstdio>
#include <cstdlib>
#include "b200moby.hpp"

static int post_steps = 0;
static void post_step(Moby::TimeSteppingSimulator*) { post_steps++; }

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 1000, every = argc > 2 ? atoi(argv[2]) : 100, n_envs = argc > 3 ? atoi(argv[3]) : 1;
  try {
    std::shared_ptr<Moby::TimeSteppingSimulator> sim(new Moby::TimeSteppingSimulator());
    std::shared_ptr<Moby::GravityForce> g(new Moby::GravityForce());
    g->gravity = Ravelin::Vector3d(0, -9.81, 0);

    Moby::PrimitivePtr b1(new Moby::BoxPrimitive(1, 1, 1));
    b1->set_density(1.0);
    Moby::PrimitivePtr halfspace(new Moby::PlanePrimitive());

    Moby::RigidBodyPtr box(new Moby::RigidBody());
    box->id = "box";
    box->set_inertia(b1->get_inertia());
    box->set_enabled(true);
    box->get_recurrent_forces().push_back(g);
    Moby::CollisionGeometryPtr cg1(new Moby::CollisionGeometry());
    cg1->set_geometry(b1);
    box->geometries.push_back(cg1);
    box->set_pose(Ravelin::Pose3d(Ravelin::Quatd(0, 0, 0, 1), Ravelin::Origin3d(0, 0.50001, 0)));

    Moby::RigidBodyPtr ground(new Moby::RigidBody());
    ground->id = "ground";
    ground->set_enabled(false);
    Moby::CollisionGeometryPtr cg2(new Moby::CollisionGeometry());
    cg2->set_geometry(halfspace);
    ground->geometries.push_back(cg2);

    sim->add_dynamic_body(box);
    sim->add_dynamic_body(ground);
    std::shared_ptr<Moby::ContactParameters> cp(new Moby::ContactParameters(ground, box));
    cp->epsilon = 0; cp->mu_coulomb = 0; cp->mu_viscous = 0; cp->NK = 8;
    sim->add_contact_parameters(cp);
    sim->cstab.max_iterations = 0;
    sim->post_step_callback_fn = post_step;
    sim->replicate(n_envs);

    for (int i = 1; i <= steps; i++) {
      sim->step(0.001);
      if (i % every == 0) {
        for (int e = 0; e < n_envs; e += (n_envs > 1 ? n_envs - 1 : 1)) {
          Ravelin::Pose3d p = box->get_pose(e);
          printf("%.3f %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", sim->current_time, e, p.x[0], p.x[1], p.x[2], p.q.x, p.q.y, p.q.z, p.q.w);
        }
      }
    }
    b200moby_counters c = sim->counters();
    printf("# env_steps %lld lcp_solves %lld max_lcp_n %lld lcp_failures %lld post_steps %d\n", c.env_steps, c.lcp_solves, c.max_lcp_n, c.lcp_failures, post_steps);

    // Moby::LCP on a 2x2 problem: M = [[2,1],[1,2]], q = [-5,-6]  ->  z = [4/3, 7/3]
    Moby::LCP lcp;
    Ravelin::MatrixNd M(2, 2); M(0, 0) = 2; M(0, 1) = 1; M(1, 0) = 1; M(1, 1) = 2;
    Ravelin::VectorNd q = {-5, -6}, z;
    const bool ok = lcp.lcp_lemke(M, q, z);
    printf("# lcp_lemke %d %.17g %.17g\n", ok ? 1 : 0, z[0], z[1]);
    Ravelin::VectorNd z2;
    const bool ok2 = lcp.lcp_fast(M, q, z2);
    printf("# lcp_fast %d %.17g %.17g\n", ok2 ? 1 : 0, z2[0], z2[1]);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
  return 0;
}
