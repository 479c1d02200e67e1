// Loads a Moby XML scene through the C++ facade's XMLReader (include/b200moby_xml.hpp) and prints what it built, one
// record per line, for tests/test_cpp_facade.py to compare with the Python loader.  Does not touch the GPU.
#include <cstdio>
#include <b200moby_xml.hpp>

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: xml_dump scene.xml\n"); return 2; }
  try {
    std::map<std::string, Moby::BasePtr> m = Moby::XMLReader::read(argv[1]);
    std::shared_ptr<Moby::TimeSteppingSimulator> sim;
    for (auto& kv : m) if (auto s = std::dynamic_pointer_cast<Moby::TimeSteppingSimulator>(kv.second)) sim = s;     // programs/driver.cpp:560-575
    if (!sim) { fprintf(stderr, "no simulator\n"); return 1; }
    printf("sim %s min_step_size %.17g contact_dist_thresh %.17g stabilization %u\n", sim->id.c_str(), sim->min_step_size, sim->contact_dist_thresh, sim->cstab.max_iterations);
    const auto& bodies = sim->get_dynamic_bodies();
    for (size_t i = 0; i < bodies.size(); i++) {
      Moby::RigidBody& rb = *std::dynamic_pointer_cast<Moby::RigidBody>(bodies[i]);     // as Moby programs do (coldet-plugin.cpp:30-35)
      Moby::PrimitivePtr p = rb.geometries.empty() ? Moby::PrimitivePtr() : rb.geometries.front()->get_geometry();
      const Ravelin::Pose3d ps = rb.get_pose(0);
      const Ravelin::SVelocityd v = rb.get_velocity(0);
      double g[3] = {0, 0, 0};
      for (auto& f : rb.get_recurrent_forces()) if (auto gf = std::dynamic_pointer_cast<Moby::GravityForce>(f)) for (int k = 0; k < 3; k++) g[k] += gf->gravity[k];
      printf("body %zu %s enabled %d shape %d dims %.17g %.17g %.17g mass %.17g J %.17g %.17g %.17g q %.17g %.17g %.17g %.17g %.17g %.17g %.17g v %.17g %.17g %.17g %.17g %.17g %.17g g %.17g %.17g %.17g\n",
             i, rb.id.c_str(), rb.is_enabled() ? 1 : 0, p ? p->shape : 0, p ? p->dims[0] : 0.0, p ? p->dims[1] : 0.0, p ? p->dims[2] : 0.0,
             rb.get_inertia().m, rb.get_inertia().J[0], rb.get_inertia().J[1], rb.get_inertia().J[2],
             ps.x[0], ps.x[1], ps.x[2], ps.q.x, ps.q.y, ps.q.z, ps.q.w, v.linear[0], v.linear[1], v.linear[2], v.angular[0], v.angular[1], v.angular[2], g[0], g[1], g[2]);
    }
    for (auto& kv : sim->contact_params) {
      int a = -1, b = -1;
      for (size_t i = 0; i < bodies.size(); i++) { if (bodies[i].get() == kv.first.first.get()) a = (int)i; if (bodies[i].get() == kv.first.second.get()) b = (int)i; }
      if (a < 0 || b < 0) continue;               // parameters for bodies the simulator does not register
      const Moby::ContactParameters& c = *kv.second;
      printf("contact %d %d eps %.17g mu_c %.17g mu_v %.17g compliance %.17g NK %u\n", a < b ? a : b, a < b ? b : a, c.epsilon, c.mu_coulomb, c.mu_viscous, c.compliance, c.NK);
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
