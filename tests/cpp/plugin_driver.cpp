// The driver's side of the plugin seam (programs/driver.cpp:310-349,560-575): build the scene, put everything into a
// read map, dlopen the controller plugin, call its extern "C" init, step.  The scene: a two-link arm (RCArticulatedBody,
// revolute joints "shoulder" and "elbow" about y) next to a sphere resting on a ground plane.
// usage: plugin_driver <plugin.so> <steps> <print_every>; prints "t q0 q1 qd0 qd1 sphere_z" then "# callbacks n min_contacts m".
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <b200moby.hpp>

typedef void (*init_t)(void*, const std::map<std::string, Moby::BasePtr>&, double);

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: plugin_driver <plugin.so> <steps> <print_every>\n"); return 2; }
  const int steps = atoi(argv[2]), every = atoi(argv[3]);
  try {
    std::shared_ptr<Moby::TimeSteppingSimulator> sim(new Moby::TimeSteppingSimulator());
    sim->id = "simulator";
    std::shared_ptr<Moby::GravityForce> g(new Moby::GravityForce());
    g->gravity = Ravelin::Vector3d(0, 0, -9.81);
    Moby::RCArticulatedBodyPtr arm(new Moby::RCArticulatedBody());
    arm->id = "arm2";
    std::vector<Moby::RigidBodyPtr> links;
    std::vector<Moby::JointPtr> joints;
    const double cx[3] = {0.0, 0.25, 0.75};
    for (int i = 0; i < 3; i++) {
      Moby::RigidBodyPtr l(new Moby::RigidBody());
      Moby::PrimitivePtr box(new Moby::BoxPrimitive(0.5, 0.05, 0.05));
      box->set_mass(1.0);
      l->id = i == 0 ? "base" : (i == 1 ? "upper" : "fore");
      l->set_inertia(box->get_inertia());
      l->set_enabled(i > 0);
      l->set_pose(Ravelin::Pose3d(Ravelin::Quatd(0, 0, 0, 1), Ravelin::Origin3d(cx[i], 0, 1.0)));
      links.push_back(l);
    }
    const char* jn[2] = {"shoulder", "elbow"};
    for (int k = 0; k < 2; k++) {
      std::shared_ptr<Moby::RevoluteJoint> j(new Moby::RevoluteJoint());
      j->id = jn[k];
      j->set_location(Ravelin::Vector3d(0.5 * k, 0, 1.0), links[k], links[k + 1]);
      j->set_axis(Ravelin::Vector3d(0, 1, 0));
      joints.push_back(j);
    }
    arm->set_links_and_joints(links, joints);
    arm->get_recurrent_forces().push_back(g);
    arm->set_floating_base(false);
    sim->add_dynamic_body(arm);

    Moby::RigidBodyPtr ball(new Moby::RigidBody());
    {
      Moby::PrimitivePtr sp(new Moby::SpherePrimitive(0.1));
      sp->set_mass(1.0);
      ball->id = "ball";
      ball->set_inertia(sp->get_inertia());
      Moby::CollisionGeometryPtr cg(new Moby::CollisionGeometry());
      cg->set_geometry(sp);
      ball->geometries.push_back(cg);
      ball->get_recurrent_forces().push_back(g);
      ball->set_pose(Ravelin::Pose3d(Ravelin::Quatd(0, 0, 0, 1), Ravelin::Origin3d(2.0, 0, 0.1)));
      sim->add_dynamic_body(ball);
    }
    Moby::RigidBodyPtr ground(new Moby::RigidBody());
    {
      Moby::PrimitivePtr pl(new Moby::PlanePrimitive());
      ground->id = "ground";
      ground->set_enabled(false);
      Moby::CollisionGeometryPtr cg(new Moby::CollisionGeometry());
      cg->set_geometry(pl);
      ground->geometries.push_back(cg);
      const double h = std::sqrt(0.5);
      ground->set_pose(Ravelin::Pose3d(Ravelin::Quatd(h, 0, 0, h), Ravelin::Origin3d(0, 0, 0)));      // plane normal +y -> +z
      sim->add_dynamic_body(ground);
    }
    sim->cstab.max_iterations = 0;

    std::map<std::string, Moby::BasePtr> read_map;
    read_map["simulator"] = sim; read_map["arm2"] = arm; read_map["ball"] = ball; read_map["ground"] = ground; read_map["gravity"] = g;
    void* plugin = dlopen(argv[1], RTLD_LAZY);
    if (!plugin) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    init_t init = (init_t)dlsym(plugin, "init");                                  // programs/driver.cpp:333-345
    if (!init) { fprintf(stderr, "no init symbol\n"); return 2; }
    (*init)(nullptr, read_map, 0.001);

    for (int k = 1; k <= steps; k++) {
      sim->step(0.001);
      if (k % every == 0) {
        Ravelin::VectorNd q, qd;
        arm->get_generalized_coordinates_euler(q);
        arm->get_generalized_velocity(Moby::DynamicBodyd::eEuler, qd);
        printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", sim->current_time, q[0], q[1], qd[0], qd[1], ball->get_pose().x[2]);
      }
    }
    int (*calls)() = (int (*)())dlsym(plugin, "plugin_callback_calls");
    int (*minc)() = (int (*)())dlsym(plugin, "plugin_min_contacts");
    printf("# callbacks %d min_contacts %d\n", calls ? calls() : -1, minc ? minc() : -1);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
  return 0;
}
