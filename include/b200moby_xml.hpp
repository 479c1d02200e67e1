// b200moby_xml.hpp -- Moby::XMLReader::read for the C++ facade (include/b200moby.hpp): loads the subset of Moby's XML
// scene format that the accelerated path covers into facade objects, so a host program written like
// programs/driver.cpp (XMLReader::read(fname) -> map of id -> BasePtr, find the simulator, call step) runs unmodified.
//
// Mirrors (Moby tree, file:line):
//   XMLReader::read -> std::map<std::string, BasePtr>                         src/XMLReader.cpp:60-132
//   Primitive: mass | density, position, rpy | quat (w x y z)                  src/Primitive.cpp:240-300, XMLTree.cpp:407-421
//   Box xlen ylen zlen / Sphere radius / Plane                                 src/BoxPrimitive.cpp:649-651, SpherePrimitive.cpp:368
//   RigidBody: enabled, mass, position, rpy | quat, linear-velocity,
//     angular-velocity; InertiaFromPrimitive, CollisionGeometry children       src/RigidBody.cpp:165-330
//   GravityForce accel                                                          src/GravityForce.cpp:81
//   ContactParameters object1-id object2-id epsilon mu-coulomb mu-viscous
//     compliance friction-cone-edges                                            src/ContactParameters.cpp:57-136
//   TimeSteppingSimulator min-step-size, contact-dist-thresh,
//     constraint-stabilization-max-iterations; DynamicBody, RecurrentForce      src/TimeSteppingSimulator.cpp:470,
//                                                                               ConstraintSimulator.cpp:585-611, Simulator.cpp:860-928
// Constructs outside the subset (joints, articulated bodies, other primitives, geometry offsets, several geometries
// per body) throw std::runtime_error naming the construct: nothing is dropped silently.  The same rules as the Python
// loader (moby_b200/xml_scene.py); tests/test_cpp_facade.py checks the two against each other.
//
// No XML library is needed (libxml2 is not part of this build): the parser below handles what Moby's scene files use --
// elements, attributes in single or double quotes, self-closing tags, comments, the <?xml?> prolog.
#ifndef B200MOBY_XML_HPP
#define B200MOBY_XML_HPP

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "b200moby.hpp"

namespace Moby {

class XMLTree {
 public:
  std::string name;
  std::map<std::string, std::string> attribs;
  std::vector<std::shared_ptr<XMLTree> > children;
  const std::string* get_attrib(const std::string& k) const { auto it = attribs.find(k); return it == attribs.end() ? nullptr : &it->second; }
  void find_all(const std::string& tag, std::vector<const XMLTree*>& out) const {        // depth first, document order
    for (const auto& c : children) { if (c->name == tag) out.push_back(c.get()); c->find_all(tag, out); }
  }
  std::vector<const XMLTree*> child_nodes(const std::string& tag) const {
    std::vector<const XMLTree*> out;
    for (const auto& c : children) if (c->name == tag) out.push_back(c.get());
    return out;
  }
  static std::shared_ptr<XMLTree> parse(const std::string& text) {
    size_t pos = 0;
    std::shared_ptr<XMLTree> root(new XMLTree);
    root->name = "#document";
    parse_children(text, pos, *root, "");
    return root;
  }

 private:
  static void skip_ws(const std::string& s, size_t& p) { while (p < s.size() && std::isspace((unsigned char)s[p])) p++; }
  static void fail(const std::string& what, size_t p) { throw std::runtime_error("XML parse error at offset " + std::to_string(p) + ": " + what); }
  static void parse_children(const std::string& s, size_t& p, XMLTree& parent, const std::string& closing) {
    for (;;) {
      const size_t lt = s.find('<', p);
      if (lt == std::string::npos) { if (!closing.empty()) fail("missing </" + closing + ">", p); p = s.size(); return; }
      p = lt;
      if (s.compare(p, 4, "<!--") == 0) { const size_t e = s.find("-->", p + 4); if (e == std::string::npos) fail("unterminated comment", p); p = e + 3; continue; }
      if (s.compare(p, 2, "<?") == 0) { const size_t e = s.find("?>", p + 2); if (e == std::string::npos) fail("unterminated prolog", p); p = e + 2; continue; }
      if (s.compare(p, 2, "<!") == 0) { const size_t e = s.find('>', p); if (e == std::string::npos) fail("unterminated declaration", p); p = e + 1; continue; }
      if (s.compare(p, 2, "</") == 0) {
        const size_t e = s.find('>', p);
        if (e == std::string::npos) fail("unterminated closing tag", p);
        std::string nm = s.substr(p + 2, e - p - 2);
        while (!nm.empty() && std::isspace((unsigned char)nm.back())) nm.pop_back();
        if (nm != closing) fail("</" + nm + "> does not close <" + closing + ">", p);
        p = e + 1;
        return;
      }
      p++;                                                   // element
      size_t q = p;
      while (q < s.size() && !std::isspace((unsigned char)s[q]) && s[q] != '>' && s[q] != '/') q++;
      std::shared_ptr<XMLTree> node(new XMLTree);
      node->name = s.substr(p, q - p);
      if (node->name.empty()) fail("empty element name", p);
      p = q;
      bool selfclose = false;
      for (;;) {
        skip_ws(s, p);
        if (p >= s.size()) fail("unterminated element <" + node->name + ">", p);
        if (s[p] == '/') { selfclose = true; p++; skip_ws(s, p); if (p >= s.size() || s[p] != '>') fail("'/' not followed by '>'", p); p++; break; }
        if (s[p] == '>') { p++; break; }
        size_t k = p;
        while (k < s.size() && s[k] != '=' && !std::isspace((unsigned char)s[k]) && s[k] != '>' && s[k] != '/') k++;
        const std::string key = s.substr(p, k - p);
        p = k; skip_ws(s, p);
        if (p >= s.size() || s[p] != '=') fail("attribute '" + key + "' without a value", p);
        p++; skip_ws(s, p);
        if (p >= s.size() || (s[p] != '"' && s[p] != '\'')) fail("attribute '" + key + "': quoted value expected", p);
        const char quote = s[p++];
        const size_t e = s.find(quote, p);
        if (e == std::string::npos) fail("unterminated attribute value", p);
        node->attribs[key] = s.substr(p, e - p);
        p = e + 1;
      }
      parent.children.push_back(node);
      if (!selfclose) parse_children(s, p, *node, node->name);
    }
  }
};

class XMLReader {
 public:
  /// Reads a Moby XML scene file; every Box / Sphere / Plane, RigidBody, GravityForce, ContactParameters and the
  /// TimeSteppingSimulator appear in the returned map under their ids (the simulator under its id, or "simulator").
  static std::map<std::string, BasePtr> read(const std::string& fname) {
    std::ifstream in(fname.c_str());
    if (!in) throw std::runtime_error("XMLReader::read() - unable to open file " + fname);
    std::stringstream ss; ss << in.rdbuf();
    return read_from_string(ss.str());
  }
  static std::map<std::string, BasePtr> read_from_string(const std::string& text) {
    std::shared_ptr<XMLTree> doc = XMLTree::parse(text);
    std::vector<const XMLTree*> mobys; doc->find_all("MOBY", mobys);
    if (mobys.empty()) throw std::runtime_error("XMLReader::read() - no <MOBY> element");
    const XMLTree& moby = *mobys[0];
    std::map<std::string, BasePtr> id_map;
    static const char* unsupported_prims[] = {"Cone", "Cylinder", "Torus", "Heightmap", "TriangleMesh", "Polyhedron", "CSG", "GaussianMixture"};
    static const char* unsupported_bodies[] = {"RCArticulatedBody", "MCArticulatedBody", "RevoluteJoint", "PrismaticJoint", "FixedJoint", "SphericalJoint", "UniversalJoint"};
    for (const char* t : unsupported_bodies) { std::vector<const XMLTree*> v; moby.find_all(t, v); if (!v.empty()) throw std::runtime_error(std::string("<") + t + ">: articulated bodies are not loaded from XML on the accelerated path"); }
    std::vector<const XMLTree*> cgs; moby.find_all("CollisionGeometry", cgs);
    for (const char* t : unsupported_prims) {
      std::vector<const XMLTree*> v; moby.find_all(t, v);
      for (const XMLTree* n : v) for (const XMLTree* cg : cgs) if (n->get_attrib("id") && cg->get_attrib("primitive-id") && *n->get_attrib("id") == *cg->get_attrib("primitive-id"))
        throw std::runtime_error(std::string("<") + t + "> collision geometry is outside the accelerated path");
    }
    // primitives (their own pose is kept for the plane composition below)
    std::map<std::string, Ravelin::Pose3d> prim_pose;
    for (const char* t : {"Box", "Sphere", "Plane"}) {
      std::vector<const XMLTree*> v; moby.find_all(t, v);
      for (const XMLTree* n : v) {
        PrimitivePtr p;
        if (std::string(t) == "Box") p.reset(new BoxPrimitive(num(*n, "xlen", 1.0), num(*n, "ylen", 1.0), num(*n, "zlen", 1.0)));
        else if (std::string(t) == "Sphere") p.reset(new SpherePrimitive(num(*n, "radius", 1.0)));
        else p.reset(new PlanePrimitive);
        p->id = n->get_attrib("id") ? *n->get_attrib("id") : "";
        if (n->get_attrib("mass")) p->set_mass(num(*n, "mass", 0.0));
        else if (n->get_attrib("density")) p->set_density(num(*n, "density", 0.0));
        prim_pose[p->id] = pose(*n);
        id_map[p->id] = p;
      }
    }
    { std::vector<const XMLTree*> v; moby.find_all("GravityForce", v);
      for (const XMLTree* n : v) {
        std::shared_ptr<GravityForce> g(new GravityForce);
        g->id = n->get_attrib("id") ? *n->get_attrib("id") : "";
        const std::vector<double> a = vec(*n, "accel", 3);
        for (int k = 0; k < 3; k++) g->gravity[k] = a[k];
        id_map[g->id] = g;
      } }
    // collision-detection plugin of the simulator (XMLReader.cpp:334-388, ConstraintSimulator.cpp:562-572): only the one the
    // accelerated path has a built-in equivalent for -- the rimless wheel's, which finds its bodies by the ids WHEEL / GROUND
    std::string coldet_plugin;
    { std::vector<const XMLTree*> sv; moby.find_all("TimeSteppingSimulator", sv);
      if (sv.size() == 1 && sv[0]->get_attrib("collision-detection-plugin")) {
        std::vector<const XMLTree*> pv; moby.find_all("CollisionDetectionPlugin", pv);
        for (const XMLTree* pn : pv) if (pn->get_attrib("id") && *pn->get_attrib("id") == *sv[0]->get_attrib("collision-detection-plugin") && pn->get_attrib("plugin")) coldet_plugin = *pn->get_attrib("plugin");
        if (coldet_plugin != "librimless-wheel-coldet-plugin.so" && coldet_plugin != "libcontact-constrained-pendulum-coldet-plugin.so") throw std::runtime_error("collision-detection-plugin '" + coldet_plugin + "': no built-in equivalent on the accelerated path");
      } }
    // rigid bodies
    { std::vector<const XMLTree*> v; moby.find_all("RigidBody", v);
      for (const XMLTree* n : v) {
        RigidBodyPtr rb(new RigidBody);
        rb->id = n->get_attrib("id") ? *n->get_attrib("id") : "";
        const std::string* en = n->get_attrib("enabled");
        const bool enabled = !en || *en == "true" || *en == "1";
        rb->set_enabled(enabled);
        Ravelin::Pose3d bp = pose(*n);
        const std::vector<const XMLTree*> cg = n->child_nodes("CollisionGeometry");
        if (cg.size() > 1) throw std::runtime_error("body '" + rb->id + "': one CollisionGeometry per body on the accelerated path");
        if (!cg.empty()) {
          for (const char* k : {"relative-origin", "relative-rpy", "relative-quat"}) if (cg[0]->get_attrib(k)) throw std::runtime_error("body '" + rb->id + "': CollisionGeometry offsets are not supported");
          const std::string pid = cg[0]->get_attrib("primitive-id") ? *cg[0]->get_attrib("primitive-id") : "";
          PrimitivePtr p = std::dynamic_pointer_cast<Primitive>(lookup(id_map, pid));
          if (!p && pid.empty() && coldet_plugin == "librimless-wheel-coldet-plugin.so" && rb->id == "WHEEL") { p.reset(new RimlessWheelPrimitive(1.0, 0.0, 6)); prim_pose[pid] = Ravelin::Pose3d(); }   // params.h:4-6
          if (!p && pid.empty() && coldet_plugin == "libcontact-constrained-pendulum-coldet-plugin.so" && (rb->id == "l1" || rb->id == "world")) {   // contact-constrained-pendulum-coldet-plugin.cpp:21-37,60-75
            if (rb->id == "l1") p.reset(new PinPrimitive(0.0, 1.0, 0.0)); else p.reset(new PinWorldPrimitive);
            prim_pose[pid] = Ravelin::Pose3d();
          }
          if (!p) throw std::runtime_error("body '" + rb->id + "': primitive '" + pid + "' is not a Box / Sphere / Plane of this file");
          const Ravelin::Pose3d& pp = prim_pose[pid];
          if (p->shape == B200MOBY_SHAPE_PLANE) {
            if (enabled) throw std::runtime_error("body '" + rb->id + "': a Plane on an enabled body is not supported");
            bp = compose(bp, pp);                                 // static half-space: primitive pose composed with the body's
          } else if (!is_identity(pp)) throw std::runtime_error("body '" + rb->id + "': a posed primitive (geometry offset from the body frame) is not supported");
          CollisionGeometryPtr g(new CollisionGeometry);
          g->set_geometry(p);
          rb->geometries.push_back(g);
        }
        rb->set_pose(bp);
        const std::vector<const XMLTree*> ifp = n->child_nodes("InertiaFromPrimitive");
        if (ifp.size() > 1) throw std::runtime_error("body '" + rb->id + "': one InertiaFromPrimitive per body on the accelerated path");
        Ravelin::SpatialRBInertiad J; J.m = 1.0; J.J[0] = J.J[1] = J.J[2] = 1.0;
        if (!ifp.empty()) {
          const std::string pid = ifp[0]->get_attrib("primitive-id") ? *ifp[0]->get_attrib("primitive-id") : "";
          PrimitivePtr p = std::dynamic_pointer_cast<Primitive>(lookup(id_map, pid));
          if (!p || p->shape == B200MOBY_SHAPE_PLANE) throw std::runtime_error("body '" + rb->id + "': InertiaFromPrimitive needs a Box or Sphere of this file");
          if (p->get_mass() <= 0.0 && !has_mass_spec(moby, pid)) p->set_density(1.0);
          J = p->get_inertia();
        } else if (enabled && !(n->get_attrib("inertia") && n->get_attrib("mass"))) throw std::runtime_error("body '" + rb->id + "': an enabled body needs InertiaFromPrimitive or both mass and inertia attributes");
        if (n->get_attrib("mass")) J.m = num(*n, "mass", J.m);                     // RigidBody.cpp:182-188: J.m only
        if (n->get_attrib("inertia")) {                                            // RigidBody.cpp:191-197: rows separated by ';'
          std::string t = *n->get_attrib("inertia");
          for (char& c : t) if (c == ';') c = ' ';
          const std::vector<double> a = numbers(t);
          if (a.size() != 9 || a[1] != 0.0 || a[2] != 0.0 || a[3] != 0.0 || a[5] != 0.0 || a[6] != 0.0 || a[7] != 0.0) throw std::runtime_error("body '" + rb->id + "': the body frame must be the principal frame (diagonal inertia matrix)");
          J.J[0] = a[0]; J.J[1] = a[4]; J.J[2] = a[8];
        }
        rb->set_inertia(J);
        Ravelin::SVelocityd vel;
        if (n->get_attrib("linear-velocity")) { const std::vector<double> a = vec(*n, "linear-velocity", 3); for (int k = 0; k < 3; k++) vel.linear[k] = a[k]; }
        if (n->get_attrib("angular-velocity")) { const std::vector<double> a = vec(*n, "angular-velocity", 3); for (int k = 0; k < 3; k++) vel.angular[k] = a[k]; }
        rb->set_velocity(vel);
        id_map[rb->id] = rb;
      } }
    // the simulator
    std::vector<const XMLTree*> sims; moby.find_all("TimeSteppingSimulator", sims);
    if (sims.size() != 1) throw std::runtime_error("exactly one <TimeSteppingSimulator> expected (other simulators are outside the accelerated path)");
    const XMLTree& sn = *sims[0];
    std::shared_ptr<TimeSteppingSimulator> sim(new TimeSteppingSimulator);
    sim->id = sn.get_attrib("id") ? *sn.get_attrib("id") : "simulator";
    if (sn.get_attrib("min-step-size")) sim->min_step_size = num(sn, "min-step-size", sim->min_step_size);
    if (sn.get_attrib("contact-dist-thresh")) sim->contact_dist_thresh = num(sn, "contact-dist-thresh", sim->contact_dist_thresh);
    // absent: the reference's default, stabilize after every step without an iteration limit (ConstraintStabilization.cpp:56-59)
    if (sn.get_attrib("constraint-stabilization-max-iterations")) sim->cstab.max_iterations = (unsigned)num(sn, "constraint-stabilization-max-iterations", 0.0);
    std::vector<RecurrentForcePtr> forces;
    for (const XMLTree* rf : sn.child_nodes("RecurrentForce")) {
      const std::string fid = rf->get_attrib("recurrent-force-id") ? *rf->get_attrib("recurrent-force-id") : "";
      RecurrentForcePtr f = std::dynamic_pointer_cast<RecurrentForce>(lookup(id_map, fid));
      if (!f) throw std::runtime_error("RecurrentForce '" + fid + "': only GravityForce is on the accelerated path");
      forces.push_back(f);
    }
    for (const XMLTree* db : sn.child_nodes("DynamicBody")) {
      const std::string bid = db->get_attrib("dynamic-body-id") ? *db->get_attrib("dynamic-body-id") : "";
      RigidBodyPtr rb = std::dynamic_pointer_cast<RigidBody>(lookup(id_map, bid));
      if (!rb) throw std::runtime_error("DynamicBody '" + bid + "' is not a <RigidBody> of this file");
      for (const RecurrentForcePtr& f : forces) rb->get_recurrent_forces().push_back(f);     // Simulator.cpp:928-950
      sim->add_dynamic_body(rb);
    }
    int ncp = 0;
    for (const XMLTree* cp : sn.child_nodes("ContactParameters")) {
      const std::string a = cp->get_attrib("object1-id") ? *cp->get_attrib("object1-id") : "", b = cp->get_attrib("object2-id") ? *cp->get_attrib("object2-id") : "";
      BasePtr oa = lookup(id_map, a), ob = lookup(id_map, b);
      if (!oa || !ob) throw std::runtime_error("ContactParameters '" + a + "' / '" + b + "': objects must be rigid bodies of this file");
      std::shared_ptr<ContactParameters> c(new ContactParameters(oa, ob));
      c->id = "contact-parameters-" + std::to_string(ncp++);
      c->epsilon = num(*cp, "epsilon", 0.0); c->mu_coulomb = num(*cp, "mu-coulomb", 0.0); c->mu_viscous = num(*cp, "mu-viscous", 0.0);
      c->compliance = num(*cp, "compliance", 0.0);
      const double nk = num(*cp, "friction-cone-edges", 4.0);
      c->NK = nk < 4.0 ? 4u : (unsigned)nk;                                                  // ContactParameters.cpp:132-136
      sim->add_contact_parameters(c);
      id_map[c->id] = c;
    }
    for (const XMLTree* dp : sn.child_nodes("DisabledPair")) {
      const std::string a = dp->get_attrib("object1-id") ? *dp->get_attrib("object1-id") : "", b = dp->get_attrib("object2-id") ? *dp->get_attrib("object2-id") : "";
      BasePtr oa = lookup(id_map, a), ob = lookup(id_map, b);
      if (oa && ob && a != b) sim->disabled_pairs.push_back(std::make_pair(oa, ob));
    }
    id_map[sim->id] = sim;
    return id_map;
  }

 private:
  static BasePtr lookup(const std::map<std::string, BasePtr>& m, const std::string& id) { auto it = m.find(id); return it == m.end() ? BasePtr() : it->second; }
  static std::vector<double> numbers(const std::string& s) {
    std::vector<double> v; std::string t = s;
    for (char& c : t) if (c == ',') c = ' ';
    std::istringstream is(t); double x;
    while (is >> x) v.push_back(x);
    return v;
  }
  static double num(const XMLTree& n, const char* key, double dflt) { const std::string* a = n.get_attrib(key); return a ? std::strtod(a->c_str(), nullptr) : dflt; }
  static std::vector<double> vec(const XMLTree& n, const char* key, size_t len) {
    const std::string* a = n.get_attrib(key);
    std::vector<double> v = a ? numbers(*a) : std::vector<double>();
    if (v.size() != len) throw std::runtime_error(std::string("attribute '") + key + "' of <" + n.name + ">: " + std::to_string(len) + " numbers expected");
    return v;
  }
  static bool has_mass_spec(const XMLTree& moby, const std::string& pid) {
    for (const char* t : {"Box", "Sphere"}) { std::vector<const XMLTree*> v; moby.find_all(t, v); for (const XMLTree* n : v) if (n->get_attrib("id") && *n->get_attrib("id") == pid) return n->get_attrib("mass") || n->get_attrib("density"); }
    return false;
  }
  // Quatd::rpy: rotation about x by roll, then y by pitch, then z by yaw (fixed axes), as moby_b200/scenes.py quat_from_rpy
  static Ravelin::Quatd quat_rpy(double roll, double pitch, double yaw) {
    const double cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5), cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5), cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5);
    return Ravelin::Quatd(sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy);
  }
  static Ravelin::Pose3d pose(const XMLTree& n) {
    Ravelin::Origin3d x(0, 0, 0);
    if (n.get_attrib("position")) { const std::vector<double> a = vec(n, "position", 3); x = Ravelin::Origin3d(a[0], a[1], a[2]); }
    Ravelin::Quatd q(0, 0, 0, 1);
    if (n.get_attrib("quat")) { const std::vector<double> a = vec(n, "quat", 4); q = Ravelin::Quatd(a[1], a[2], a[3], a[0]); }   // XMLTree.cpp:415-419: w x y z
    else if (n.get_attrib("rpy")) { const std::vector<double> a = vec(n, "rpy", 3); q = quat_rpy(a[0], a[1], a[2]); }
    else if (n.get_attrib("aangle")) {
      const std::vector<double> a = vec(n, "aangle", 4);
      const double nrm = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), s = std::sin(0.5 * a[3]) / nrm;
      q = Ravelin::Quatd(a[0] * s, a[1] * s, a[2] * s, std::cos(0.5 * a[3]));
    }
    const double nq = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q = Ravelin::Quatd(q.x / nq, q.y / nq, q.z / nq, q.w / nq);
    return Ravelin::Pose3d(q, x);
  }
  static bool is_identity(const Ravelin::Pose3d& p) { return p.x[0] == 0.0 && p.x[1] == 0.0 && p.x[2] == 0.0 && std::fabs(std::fabs(p.q.w) - 1.0) < 1e-15; }
  static Ravelin::Pose3d compose(const Ravelin::Pose3d& a, const Ravelin::Pose3d& b) {       // a * b: b expressed in a's frame
    const Ravelin::Quatd& p = a.q; const Ravelin::Quatd& q = b.q;
    Ravelin::Quatd r(p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y, p.w * q.y - p.x * q.z + p.y * q.w + p.z * q.x,
                     p.w * q.z + p.x * q.y - p.y * q.x + p.z * q.w, p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z);
    // rotate b.x by a.q
    const double vx = b.x[0], vy = b.x[1], vz = b.x[2];
    const double tx = 2.0 * (p.y * vz - p.z * vy), ty = 2.0 * (p.z * vx - p.x * vz), tz = 2.0 * (p.x * vy - p.y * vx);
    Ravelin::Origin3d x(a.x[0] + vx + p.w * tx + (p.y * tz - p.z * ty), a.x[1] + vy + p.w * ty + (p.z * tx - p.x * tz), a.x[2] + vz + p.w * tz + (p.x * ty - p.y * tx));
    return Ravelin::Pose3d(r, x);
  }
};

}  // namespace Moby

#endif  // B200MOBY_XML_HPP
