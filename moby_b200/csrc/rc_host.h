// Host-side validation of b200moby_rc_desc and its translation into the device's RCTree.
#pragma once
#include <cmath>
#include <cstring>
#include "../../include/b200moby.h"
#include "sim_device.cuh"

// Returns nullptr on success, else a static message; *unsupported tells INVALID from UNSUPPORTED.
inline const char* b2m_rc_tree_from_desc(const b200moby_rc_desc& r, int n_bodies, b2m::RCTree& T, bool* unsupported) {
  using namespace b2m;
  *unsupported = false;
  if (r.n_links < 2 || r.n_links > B2M_MAX_LINKS || r.first_body < 0 || r.first_body + r.n_links > n_bodies)
    return "articulated body: 2 <= n_links <= 16 and the links must be bodies of the scene";
  if (!r.parent || !r.joint_type || !r.joint_axis || !r.loc_parent || !r.loc_child || !r.rel_quat) return "articulated body: null array";
  if (r.fdyn_algorithm != B200MOBY_FDYN_FSAB && r.fdyn_algorithm != B200MOBY_FDYN_CRB) return "articulated body: unknown fdyn_algorithm";
  memset(&T, 0, sizeof(T));
  T.n_links = r.n_links; T.first_body = r.first_body; T.fdyn = r.fdyn_algorithm; T.has_ctrl = r.ctrl_kp ? 1 : 0;
  for (int i = 1; i < r.n_links; i++) {
    if (r.parent[i] < 0 || r.parent[i] >= i) return "articulated body: parent[i] must be in [0, i)";
    if (r.joint_type[i] != B200MOBY_JOINT_REVOLUTE && r.joint_type[i] != B200MOBY_JOINT_PRISMATIC) { *unsupported = true; return "articulated body: only revolute and prismatic joints are on the accelerated path"; }
    T.parent[i] = r.parent[i]; T.jtype[i] = r.joint_type[i];
    double an = 0.0, qn = 0.0;
    for (int c = 0; c < 3; c++) an += r.joint_axis[3 * i + c] * r.joint_axis[3 * i + c];
    for (int c = 0; c < 4; c++) qn += r.rel_quat[4 * i + c] * r.rel_quat[4 * i + c];
    an = std::sqrt(an); qn = std::sqrt(qn);
    if (!(an > 0.0) || !(qn > 0.0)) return "articulated body: zero joint axis or quaternion";
    double qt[4];
    for (int c = 0; c < 3; c++) { T.axis[i][c] = r.joint_axis[3 * i + c] / an; T.loc_parent[i][c] = r.loc_parent[3 * i + c]; T.loc_child[i][c] = r.loc_child[3 * i + c]; }
    for (int c = 0; c < 4; c++) qt[c] = r.rel_quat[4 * i + c] / qn;
    quat_to_R(qt, T.R0[i]);
    if (r.ctrl_kp) {
      T.kp[i - 1] = r.ctrl_kp[i - 1]; T.kv[i - 1] = r.ctrl_kv ? r.ctrl_kv[i - 1] : 0.0;
      T.amp[i - 1] = r.ctrl_amp ? r.ctrl_amp[i - 1] : 0.0; T.freq[i - 1] = r.ctrl_freq ? r.ctrl_freq[i - 1] : 0.0;
    }
  }
  return nullptr;
}

// Largest island dimension of a scene with an articulated body: six coordinates per enabled free body plus the joints.
inline int b2m_dense_ngc(const b200moby_scene_desc* d) {
  if (!d->rc || d->rc->n_links < 2) return 0;
  const int ne = d->n_envs, nb = d->n_bodies, f = d->rc->first_body, nl = d->rc->n_links;
  int best = 0;
  for (int e = 0; e < ne; e++) {
    int n = 0;
    for (int b = 0; b < nb; b++) if ((b < f || b >= f + nl) && d->enabled[(size_t)b * ne + e]) n += 6;
    best = n > best ? n : best;
  }
  return best + nl - 1;
}
