// mjcf_compile.cpp — MJCF subset -> b2mjModel.
//
// Stands in for mj_loadXML at the reference's load path (mujoco_ros/src/mujoco_env.cpp:840-843).
// Subset: compiler(angle, eulerseq, autolimits, inertiafromgeom), option(+flag), size(nconmax,njmax),
// default classes, worldbody tree (body, inertial, joint, freejoint, geom, site; light/camera ignored),
// actuator (motor, position, velocity, general), tendon/fixed, equality (connect, weld, joint, tendon),
// contact/exclude, sensor.  visual/asset/statistic/keyframe are accepted and ignored.
// It parses the reference's five worlds (pendulum_world, sensors_world, equality_world, empty_world,
// mocap_world) unmodified; tests/test_model_compile.py pins the facts the reference's tests pin
// (ros_interface_test.cpp:290-298 qpos0; :769-829 equality data).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "hostmath.h"
#include "mesh.h"
#include "png_gray.h"
#include "model_core.h"
#include "xml_mini.h"

namespace b2mj {
using namespace hm;

void model_body_poses0(const b2mjModel* m, std::vector<double>& xpos, std::vector<double>& xquat);

namespace {

const double kPi = 3.14159265358979323846;

struct CompileError {
  std::string msg;
};
[[noreturn]] void fail(const XmlNode* n, const std::string& msg) {
  throw CompileError{"MJCF line " + std::to_string(n ? n->line : 0) + " <" + (n ? n->tag : std::string("?")) + ">: " + msg};
}

typedef std::map<std::string, std::string> AttrMap;

struct Ctx {
  bool degrees = true;
  bool autolimits = true;
  std::string eulerseq = "xyz";
  int inertiafromgeom = 2;  // 0 false, 1 true, 2 auto
  double boundmass = 0, boundinertia = 0, settotalmass = -1;
  int inertiagroup[2] = {0, 5};  // geom groups that count for inertia inference (compiler inertiagrouprange)
  std::map<std::string, std::map<std::string, AttrMap>> defaults;  // class -> group -> attrs
  std::map<std::string, std::string> class_parent;
  std::string meshdir;
  std::vector<MeshData> meshes;
  std::map<std::string, int> mesh_id;
  struct HField { int nrow = 0, ncol = 0; double size[4] = {0, 0, 0, 0}; std::vector<double> data; };
  std::vector<HField> hfields;
  std::map<std::string, int> hfield_id;
};

std::vector<double> parse_nums(const std::string& s) {
  std::vector<double> v;
  std::istringstream is(s);
  std::string tok;
  while (is >> tok) v.push_back(std::strtod(tok.c_str(), nullptr));
  return v;
}

std::string group_of(const std::string& tag) {
  if (tag == "motor" || tag == "position" || tag == "velocity" || tag == "general" || tag == "intvelocity" || tag == "damper" ||
      tag == "cylinder")
    return "actuator";
  if (tag == "freejoint") return "freejoint";
  if (tag == "fixed" || tag == "spatial") return "tendon";
  if (tag == "connect" || tag == "weld") return "equality";
  return tag;
}

// effective attributes of element n: class defaults overlaid by own attributes
AttrMap effective(const Ctx& c, const XmlNode* n, const std::string& childclass) {
  AttrMap out;
  std::string cls = "main";
  if (auto* a = n->attr("class")) cls = *a;
  else if (!childclass.empty()) cls = childclass;
  std::string g = group_of(n->tag);
  if (g != "freejoint") {
    auto it = c.defaults.find(cls);
    if (it == c.defaults.end() && cls != "main") fail(n, "unknown default class '" + cls + "'");
    if (it != c.defaults.end()) {
      auto jt = it->second.find(g);
      if (jt != it->second.end()) out = jt->second;
    }
  }
  for (auto& kv : n->attrs) out[kv.first] = kv.second;
  return out;
}

struct A {  // attribute accessor with typed getters
  const AttrMap& m;
  const XmlNode* n;
  bool has(const char* k) const { return m.count(k) > 0; }
  std::string str(const char* k, const std::string& d = "") const {
    auto it = m.find(k);
    return it == m.end() ? d : it->second;
  }
  double num(const char* k, double d) const {
    auto it = m.find(k);
    if (it == m.end()) return d;
    auto v = parse_nums(it->second);
    if (v.size() != 1) fail(n, std::string("attribute '") + k + "' needs 1 number");
    return v[0];
  }
  int integer(const char* k, int d) const { return (int)std::lround(num(k, d)); }
  // fills up to n values; returns how many were given
  int vec(const char* k, double* dst, int nmax, int nmin = 0) const {
    auto it = m.find(k);
    if (it == m.end()) return 0;
    auto v = parse_nums(it->second);
    if ((int)v.size() > nmax || (int)v.size() < std::max(nmin, 1))
      fail(n, std::string("attribute '") + k + "' has wrong number of values");
    for (size_t i = 0; i < v.size(); i++) dst[i] = v[i];
    return (int)v.size();
  }
  // tri-state: -1 auto / 0 / 1
  int tristate(const char* k, int d) const {
    auto it = m.find(k);
    if (it == m.end()) return d;
    if (it->second == "true") return 1;
    if (it->second == "false") return 0;
    if (it->second == "auto") return -1;
    fail(n, std::string("attribute '") + k + "' must be true/false/auto");
  }
};

void euler2quat(const Ctx& c, const double* e_in, double* q) {
  double e[3] = {e_in[0], e_in[1], e_in[2]};
  if (c.degrees) for (double& x : e) x *= kPi / 180;
  q[0] = 1; q[1] = q[2] = q[3] = 0;
  for (int i = 0; i < 3; i++) {
    char ch = c.eulerseq[i];
    double ax[3] = {0, 0, 0};
    int k = (ch == 'x' || ch == 'X') ? 0 : (ch == 'y' || ch == 'Y') ? 1 : 2;
    ax[k] = 1;
    double r[4];
    axisangle2quat(r, ax, e[i]);
    if (ch >= 'a') mulquat(q, q, r);  // intrinsic (rotating frame)
    else mulquat(q, r, q);            // extrinsic (fixed frame)
  }
  normalize4(q);
}

// resolve quat / euler / axisangle / zaxis / xyaxes into q (identity if none)
void orientation(const Ctx& c, const A& a, double* q) {
  q[0] = 1; q[1] = q[2] = q[3] = 0;
  double v[6];
  if (a.has("quat")) {
    a.vec("quat", q, 4, 4);
    normalize4(q);
  } else if (a.has("euler")) {
    a.vec("euler", v, 3, 3);
    euler2quat(c, v, q);
  } else if (a.has("axisangle")) {
    double aa[4];
    a.vec("axisangle", aa, 4, 4);
    double ang = c.degrees ? aa[3] * kPi / 180 : aa[3];
    normalize3(aa);
    axisangle2quat(q, aa, ang);
  } else if (a.has("zaxis")) {
    a.vec("zaxis", v, 3, 3);
    z2quat(q, v);
  } else if (a.has("xyaxes")) {
    a.vec("xyaxes", v, 6, 6);
    double x[3] = {v[0], v[1], v[2]}, y[3] = {v[3], v[4], v[5]}, z[3];
    normalize3(x);
    double d = dot3(x, y);
    for (int i = 0; i < 3; i++) y[i] -= d * x[i];
    normalize3(y);
    cross(z, x, y);
    double mat[9] = {x[0], y[0], z[0], x[1], y[1], z[1], x[2], y[2], z[2]};
    mat2quat(q, mat);
  }
}

struct CJoint {
  std::string name;
  int type = B2MJ_JNT_HINGE;
  double pos[3] = {0, 0, 0}, axis[3] = {0, 0, 1};
  double range[2] = {0, 0};
  int limited = 0;
  double margin = 0, ref = 0, springref = 0, stiffness = 0, damping = 0, armature = 0, frictionloss = 0;
  double springdamper[2] = {0, 0};  // (time constant, damping ratio): stiffness / damping derived from dof_invweight0 after set_const
  double solref_lim[2] = {0.02, 1}, solimp_lim[5] = {0.9, 0.95, 0.001, 0.5, 2};
  double solref_fri[2] = {0.02, 1}, solimp_fri[5] = {0.9, 0.95, 0.001, 0.5, 2};
};
struct CGeom {
  std::string name;
  int type = B2MJ_GEOM_SPHERE;
  double size[3] = {0, 0, 0}, pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};
  double friction[3] = {1, 0.005, 0.0001};
  double solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2};
  double solmix = 1, margin = 0, gap = 0, mass = 0;
  double inertia[3] = {0, 0, 0};
  double rgba[4] = {0.5, 0.5, 0.5, 1};
  int meshid = -1;
  double mesh_volume = 0, mesh_inertia[3] = {0, 0, 0}, rbound = 0;
  int contype = 1, conaffinity = 1, condim = 3, priority = 0, group = 0;
};
struct CSite {
  std::string name;
  int type = B2MJ_GEOM_SPHERE;
  double size[3] = {0.005, 0.005, 0.005}, pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};
};
struct CBody {
  std::string name;
  int parent = 0;
  double pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};
  bool mocap = false, has_inertial = false;
  double ipos[3] = {0, 0, 0}, iquat[4] = {1, 0, 0, 0}, mass = 0, inertia[3] = {0, 0, 0};
  double gravcomp = 0;
  std::vector<CJoint> joints;
  std::vector<CGeom> geoms;
  std::vector<CSite> sites;
};

int geom_type_from(const XmlNode* n, const std::string& s) {
  if (s == "plane") return B2MJ_GEOM_PLANE;
  if (s == "sphere") return B2MJ_GEOM_SPHERE;
  if (s == "capsule") return B2MJ_GEOM_CAPSULE;
  if (s == "ellipsoid") return B2MJ_GEOM_ELLIPSOID;
  if (s == "cylinder") return B2MJ_GEOM_CYLINDER;
  if (s == "box") return B2MJ_GEOM_BOX;
  if (s == "mesh") return B2MJ_GEOM_MESH;
  if (s == "hfield") return B2MJ_GEOM_HFIELD;
  fail(n, "unsupported geom/site type '" + s + "'");
}

// volume and principal inertia (unit density scaled by mass later) of a primitive
void geom_mass_props(const CGeom& g, double density, bool explicit_mass, double* mass, double* inertia) {
  const double r = g.size[0], h = g.size[1];
  double vol = 0, I[3] = {0, 0, 0};  // I per unit mass
  switch (g.type) {
    case B2MJ_GEOM_SPHERE:
      vol = 4.0 / 3.0 * kPi * r * r * r;
      I[0] = I[1] = I[2] = 0.4 * r * r;
      break;
    case B2MJ_GEOM_CAPSULE: {
      double L = 2 * h;
      double vc = kPi * r * r * L, vs = 4.0 / 3.0 * kPi * r * r * r;
      vol = vc + vs;
      double fc = vc / vol, fs = vs / vol;  // mass fractions
      I[0] = I[1] = fc * (3 * r * r + L * L) / 12 + fs * (0.4 * r * r + L * L / 4 + 0.375 * L * r);
      I[2] = fc * r * r / 2 + fs * 0.4 * r * r;
      break;
    }
    case B2MJ_GEOM_CYLINDER: {
      double L = 2 * h;
      vol = kPi * r * r * L;
      I[0] = I[1] = (3 * r * r + L * L) / 12;
      I[2] = r * r / 2;
      break;
    }
    case B2MJ_GEOM_ELLIPSOID:
      vol = 4.0 / 3.0 * kPi * g.size[0] * g.size[1] * g.size[2];
      I[0] = 0.2 * (g.size[1] * g.size[1] + g.size[2] * g.size[2]);
      I[1] = 0.2 * (g.size[0] * g.size[0] + g.size[2] * g.size[2]);
      I[2] = 0.2 * (g.size[0] * g.size[0] + g.size[1] * g.size[1]);
      break;
    case B2MJ_GEOM_BOX:
      vol = 8 * g.size[0] * g.size[1] * g.size[2];
      I[0] = (g.size[1] * g.size[1] + g.size[2] * g.size[2]) / 3;
      I[1] = (g.size[0] * g.size[0] + g.size[2] * g.size[2]) / 3;
      I[2] = (g.size[0] * g.size[0] + g.size[1] * g.size[1]) / 3;
      break;
    case B2MJ_GEOM_MESH:
      vol = g.mesh_volume;
      for (int k = 0; k < 3; k++) I[k] = g.mesh_inertia[k] / vol;
      break;
    default: break;  // plane: massless
  }
  double m = explicit_mass ? *mass : density * vol;
  *mass = m;
  for (int i = 0; i < 3; i++) inertia[i] = m * I[i];
}

// Jacobi eigen-decomposition of a symmetric 3x3 (row-major A); V columns = eigenvectors
void eig3(const double* A_in, double* eval, double* V) {
  double A[9];
  std::memcpy(A, A_in, sizeof(A));
  for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1 : 0;
  for (int sweep = 0; sweep < 50; sweep++) {
    double off = std::fabs(A[1]) + std::fabs(A[2]) + std::fabs(A[5]);
    if (off < 1e-30) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double apq = A[3 * p + q];
        if (std::fabs(apq) < 1e-300) continue;
        double theta = (A[3 * q + q] - A[3 * p + p]) / (2 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double cs = 1 / std::sqrt(t * t + 1), sn = t * cs;
        for (int k = 0; k < 3; k++) {  // A <- A*J
          double akp = A[3 * k + p], akq = A[3 * k + q];
          A[3 * k + p] = cs * akp - sn * akq;
          A[3 * k + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- J'*A
          double apk = A[3 * p + k], aqk = A[3 * q + k];
          A[3 * p + k] = cs * apk - sn * aqk;
          A[3 * q + k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = cs * vkp - sn * vkq;
          V[3 * k + q] = sn * vkp + cs * vkq;
        }
      }
  }
  eval[0] = A[0]; eval[1] = A[4]; eval[2] = A[8];
  // make right-handed
  double c0[3] = {V[0], V[3], V[6]}, c1[3] = {V[1], V[4], V[7]}, c2[3] = {V[2], V[5], V[8]}, cr[3];
  cross(cr, c0, c1);
  if (dot3(cr, c2) < 0) { V[2] = -V[2]; V[5] = -V[5]; V[8] = -V[8]; }
}

// MuJoCo validates every element against its schema: an attribute it does not know is an error, not a no-op.  Same here
// for the elements that carry physics (MuJoCo 2.3.7 attribute lists); a misspelt or newer-version attribute would
// otherwise be dropped silently and the model would simulate something else.
void check_attributes(const XmlNode* n, const char* const* known) {
  for (auto& kv : n->attrs) {
    bool ok = false;
    for (const char* const* k = known; *k && !ok; k++) ok = kv.first == *k;
    if (!ok) fail(n, "unknown attribute '" + kv.first + "'");
  }
}
const char* const kGeomAttrs[] = {"name", "class", "type", "contype", "conaffinity", "condim", "group", "priority", "size",
    "material", "rgba", "friction", "mass", "density", "shellinertia", "solmix", "solref", "solimp", "margin", "gap", "fromto",
    "pos", "quat", "axisangle", "xyaxes", "zaxis", "euler", "hfield", "mesh", "fitscale", "user", "fluidshape", "fluidcoef", nullptr};
const char* const kSiteAttrs[] = {"name", "class", "type", "group", "pos", "quat", "axisangle", "xyaxes", "zaxis", "euler",
    "material", "size", "fromto", "rgba", "user", nullptr};
const char* const kJointAttrs[] = {"name", "class", "type", "group", "pos", "axis", "springdamper", "limited", "actuatorfrclimited",
    "solreflimit", "solimplimit", "solreffriction", "solimpfriction", "stiffness", "range", "actuatorfrcrange", "margin", "ref",
    "springref", "armature", "damping", "frictionloss", "user", nullptr};
const char* const kFreeJointAttrs[] = {"name", "group", nullptr};
const char* const kBodyAttrs[] = {"name", "childclass", "mocap", "pos", "quat", "axisangle", "xyaxes", "zaxis", "euler", "gravcomp",
    "user", nullptr};
const char* const kInertialAttrs[] = {"pos", "quat", "axisangle", "xyaxes", "zaxis", "euler", "mass", "diaginertia", "fullinertia", nullptr};
const char* const kFrameAttrs[] = {"name", "childclass", "pos", "quat", "axisangle", "xyaxes", "zaxis", "euler", nullptr};

// <frame pos=.. quat=.. childclass=..> (MuJoCo 3 files): a pure coordinate transform applied to everything it
// contains; it leaves no trace in the model.  F = pose of the enclosing frames relative to the body.
struct Frame {
  double pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};
  bool identity = true;
};
struct PendingBody {
  const XmlNode* node;
  Frame frame;
  std::string childclass;
};

struct Builder {
  Ctx ctx;
  std::vector<CBody> bodies;
  std::map<std::string, int> body_id, joint_id, geom_id, site_id, tendon_id;

  void read_defaults(const XmlNode* dn, const std::string& cls, const std::string& parent) {
    auto& mine = ctx.defaults[cls];
    if (!parent.empty()) {
      auto pit = ctx.defaults.find(parent);
      if (pit != ctx.defaults.end())
        for (auto& g : pit->second)
          for (auto& kv : g.second)
            if (!mine[g.first].count(kv.first)) mine[g.first][kv.first] = kv.second;
    }
    for (auto& ch : dn->children) {
      if (ch->tag == "default") continue;
      auto& grp = mine[group_of(ch->tag)];
      for (auto& kv : ch->attrs) grp[kv.first] = kv.second;
    }
    for (auto& ch : dn->children) {
      if (ch->tag != "default") continue;
      auto* cn = ch->attr("class");
      if (!cn) fail(ch.get(), "nested <default> needs a class");
      // child starts from a copy of this class
      ctx.defaults[*cn] = ctx.defaults[cls];
      read_defaults(ch.get(), *cn, "");
    }
  }

  void read_joint(const XmlNode* n, const std::string& childclass, CBody& b) {
    check_attributes(n, n->tag == "freejoint" ? kFreeJointAttrs : kJointAttrs);
    CJoint j;
    AttrMap em = effective(ctx, n, childclass);
    A a{em, n};
    j.name = a.str("name");
    if (n->tag == "freejoint") {
      j.type = B2MJ_JNT_FREE;
      b.joints.push_back(j);
      return;
    }
    std::string t = a.str("type", "hinge");
    if (t == "free") j.type = B2MJ_JNT_FREE;
    else if (t == "ball") j.type = B2MJ_JNT_BALL;
    else if (t == "slide") j.type = B2MJ_JNT_SLIDE;
    else if (t == "hinge") j.type = B2MJ_JNT_HINGE;
    else fail(n, "unknown joint type '" + t + "'");
    a.vec("pos", j.pos, 3, 3);
    if (a.vec("axis", j.axis, 3, 3)) {
      if (normalize3(j.axis) < 1e-14) fail(n, "zero joint axis");
    }
    if (j.type == B2MJ_JNT_BALL || j.type == B2MJ_JNT_FREE) { j.axis[0] = 0; j.axis[1] = 0; j.axis[2] = 1; }
    bool has_range = a.vec("range", j.range, 2, 2) == 2;
    int lim = a.tristate("limited", -1);
    if (lim == -1) lim = (ctx.autolimits && has_range) ? 1 : 0;
    j.limited = lim;
    bool ang = (j.type == B2MJ_JNT_HINGE || j.type == B2MJ_JNT_BALL) && ctx.degrees;
    if (ang) { j.range[0] *= kPi / 180; j.range[1] *= kPi / 180; }
    j.margin = a.num("margin", 0);
    if (a.has("actuatorfrcrange") || (a.has("actuatorfrclimited") && a.str("actuatorfrclimited") == "true"))
      fail(n, "joint actuatorfrcrange (clamp on the total actuator force of a joint) is not supported");
    if (a.has("springdamper")) a.vec("springdamper", j.springdamper, 2, 2);
    j.ref = a.num("ref", 0);
    j.springref = a.num("springref", 0);
    if (j.type == B2MJ_JNT_HINGE && ctx.degrees) { j.ref *= kPi / 180; j.springref *= kPi / 180; }
    j.stiffness = a.num("stiffness", 0);
    j.damping = a.num("damping", 0);
    j.armature = a.num("armature", 0);
    j.frictionloss = a.num("frictionloss", 0);
    a.vec("solreflimit", j.solref_lim, 2, 2);
    a.vec("solimplimit", j.solimp_lim, 5, 3);
    a.vec("solreffriction", j.solref_fri, 2, 2);
    a.vec("solimpfriction", j.solimp_fri, 5, 3);
    if (j.type == B2MJ_JNT_FREE && b.parent != 0) fail(n, "free joint only allowed on children of the world");
    b.joints.push_back(j);
  }

  void read_geom(const XmlNode* n, const std::string& childclass, CBody& b) {
    check_attributes(n, kGeomAttrs);
    CGeom g;
    AttrMap em = effective(ctx, n, childclass);
    A a{em, n};
    g.name = a.str("name");
    g.type = geom_type_from(n, a.str("type", a.has("mesh") ? "mesh" : a.has("hfield") ? "hfield" : "sphere"));
    int ns = a.vec("size", g.size, 3);
    a.vec("pos", g.pos, 3, 3);
    orientation(ctx, a, g.quat);
    if (g.type != B2MJ_GEOM_MESH && a.has("mesh"))  // MuJoCo then fits the primitive to the mesh and ignores size
      fail(n, "fitting a primitive geom to a mesh (type other than mesh together with a mesh attribute) is not supported");
    if (g.type == B2MJ_GEOM_MESH) {
      // the mesh frame (centre of mass + principal axes) is folded into the geom pose, as the MuJoCo compiler does
      if (!a.has("mesh")) fail(n, "mesh geom needs a mesh attribute");
      auto it = ctx.mesh_id.find(a.str("mesh"));
      if (it == ctx.mesh_id.end()) fail(n, "unknown mesh '" + a.str("mesh") + "'");
      const MeshData& md = ctx.meshes[it->second];
      g.meshid = it->second;
      double off[3], q[4];
      rotvecquat(off, md.pos, g.quat);
      for (int k = 0; k < 3; k++) g.pos[k] += off[k];
      mulquat(q, g.quat, md.quat);
      std::copy(q, q + 4, g.quat);
      std::copy(md.aabb, md.aabb + 3, g.size);
      g.mesh_volume = md.volume;
      std::copy(md.inertia, md.inertia + 3, g.mesh_inertia);
      g.rbound = md.rbound;
      ns = 3;
    }
    if (g.type == B2MJ_GEOM_HFIELD) {
      // mjCGeom::Compile: the geom takes its size from the hfield asset (x, y radii; z = elevation / 4 + base / 2, a
      // bounding value only: the collision code reads hfield_size); rbound covers the whole slab
      if (!a.has("hfield")) fail(n, "hfield geom needs an hfield attribute");
      auto it = ctx.hfield_id.find(a.str("hfield"));
      if (it == ctx.hfield_id.end()) fail(n, "unknown hfield '" + a.str("hfield") + "'");
      const Ctx::HField& hf = ctx.hfields[it->second];
      g.meshid = it->second;
      g.size[0] = hf.size[0];
      g.size[1] = hf.size[1];
      g.size[2] = 0.25 * hf.size[2] + 0.5 * hf.size[3];
      const double zmax = std::max(hf.size[2], hf.size[3]);
      g.rbound = std::sqrt(hf.size[0] * hf.size[0] + hf.size[1] * hf.size[1] + zmax * zmax);
      ns = 3;
    }
    if (a.has("fromto")) {
      double ft[6];
      a.vec("fromto", ft, 6, 6);
      if (g.type != B2MJ_GEOM_CAPSULE && g.type != B2MJ_GEOM_CYLINDER && g.type != B2MJ_GEOM_BOX &&
          g.type != B2MJ_GEOM_ELLIPSOID)
        fail(n, "fromto needs capsule/cylinder/box/ellipsoid");
      double v[3] = {ft[0] - ft[3], ft[1] - ft[4], ft[2] - ft[5]};
      double len = norm3(v);
      if (len < 1e-14) fail(n, "fromto points coincide");
      for (int i = 0; i < 3; i++) g.pos[i] = 0.5 * (ft[i] + ft[i + 3]);
      z2quat(g.quat, v);
      if (g.type == B2MJ_GEOM_BOX || g.type == B2MJ_GEOM_ELLIPSOID) { g.size[1] = g.size[0]; g.size[2] = len / 2; }
      else g.size[1] = len / 2;
    } else {
      int need = (g.type == B2MJ_GEOM_SPHERE) ? 1 : (g.type == B2MJ_GEOM_CAPSULE || g.type == B2MJ_GEOM_CYLINDER) ? 2
                 : (g.type == B2MJ_GEOM_PLANE || g.type == B2MJ_GEOM_MESH || g.type == B2MJ_GEOM_HFIELD) ? 0 : 3;
      if (ns < need) fail(n, "geom size needs " + std::to_string(need) + " values");
    }
    int nf = a.vec("friction", g.friction, 3);
    (void)nf;
    a.vec("solref", g.solref, 2, 2);
    a.vec("solimp", g.solimp, 5, 3);
    // physics-relevant geom attributes this compiler does not implement are refused, not ignored
    if (a.has("fluidshape") && a.str("fluidshape") != "none")
      fail(n, "geom fluidshape='" + a.str("fluidshape") + "' (ellipsoid fluid model) is not supported: only the inertia-box model is");
    if (a.has("fluidcoef")) fail(n, "geom fluidcoef (ellipsoid fluid model) is not supported");
    if (a.has("shellinertia") && a.str("shellinertia") == "true") fail(n, "geom shellinertia is not supported");
    g.solmix = a.num("solmix", 1);
    g.margin = a.num("margin", 0);
    g.gap = a.num("gap", 0);
    a.vec("rgba", g.rgba, 4, 1);  // only alpha matters to the physics: mj_ray skips fully transparent geoms
    g.contype = a.integer("contype", 1);
    g.group = a.integer("group", 0);
    g.conaffinity = a.integer("conaffinity", 1);
    g.condim = a.integer("condim", 3);
    g.priority = a.integer("priority", 0);
    if (g.condim != 1 && g.condim != 3 && g.condim != 4 && g.condim != 6) fail(n, "condim must be 1,3,4 or 6");
    bool explicit_mass = a.has("mass");
    g.mass = explicit_mass ? a.num("mass", 0) : 0;
    geom_mass_props(g, a.num("density", 1000), explicit_mass, &g.mass, g.inertia);
    b.geoms.push_back(g);
  }

  void read_site(const XmlNode* n, const std::string& childclass, CBody& b) {
    check_attributes(n, kSiteAttrs);
    CSite s;
    AttrMap em = effective(ctx, n, childclass);
    A a{em, n};
    s.name = a.str("name");
    s.type = geom_type_from(n, a.str("type", "sphere"));
    a.vec("size", s.size, 3);
    a.vec("pos", s.pos, 3, 3);
    orientation(ctx, a, s.quat);
    if (a.has("fromto")) {  // as for geoms: midpoint, z axis along the segment, half length into the size
      double ft[6];
      a.vec("fromto", ft, 6, 6);
      if (s.type != B2MJ_GEOM_CAPSULE && s.type != B2MJ_GEOM_CYLINDER && s.type != B2MJ_GEOM_BOX && s.type != B2MJ_GEOM_ELLIPSOID)
        fail(n, "fromto needs capsule/cylinder/box/ellipsoid");
      double v[3] = {ft[0] - ft[3], ft[1] - ft[4], ft[2] - ft[5]};
      const double len = norm3(v);
      if (len < 1e-14) fail(n, "fromto points coincide");
      for (int i = 0; i < 3; i++) s.pos[i] = 0.5 * (ft[i] + ft[i + 3]);
      z2quat(s.quat, v);
      if (s.type == B2MJ_GEOM_BOX || s.type == B2MJ_GEOM_ELLIPSOID) { s.size[1] = s.size[0]; s.size[2] = len / 2; }
      else s.size[1] = len / 2;
    }
    b.sites.push_back(s);
  }

  static void frame_apply(const Frame& F, double* pos, double* quat) {
    if (F.identity) return;
    double p[3], q[4];
    rotvecquat(p, pos, F.quat);
    for (int k = 0; k < 3; k++) pos[k] = F.pos[k] + p[k];
    if (quat) {
      mulquat(q, F.quat, quat);
      std::copy(q, q + 4, quat);
    }
  }

  // the non-body children of a body or frame node, in document order; child bodies are queued with their frame
  void read_children(const XmlNode* n, int id, const std::string& childclass, const Frame& F, std::vector<PendingBody>& queue) {
    for (auto& ch : n->children) {
      const std::string& t = ch->tag;
      if (t == "body") {
        queue.push_back({ch.get(), F, childclass});
      } else if (t == "frame") {
        check_attributes(ch.get(), kFrameAttrs);
        AttrMap em;
        for (auto& kv : ch->attrs) em[kv.first] = kv.second;
        A a{em, ch.get()};
        Frame G;
        a.vec("pos", G.pos, 3, 3);
        orientation(ctx, a, G.quat);
        frame_apply(F, G.pos, G.quat);
        G.identity = false;
        read_children(ch.get(), id, a.has("childclass") ? a.str("childclass") : childclass, G, queue);
      } else if (t == "joint" || t == "freejoint") {
        if (id == 0) fail(ch.get(), "joints are not allowed in the world body");
        read_joint(ch.get(), childclass, bodies[id]);
        CJoint& j = bodies[id].joints.back();
        if (!F.identity) {
          double ax[3];
          frame_apply(F, j.pos, nullptr);
          rotvecquat(ax, j.axis, F.quat);
          std::copy(ax, ax + 3, j.axis);
        }
      } else if (t == "geom") {
        read_geom(ch.get(), childclass, bodies[id]);
        frame_apply(F, bodies[id].geoms.back().pos, bodies[id].geoms.back().quat);
      } else if (t == "site") {
        read_site(ch.get(), childclass, bodies[id]);
        frame_apply(F, bodies[id].sites.back().pos, bodies[id].sites.back().quat);
      } else if (t == "inertial") {
        if (!F.identity) fail(ch.get(), "inertial is not allowed inside a frame");
        check_attributes(ch.get(), kInertialAttrs);
        CBody& b = bodies[id];
        AttrMap em;
        for (auto& kv : ch->attrs) em[kv.first] = kv.second;
        A a{em, ch.get()};
        b.has_inertial = true;
        a.vec("pos", b.ipos, 3, 3);
        orientation(ctx, a, b.iquat);
        b.mass = a.num("mass", 0);
        if (a.has("diaginertia")) a.vec("diaginertia", b.inertia, 3, 3);
        else if (a.has("fullinertia")) {
          double f[6];
          a.vec("fullinertia", f, 6, 6);
          double Am[9] = {f[0], f[3], f[4], f[3], f[1], f[5], f[4], f[5], f[2]}, V[9];
          eig3(Am, b.inertia, V);
          double q[4];
          mat2quat(q, V);
          mulquat(b.iquat, b.iquat, q);
        } else fail(ch.get(), "inertial needs diaginertia or fullinertia");
      } else if (t == "light" || t == "camera") {
        // rendering only: no effect on the physics
      } else if (t == "composite") {
        fail(ch.get(), "composite bodies are not supported (expand them in the model file)");
      } else if (t == "replicate") {
        fail(ch.get(), "replicate is not supported (expand the copies in the model file)");
      } else {
        fail(ch.get(), "unsupported element inside body");
      }
    }
  }

  void read_body(const XmlNode* n, int parent, std::string childclass, const Frame& F = Frame()) {
    int id = (int)bodies.size();
    if (n->tag == "worldbody") {
      id = 0;
    } else {
      check_attributes(n, kBodyAttrs);
      bodies.emplace_back();
      CBody& b = bodies.back();
      b.parent = parent;
      AttrMap em;
      for (auto& kv : n->attrs) em[kv.first] = kv.second;
      A a{em, n};
      b.name = a.str("name");
      a.vec("pos", b.pos, 3, 3);
      orientation(ctx, a, b.quat);
      frame_apply(F, b.pos, b.quat);
      b.mocap = a.str("mocap", "false") == "true";
      b.gravcomp = a.num("gravcomp", 0);
      if (a.has("childclass")) childclass = a.str("childclass");
    }
    std::vector<PendingBody> queue;
    read_children(n, id, childclass, Frame(), queue);
    if (id != 0 && bodies[id].mocap && !bodies[id].joints.empty()) fail(n, "mocap body cannot have joints");
    for (auto& pb : queue) read_body(pb.node, id, pb.childclass, pb.frame);
  }

  void body_inertia_from_geoms(CBody& b) {
    if (b.has_inertial && ctx.inertiafromgeom != 1) return;
    if (ctx.inertiafromgeom == 0) return;
    std::vector<const CGeom*> gs;
    for (auto& g : b.geoms)
      if (g.mass > 0 && g.group >= ctx.inertiagroup[0] && g.group <= ctx.inertiagroup[1]) gs.push_back(&g);
    if (gs.empty()) return;
    if (gs.size() == 1) {
      copy3(b.ipos, gs[0]->pos);
      copy4(b.iquat, gs[0]->quat);
      b.mass = gs[0]->mass;
      copy3(b.inertia, gs[0]->inertia);
      return;
    }
    double M = 0, com[3] = {0, 0, 0};
    for (auto* g : gs) {
      M += g->mass;
      for (int i = 0; i < 3; i++) com[i] += g->mass * g->pos[i];
    }
    for (int i = 0; i < 3; i++) com[i] /= M;
    double I[9] = {0};
    for (auto* g : gs) {
      double R[9];
      quat2mat(R, g->quat);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          double s = 0;
          for (int k = 0; k < 3; k++) s += R[3 * r + k] * g->inertia[k] * R[3 * c + k];
          I[3 * r + c] += s;
        }
      double d[3] = {g->pos[0] - com[0], g->pos[1] - com[1], g->pos[2] - com[2]};
      double d2 = dot3(d, d);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) I[3 * r + c] += g->mass * ((r == c ? d2 : 0) - d[r] * d[c]);
    }
    double V[9];
    eig3(I, b.inertia, V);
    mat2quat(b.iquat, V);
    copy3(b.ipos, com);
    b.mass = M;
  }
};

struct Names {
  std::string buf;
  int add(const std::string& s) {
    int adr = (int)buf.size();
    buf += s;
    buf.push_back('\0');
    return adr;
  }
};

int lookup(const std::map<std::string, int>& mp, const XmlNode* n, const std::string& kind, const std::string& name) {
  auto it = mp.find(name);
  if (it == mp.end()) fail(n, "unknown " + kind + " '" + name + "'");
  return it->second;
}

struct SensorSpec {
  const char* tag;
  int type, dim, stage, datatype;
  const char* objattr;  // attribute naming the object ("site","joint",...,"objname" for frame sensors)
  int objtype;
};

const SensorSpec kSensors[] = {
    {"touch", B2MJ_SENS_TOUCH, 1, B2MJ_STAGE_ACC, B2MJ_DATATYPE_POSITIVE, "site", B2MJ_OBJ_SITE},
    {"accelerometer", B2MJ_SENS_ACCELEROMETER, 3, B2MJ_STAGE_ACC, 0, "site", B2MJ_OBJ_SITE},
    {"velocimeter", B2MJ_SENS_VELOCIMETER, 3, B2MJ_STAGE_VEL, 0, "site", B2MJ_OBJ_SITE},
    {"gyro", B2MJ_SENS_GYRO, 3, B2MJ_STAGE_VEL, 0, "site", B2MJ_OBJ_SITE},
    {"force", B2MJ_SENS_FORCE, 3, B2MJ_STAGE_ACC, 0, "site", B2MJ_OBJ_SITE},
    {"torque", B2MJ_SENS_TORQUE, 3, B2MJ_STAGE_ACC, 0, "site", B2MJ_OBJ_SITE},
    {"magnetometer", B2MJ_SENS_MAGNETOMETER, 3, B2MJ_STAGE_POS, 0, "site", B2MJ_OBJ_SITE},
    {"rangefinder", B2MJ_SENS_RANGEFINDER, 1, B2MJ_STAGE_POS, B2MJ_DATATYPE_POSITIVE, "site", B2MJ_OBJ_SITE},
    {"jointpos", B2MJ_SENS_JOINTPOS, 1, B2MJ_STAGE_POS, 0, "joint", B2MJ_OBJ_JOINT},
    {"jointvel", B2MJ_SENS_JOINTVEL, 1, B2MJ_STAGE_VEL, 0, "joint", B2MJ_OBJ_JOINT},
    {"tendonpos", B2MJ_SENS_TENDONPOS, 1, B2MJ_STAGE_POS, 0, "tendon", B2MJ_OBJ_TENDON},
    {"tendonvel", B2MJ_SENS_TENDONVEL, 1, B2MJ_STAGE_VEL, 0, "tendon", B2MJ_OBJ_TENDON},
    {"actuatorpos", B2MJ_SENS_ACTUATORPOS, 1, B2MJ_STAGE_POS, 0, "actuator", B2MJ_OBJ_ACTUATOR},
    {"actuatorvel", B2MJ_SENS_ACTUATORVEL, 1, B2MJ_STAGE_VEL, 0, "actuator", B2MJ_OBJ_ACTUATOR},
    {"actuatorfrc", B2MJ_SENS_ACTUATORFRC, 1, B2MJ_STAGE_ACC, 0, "actuator", B2MJ_OBJ_ACTUATOR},
    {"jointactuatorfrc", B2MJ_SENS_JOINTACTFRC, 1, B2MJ_STAGE_ACC, 0, "joint", B2MJ_OBJ_JOINT},
    {"ballquat", B2MJ_SENS_BALLQUAT, 4, B2MJ_STAGE_POS, B2MJ_DATATYPE_QUATERNION, "joint", B2MJ_OBJ_JOINT},
    {"ballangvel", B2MJ_SENS_BALLANGVEL, 3, B2MJ_STAGE_VEL, 0, "joint", B2MJ_OBJ_JOINT},
    {"jointlimitpos", B2MJ_SENS_JOINTLIMITPOS, 1, B2MJ_STAGE_POS, 0, "joint", B2MJ_OBJ_JOINT},
    {"jointlimitvel", B2MJ_SENS_JOINTLIMITVEL, 1, B2MJ_STAGE_VEL, 0, "joint", B2MJ_OBJ_JOINT},
    {"jointlimitfrc", B2MJ_SENS_JOINTLIMITFRC, 1, B2MJ_STAGE_ACC, 0, "joint", B2MJ_OBJ_JOINT},
    {"tendonlimitpos", B2MJ_SENS_TENDONLIMITPOS, 1, B2MJ_STAGE_POS, 0, "tendon", B2MJ_OBJ_TENDON},
    {"tendonlimitvel", B2MJ_SENS_TENDONLIMITVEL, 1, B2MJ_STAGE_VEL, 0, "tendon", B2MJ_OBJ_TENDON},
    {"tendonlimitfrc", B2MJ_SENS_TENDONLIMITFRC, 1, B2MJ_STAGE_ACC, 0, "tendon", B2MJ_OBJ_TENDON},
    {"framepos", B2MJ_SENS_FRAMEPOS, 3, B2MJ_STAGE_POS, 0, "objname", -1},
    {"framequat", B2MJ_SENS_FRAMEQUAT, 4, B2MJ_STAGE_POS, B2MJ_DATATYPE_QUATERNION, "objname", -1},
    {"framexaxis", B2MJ_SENS_FRAMEXAXIS, 3, B2MJ_STAGE_POS, B2MJ_DATATYPE_AXIS, "objname", -1},
    {"frameyaxis", B2MJ_SENS_FRAMEYAXIS, 3, B2MJ_STAGE_POS, B2MJ_DATATYPE_AXIS, "objname", -1},
    {"framezaxis", B2MJ_SENS_FRAMEZAXIS, 3, B2MJ_STAGE_POS, B2MJ_DATATYPE_AXIS, "objname", -1},
    {"framelinvel", B2MJ_SENS_FRAMELINVEL, 3, B2MJ_STAGE_VEL, 0, "objname", -1},
    {"frameangvel", B2MJ_SENS_FRAMEANGVEL, 3, B2MJ_STAGE_VEL, 0, "objname", -1},
    {"framelinacc", B2MJ_SENS_FRAMELINACC, 3, B2MJ_STAGE_ACC, 0, "objname", -1},
    {"frameangacc", B2MJ_SENS_FRAMEANGACC, 3, B2MJ_STAGE_ACC, 0, "objname", -1},
    {"subtreecom", B2MJ_SENS_SUBTREECOM, 3, B2MJ_STAGE_POS, 0, "body", B2MJ_OBJ_BODY},
    {"subtreelinvel", B2MJ_SENS_SUBTREELINVEL, 3, B2MJ_STAGE_VEL, 0, "body", B2MJ_OBJ_BODY},
    {"subtreeangmom", B2MJ_SENS_SUBTREEANGMOM, 3, B2MJ_STAGE_VEL, 0, "body", B2MJ_OBJ_BODY},
    {"clock", B2MJ_SENS_CLOCK, 1, B2MJ_STAGE_POS, 0, nullptr, B2MJ_OBJ_UNKNOWN},
};

// <include file="..."/>: the children of the included file's root element replace the include element, wherever it
// stands (MuJoCo's semantics); relative paths resolve against the directory of the main model file.
void expand_includes(XmlNode* n, int depth) {
  if (depth > 16) fail(n, "includes nested too deeply (cycle?)");
  for (size_t i = 0; i < n->children.size();) {
    XmlNode* ch = n->children[i].get();
    if (ch->tag != "include") {
      expand_includes(ch, depth);
      i++;
      continue;
    }
    auto* fa = ch->attr("file");
    if (!fa) fail(ch, "include needs a file attribute");
    std::string path = *fa;
    if (path.empty() || path[0] != '/') path = model_dir() + (model_dir().empty() ? "" : "/") + path;
    std::ifstream f(path, std::ios::binary);
    if (!f) fail(ch, "cannot open included file '" + path + "'");
    std::stringstream ss;
    ss << f.rdbuf();
    std::string err;
    auto inc = xml_parse(ss.str(), err);
    if (!inc) fail(ch, "included file '" + path + "': " + err);
    if (inc->tag != "mujoco" && inc->tag != "mujocoinclude") fail(ch, "included file '" + path + "' must have a <mujoco> or <mujocoinclude> root");
    expand_includes(inc.get(), depth + 1);
    std::vector<std::unique_ptr<XmlNode>> kids = std::move(inc->children);
    n->children.erase(n->children.begin() + (long)i);
    for (size_t k = 0; k < kids.size(); k++) n->children.insert(n->children.begin() + (long)(i + k), std::move(kids[k]));
    i += kids.size();
  }
}

b2mjModel* compile(const XmlNode* root) {
  if (root->tag != "mujoco") fail(root, "root element must be <mujoco>");
  Builder B;
  Ctx& ctx = B.ctx;
  b2mjModel* m = model_new();
  if (!m) throw CompileError{"out of memory"};
  struct Guard {
    b2mjModel* m;
    ~Guard() { if (m) b2mj_model_free(m); }
  } guard{m};
  int nconmax_user = -1, njmax_user = -1;
  bool override_contacts = false;  // <flag override="enable">

  // ---- pass 1: compiler / option / size / defaults (they may appear anywhere, any number of times)
  for (auto& sec : root->children) {
    AttrMap em;
    for (auto& kv : sec->attrs) em[kv.first] = kv.second;
    A a{em, sec.get()};
    if (sec->tag == "compiler") {
      if (a.has("angle")) {
        std::string s = a.str("angle");
        if (s == "radian") ctx.degrees = false;
        else if (s == "degree") ctx.degrees = true;
        else fail(sec.get(), "angle must be radian or degree");
      }
      if (a.has("coordinate") && a.str("coordinate") != "local") fail(sec.get(), "only coordinate=local is supported");
      if (a.has("eulerseq")) {
        ctx.eulerseq = a.str("eulerseq");
        if (ctx.eulerseq.size() != 3) fail(sec.get(), "eulerseq needs 3 characters");
      }
      if (a.has("autolimits")) ctx.autolimits = a.str("autolimits") == "true";
      if (a.has("meshdir")) ctx.meshdir = a.str("meshdir");
      if (a.has("inertiafromgeom")) {
        std::string s = a.str("inertiafromgeom");
        ctx.inertiafromgeom = s == "true" ? 1 : s == "false" ? 0 : 2;
      }
      ctx.boundmass = a.num("boundmass", ctx.boundmass);
      ctx.boundinertia = a.num("boundinertia", ctx.boundinertia);
      ctx.settotalmass = a.num("settotalmass", ctx.settotalmass);
      if (a.has("inertiagrouprange")) {
        double r[2] = {0, 5};
        a.vec("inertiagrouprange", r, 2, 2);
        ctx.inertiagroup[0] = (int)std::lround(r[0]);
        ctx.inertiagroup[1] = (int)std::lround(r[1]);
      }
    } else if (sec->tag == "option") {
      b2mjOption& o = m->opt;
      o.timestep = a.num("timestep", o.timestep);
      o.impratio = a.num("impratio", o.impratio);
      o.tolerance = a.num("tolerance", o.tolerance);
      o.ls_tolerance = a.num("ls_tolerance", o.ls_tolerance);
      o.noslip_tolerance = a.num("noslip_tolerance", o.noslip_tolerance);
      o.mpr_tolerance = a.num("mpr_tolerance", o.mpr_tolerance);
      a.vec("gravity", o.gravity, 3, 3);
      a.vec("wind", o.wind, 3, 3);
      a.vec("magnetic", o.magnetic, 3, 3);
      o.density = a.num("density", o.density);
      o.viscosity = a.num("viscosity", o.viscosity);
      o.o_margin = a.num("o_margin", o.o_margin);
      a.vec("o_solref", o.o_solref, 2, 2);
      a.vec("o_solimp", o.o_solimp, 5, 3);
      o.iterations = a.integer("iterations", o.iterations);
      o.ls_iterations = a.integer("ls_iterations", o.ls_iterations);
      o.noslip_iterations = a.integer("noslip_iterations", o.noslip_iterations);
      o.mpr_iterations = a.integer("mpr_iterations", o.mpr_iterations);
      if (a.has("integrator")) {
        std::string s = a.str("integrator");
        o.integrator = s == "Euler" ? B2MJ_INT_EULER : s == "RK4" ? B2MJ_INT_RK4 : s == "implicit" ? B2MJ_INT_IMPLICIT
                       : s == "implicitfast" ? B2MJ_INT_IMPLICITFAST : -1;
        if (o.integrator < 0) fail(sec.get(), "unknown integrator");
      }
      if (a.has("collision")) {
        std::string s = a.str("collision");
        o.collision = s == "all" ? 0 : s == "predefined" ? 1 : s == "dynamic" ? 2 : -1;
        if (o.collision < 0) fail(sec.get(), "unknown collision mode");
      }
      if (a.has("cone")) {
        std::string s = a.str("cone");
        o.cone = s == "pyramidal" ? B2MJ_CONE_PYRAMIDAL : s == "elliptic" ? B2MJ_CONE_ELLIPTIC : -1;
        if (o.cone < 0) fail(sec.get(), "unknown cone");
      }
      if (a.has("jacobian")) {
        std::string s = a.str("jacobian");
        o.jacobian = s == "dense" ? 0 : s == "sparse" ? 1 : 2;
      }
      if (a.has("solver")) {
        std::string s = a.str("solver");
        o.solver = s == "PGS" ? B2MJ_SOL_PGS : s == "CG" ? B2MJ_SOL_CG : s == "Newton" ? B2MJ_SOL_NEWTON : -1;
        if (o.solver < 0) fail(sec.get(), "unknown solver");
      }
      if (auto* fl = sec->child("flag")) {
        static const std::pair<const char*, int> dis[] = {
            {"constraint", B2MJ_DSBL_CONSTRAINT}, {"equality", B2MJ_DSBL_EQUALITY},
            {"frictionloss", B2MJ_DSBL_FRICTIONLOSS}, {"limit", B2MJ_DSBL_LIMIT}, {"contact", B2MJ_DSBL_CONTACT},
            {"passive", B2MJ_DSBL_PASSIVE}, {"gravity", B2MJ_DSBL_GRAVITY}, {"clampctrl", B2MJ_DSBL_CLAMPCTRL},
            {"warmstart", B2MJ_DSBL_WARMSTART}, {"filterparent", B2MJ_DSBL_FILTERPARENT},
            {"actuation", B2MJ_DSBL_ACTUATION}, {"refsafe", B2MJ_DSBL_REFSAFE}, {"sensor", B2MJ_DSBL_SENSOR},
            {"midphase", B2MJ_DSBL_MIDPHASE}, {"eulerdamp", B2MJ_DSBL_EULERDAMP}};
        for (auto& kv : fl->attrs) {
          bool known = false;
          for (auto& d : dis)
            if (kv.first == d.first) {
              known = true;
              if (kv.second == "disable") o.disableflags |= d.second;
              else o.disableflags &= ~d.second;
            }
          if (known) continue;
          // enable flags.  energy / fwdinv only add outputs this library does not expose: accepted, no effect on the
          // state.  The others change the dynamics or the sensor values and are not implemented: refused when enabled.
          if (kv.first == "energy" || kv.first == "fwdinv") continue;
          if (kv.first == "override") {  // implemented at compile time, below: contact parameters rewritten to the o_* values
            override_contacts = kv.second == "enable";
            continue;
          }
          if (kv.first == "sensornoise" || kv.first == "multiccd" || kv.first == "island") {
            if (kv.second == "enable") fail(fl, "option flag " + kv.first + "=\"enable\" is not supported");
            continue;
          }
          fail(fl, "unknown option flag '" + kv.first + "'");
        }
      }
    } else if (sec->tag == "size") {
      nconmax_user = a.integer("nconmax", -1);
      njmax_user = a.integer("njmax", -1);
    } else if (sec->tag == "default") {
      B.read_defaults(sec.get(), "main", "");
    }
  }
  // ---- assets: meshes (textures / materials only matter to rendering and are skipped)
  for (auto& sec : root->children) {
    if (sec->tag != "asset") continue;
    for (auto& ch : sec->children) {
      if (ch->tag != "mesh") continue;
      AttrMap em = effective(ctx, ch.get(), "");
      A a{em, ch.get()};
      std::vector<double> pts;
      std::string name = a.str("name"), err;
      if (a.has("vertex")) {
        pts = parse_nums(a.str("vertex"));
        if (pts.size() % 3 || pts.size() < 12) fail(ch.get(), "mesh vertex data needs at least 4 points of 3 numbers");
      } else if (a.has("file")) {
        std::string file = a.str("file");
        if (name.empty()) {
          const size_t sl = file.find_last_of("/\\"), dot = file.find_last_of('.');
          name = file.substr(sl == std::string::npos ? 0 : sl + 1, dot == std::string::npos ? std::string::npos : dot - (sl == std::string::npos ? 0 : sl + 1));
        }
        if (file.empty() || file[0] != '/') {
          std::string base = ctx.meshdir;
          if (base.empty() || base[0] != '/') base = model_dir() + (model_dir().empty() || base.empty() ? "" : "/") + base;
          file = base + (base.empty() ? "" : "/") + file;
        }
        if (!mesh_read_file(file, pts, err)) fail(ch.get(), err);
      } else {
        fail(ch.get(), "mesh needs a file or a vertex attribute");
      }
      if (a.has("refpos") || a.has("refquat")) fail(ch.get(), "mesh refpos / refquat are not supported");
      if (name.empty()) fail(ch.get(), "mesh needs a name");
      double scale[3] = {1, 1, 1};
      a.vec("scale", scale, 3, 3);
      MeshData md;
      if (!mesh_process(pts, scale, md, err)) fail(ch.get(), "mesh '" + name + "': " + err);
      if (ctx.mesh_id.count(name)) fail(ch.get(), "repeated mesh name '" + name + "'");
      ctx.mesh_id[name] = (int)ctx.meshes.size();
      ctx.meshes.push_back(std::move(md));
    }
  }
  // ---- assets: height fields.  MuJoCo 2.3.7 takes the elevation from a file (PNG, or the custom binary format
  // (int32) nrow, (int32) ncol, (float32) data[nrow * ncol]) or leaves it zero for the program to fill; the inline
  // `elevation` attribute of later MuJoCo versions is accepted too.  PNG files go through png_gray.cpp (our own decoder;
  // the reference's vendored lodepng is out of scope).  mjCHField::Compile normalises the data to [0, 1].
  for (auto& sec : root->children) {
    if (sec->tag != "asset") continue;
    for (auto& ch : sec->children) {
      if (ch->tag != "hfield") continue;
      AttrMap em = effective(ctx, ch.get(), "");
      A a{em, ch.get()};
      Ctx::HField hf;
      std::string name = a.str("name");
      if (a.vec("size", hf.size, 4) != 4) fail(ch.get(), "hfield size needs 4 values (x radius, y radius, elevation, base)");
      for (int k = 0; k < 4; k++)
        if (!(hf.size[k] > 0)) fail(ch.get(), "hfield size values must be positive");
      if (a.has("file")) {
        std::string file = a.str("file");
        if (name.empty()) {
          const size_t sl = file.find_last_of("/\\"), dot = file.find_last_of('.');
          const size_t b0 = sl == std::string::npos ? 0 : sl + 1;
          name = file.substr(b0, dot == std::string::npos || dot < b0 ? std::string::npos : dot - b0);
        }
        const bool png = file.size() >= 4 && (file.substr(file.size() - 4) == ".png" || file.substr(file.size() - 4) == ".PNG");
        if (file.empty() || file[0] != '/') {
          std::string base = ctx.meshdir;
          if (base.empty() || base[0] != '/') base = model_dir() + (model_dir().empty() || base.empty() ? "" : "/") + base;
          file = base + (base.empty() ? "" : "/") + file;
        }
        FILE* fp = std::fopen(file.c_str(), "rb");
        if (!fp) fail(ch.get(), "cannot open hfield file '" + file + "'");
        if (png) {  // 8-bit grey image, image row 0 = the far (+y) edge of the field (mjCHField::LoadPNG flips the rows)
          std::vector<uint8_t> bytes, img;
          uint8_t chunk[65536];
          for (size_t got; (got = std::fread(chunk, 1, sizeof(chunk), fp)) > 0;) bytes.insert(bytes.end(), chunk, chunk + got);
          std::fclose(fp);
          unsigned w = 0, h = 0;
          std::string perr;
          if (!png_decode_gray8(bytes.data(), bytes.size(), img, w, h, perr)) fail(ch.get(), "hfield file '" + file + "': " + perr);
          if (w < 2 || h < 2) fail(ch.get(), "hfield file '" + file + "' needs at least 2 x 2 pixels");
          hf.nrow = (int)h;
          hf.ncol = (int)w;
          hf.data.resize((size_t)w * h);
          for (unsigned r = 0; r < h; r++)
            for (unsigned c = 0; c < w; c++) hf.data[c + (size_t)(h - 1 - r) * w] = img[c + (size_t)r * w];
        } else {
        int32_t dims[2] = {0, 0};
        bool ok = std::fread(dims, sizeof(int32_t), 2, fp) == 2 && dims[0] >= 2 && dims[1] >= 2 && (int64_t)dims[0] * dims[1] < (1 << 26);
        std::vector<float> buf;
        if (ok) {
          buf.resize((size_t)dims[0] * dims[1]);
          ok = std::fread(buf.data(), sizeof(float), buf.size(), fp) == buf.size();
        }
        std::fclose(fp);
        if (!ok) fail(ch.get(), "hfield file '" + file + "' is not (int32 nrow, int32 ncol, float32 data[nrow * ncol])");
        hf.nrow = dims[0];
        hf.ncol = dims[1];
        hf.data.assign(buf.begin(), buf.end());
        }
      } else {
        hf.nrow = a.integer("nrow", 0);
        hf.ncol = a.integer("ncol", 0);
        if (hf.nrow < 2 || hf.ncol < 2) fail(ch.get(), "hfield needs a file, or nrow >= 2 and ncol >= 2");
        hf.data.assign((size_t)hf.nrow * hf.ncol, 0.0);
        if (a.has("elevation")) {
          std::vector<double> ev = parse_nums(a.str("elevation"));
          if (ev.size() != hf.data.size()) fail(ch.get(), "hfield elevation needs nrow * ncol values");
          for (size_t k = 0; k < ev.size(); k++) hf.data[k] = (double)(float)ev[k];  // mjModel.hfield_data is float
        }
      }
      if (name.empty()) fail(ch.get(), "hfield needs a name");
      double emin = 1e10, emax = -1e10;
      for (double v : hf.data) { emin = std::min(emin, v); emax = std::max(emax, v); }
      for (double& v : hf.data) {
        v -= emin;
        if (emax - emin > 1e-15) v = (double)(float)(v / (emax - emin));
      }
      if (ctx.hfield_id.count(name)) fail(ch.get(), "repeated hfield name '" + name + "'");
      ctx.hfield_id[name] = (int)ctx.hfields.size();
      ctx.hfields.push_back(std::move(hf));
    }
  }

  // ---- pass 2: kinematic tree
  B.bodies.emplace_back();  // world
  B.bodies[0].name = "world";
  for (auto& sec : root->children)
    if (sec->tag == "worldbody") B.read_body(sec.get(), 0, "");
  for (size_t i = 1; i < B.bodies.size(); i++) B.body_inertia_from_geoms(B.bodies[i]);

  const int nbody = (int)B.bodies.size();
  int njnt = 0, nq = 0, nv = 0, ngeom = 0, nsite = 0, nmocap = 0;
  for (auto& b : B.bodies) {
    ngeom += (int)b.geoms.size();
    nsite += (int)b.sites.size();
    if (b.mocap) nmocap++;
    for (auto& j : b.joints) {
      njnt++;
      nq += j.type == B2MJ_JNT_FREE ? 7 : j.type == B2MJ_JNT_BALL ? 4 : 1;
      nv += j.type == B2MJ_JNT_FREE ? 6 : j.type == B2MJ_JNT_BALL ? 3 : 1;
    }
  }

  // ---- tendons, actuators, equality, excludes, sensors: collected as nodes, resolved after ids exist
  std::vector<const XmlNode*> tendon_nodes, act_nodes, eq_nodes, excl_nodes, sens_nodes, key_nodes, pair_nodes;
  for (auto& sec : root->children) {
    if (sec->tag == "tendon")
      for (auto& ch : sec->children) {
        if (ch->tag != "fixed" && ch->tag != "spatial") fail(ch.get(), "tendons are <fixed> or <spatial>");
        tendon_nodes.push_back(ch.get());
      }
    else if (sec->tag == "actuator")
      for (auto& ch : sec->children) act_nodes.push_back(ch.get());
    else if (sec->tag == "equality")
      for (auto& ch : sec->children) eq_nodes.push_back(ch.get());
    else if (sec->tag == "contact")
      for (auto& ch : sec->children) {
        if (ch->tag == "exclude") excl_nodes.push_back(ch.get());
        else if (ch->tag == "pair") pair_nodes.push_back(ch.get());
        else fail(ch.get(), "<contact> accepts <pair> and <exclude>");
      }
    else if (sec->tag == "sensor")
      for (auto& ch : sec->children) sens_nodes.push_back(ch.get());
    else if (sec->tag == "keyframe")
      for (auto& ch : sec->children) {
        if (ch->tag != "key") fail(ch.get(), "only <key> is allowed inside <keyframe>");
        key_nodes.push_back(ch.get());
      }
    else if (sec->tag == "compiler" || sec->tag == "option" || sec->tag == "size" || sec->tag == "default" ||
             sec->tag == "worldbody" || sec->tag == "visual" || sec->tag == "asset" || sec->tag == "statistic" ||
             sec->tag == "custom") {
    } else {
      fail(sec.get(), "unsupported top-level section");
    }
  }
  int nwrap = 0;
  for (auto* t : tendon_nodes) nwrap += (int)t->children.size();
  int nsensordata = 0;
  for (auto* s : sens_nodes) {
    const SensorSpec* sp = nullptr;
    for (auto& k : kSensors)
      if (s->tag == k.tag) sp = &k;
    if (!sp) fail(s, "unsupported sensor type");
    nsensordata += sp->dim;
  }

  m->nq = nq; m->nv = nv; m->nbody = nbody; m->njnt = njnt; m->ngeom = ngeom; m->nsite = nsite;
  m->ntendon = (int)tendon_nodes.size(); m->nwrap = nwrap; m->nu = (int)act_nodes.size();
  m->neq = (int)eq_nodes.size(); m->nexclude = (int)excl_nodes.size(); m->nsensor = (int)sens_nodes.size();
  m->nsensordata = nsensordata; m->nmocap = nmocap;
  m->nmesh = (int)ctx.meshes.size();
  m->nmeshvert = 0;
  for (auto& md : ctx.meshes) m->nmeshvert += (int)md.vert.size() / 3;
  m->nhfield = (int)ctx.hfields.size();
  m->nhfielddata = 0;
  for (auto& hf : ctx.hfields) m->nhfielddata += (int)hf.data.size();
  m->npair = (int)pair_nodes.size();
  m->nkey = (int)key_nodes.size();
  m->nkeyq = m->nkey * nq; m->nkeyv = m->nkey * nv; m->nkeyu = m->nkey * m->nu;
  m->nkeymp = m->nkey * 3 * nmocap; m->nkeymq = m->nkey * 4 * nmocap;
  m->nkeya = 0;  // set once the actuators are parsed (na)

  // nM and levels need the dof tree; names need everything: compute before allocating
  Names names;
  std::vector<int> dof_parent(nv), dof_madr(nv), body_lastdof(nbody, -1);
  {
    int d = 0, nM = 0;
    for (int i = 0; i < nbody; i++) {
      int last = i ? body_lastdof[B.bodies[i].parent] : -1;
      for (auto& j : B.bodies[i].joints) {
        int n = j.type == B2MJ_JNT_FREE ? 6 : j.type == B2MJ_JNT_BALL ? 3 : 1;
        for (int k = 0; k < n; k++) {
          dof_parent[d] = last;
          int depth = 1;
          for (int p = last; p >= 0; p = dof_parent[p]) depth++;
          dof_madr[d] = nM;
          nM += depth;
          last = d++;
        }
      }
      body_lastdof[i] = last;
    }
    m->nM = nM;
  }
  std::vector<int> level(nbody, 0);
  int nlevel = 1;
  for (int i = 1; i < nbody; i++) {
    level[i] = level[B.bodies[i].parent] + 1;
    nlevel = std::max(nlevel, level[i] + 1);
  }
  m->nlevel = nlevel;

  // pre-build names so nnames is known
  std::vector<int> nb_adr(nbody), nj_adr(njnt), ng_adr(ngeom), ns_adr(nsite), nt_adr(m->ntendon), na_adr(m->nu),
      nsn_adr(m->nsensor), ne_adr(m->neq), nk_adr(m->nkey), np_adr(m->npair);
  {
    int j = 0, g = 0, s = 0;
    for (int i = 0; i < nbody; i++) {
      nb_adr[i] = names.add(B.bodies[i].name);
      if (!B.bodies[i].name.empty()) B.body_id[B.bodies[i].name] = i;
    }
    for (int i = 0; i < nbody; i++)
      for (auto& jn : B.bodies[i].joints) {
        nj_adr[j] = names.add(jn.name);
        if (!jn.name.empty()) B.joint_id[jn.name] = j;
        j++;
      }
    for (int i = 0; i < nbody; i++)
      for (auto& gm : B.bodies[i].geoms) {
        ng_adr[g] = names.add(gm.name);
        if (!gm.name.empty()) B.geom_id[gm.name] = g;
        g++;
      }
    for (int i = 0; i < nbody; i++)
      for (auto& st : B.bodies[i].sites) {
        ns_adr[s] = names.add(st.name);
        if (!st.name.empty()) B.site_id[st.name] = s;
        s++;
      }
    for (size_t i = 0; i < tendon_nodes.size(); i++) {
      std::string nm = tendon_nodes[i]->attr("name") ? *tendon_nodes[i]->attr("name") : "";
      nt_adr[i] = names.add(nm);
      if (!nm.empty()) B.tendon_id[nm] = (int)i;
    }
    for (size_t i = 0; i < act_nodes.size(); i++)
      na_adr[i] = names.add(act_nodes[i]->attr("name") ? *act_nodes[i]->attr("name") : "");
    for (size_t i = 0; i < sens_nodes.size(); i++)
      nsn_adr[i] = names.add(sens_nodes[i]->attr("name") ? *sens_nodes[i]->attr("name") : "");
    for (size_t i = 0; i < eq_nodes.size(); i++)
      ne_adr[i] = names.add(eq_nodes[i]->attr("name") ? *eq_nodes[i]->attr("name") : "");
    for (size_t i = 0; i < key_nodes.size(); i++)
      nk_adr[i] = names.add(key_nodes[i]->attr("name") ? *key_nodes[i]->attr("name") : "");
    for (size_t i = 0; i < pair_nodes.size(); i++)
      np_adr[i] = names.add(pair_nodes[i]->attr("name") ? *pair_nodes[i]->attr("name") : "");
  }
  m->nnames = (int)names.buf.size();
  m->ntree = 0;
  model_alloc_arrays(m);
  std::memcpy(m->names, names.buf.data(), names.buf.size());
  std::copy(nb_adr.begin(), nb_adr.end(), m->name_bodyadr);
  std::copy(nj_adr.begin(), nj_adr.end(), m->name_jntadr);
  std::copy(ng_adr.begin(), ng_adr.end(), m->name_geomadr);
  std::copy(ns_adr.begin(), ns_adr.end(), m->name_siteadr);
  std::copy(nt_adr.begin(), nt_adr.end(), m->name_tendonadr);
  std::copy(na_adr.begin(), na_adr.end(), m->name_actuatoradr);
  std::copy(nsn_adr.begin(), nsn_adr.end(), m->name_sensoradr);
  std::copy(ne_adr.begin(), ne_adr.end(), m->name_eqadr);
  std::copy(nk_adr.begin(), nk_adr.end(), m->name_keyadr);
  std::copy(np_adr.begin(), np_adr.end(), m->name_pairadr);
  {
    int adr = 0;
    for (int i = 0; i < m->nmesh; i++) {
      const MeshData& md = ctx.meshes[i];
      m->mesh_vertadr[i] = adr;
      m->mesh_vertnum[i] = (int)md.vert.size() / 3;
      std::copy(md.vert.begin(), md.vert.end(), m->mesh_vert + 3 * adr);
      adr += m->mesh_vertnum[i];
    }
    adr = 0;
    for (int i = 0; i < m->nhfield; i++) {
      const Ctx::HField& hf = ctx.hfields[i];
      m->hfield_nrow[i] = hf.nrow;
      m->hfield_ncol[i] = hf.ncol;
      m->hfield_adr[i] = adr;
      std::copy(hf.size, hf.size + 4, m->hfield_size + 4 * i);
      std::copy(hf.data.begin(), hf.data.end(), m->hfield_data + adr);
      adr += (int)hf.data.size();
    }
  }

  // ---- fill body / joint / dof / geom / site arrays
  {
    int j = 0, d = 0, q = 0, g = 0, s = 0, mc = 0;
    for (int i = 0; i < nbody; i++) {
      CBody& b = B.bodies[i];
      m->body_parentid[i] = b.parent;
      m->body_rootid[i] = i == 0 ? 0 : (b.parent == 0 ? i : m->body_rootid[b.parent]);
      m->body_mocapid[i] = b.mocap ? mc++ : -1;
      // weld group: jointless bodies belong to the parent's group (collision filter)
      m->body_weldid[i] = i == 0 ? 0 : (b.joints.empty() ? m->body_weldid[b.parent] : i);
      m->body_level[i] = level[i];
      copy3(m->body_pos + 3 * i, b.pos);
      copy4(m->body_quat + 4 * i, b.quat);
      copy3(m->body_ipos + 3 * i, b.ipos);
      copy4(m->body_iquat + 4 * i, b.iquat);
      m->body_mass[i] = b.mass;
      copy3(m->body_inertia + 3 * i, b.inertia);
      m->body_gravcomp[i] = b.gravcomp;
      m->body_jntnum[i] = (int)b.joints.size();
      m->body_jntadr[i] = b.joints.empty() ? -1 : j;
      m->body_dofadr[i] = -1;
      m->body_dofnum[i] = 0;
      m->body_geomnum[i] = (int)b.geoms.size();
      m->body_geomadr[i] = b.geoms.empty() ? -1 : g;
      for (auto& jn : b.joints) {
        int nqj = jn.type == B2MJ_JNT_FREE ? 7 : jn.type == B2MJ_JNT_BALL ? 4 : 1;
        int nvj = jn.type == B2MJ_JNT_FREE ? 6 : jn.type == B2MJ_JNT_BALL ? 3 : 1;
        if (m->body_dofadr[i] < 0) m->body_dofadr[i] = d;
        m->body_dofnum[i] += nvj;
        m->jnt_type[j] = jn.type;
        m->jnt_qposadr[j] = q;
        m->jnt_dofadr[j] = d;
        m->jnt_bodyid[j] = i;
        m->jnt_limited[j] = jn.limited;
        copy3(m->jnt_pos + 3 * j, jn.pos);
        copy3(m->jnt_axis + 3 * j, jn.axis);
        m->jnt_stiffness[j] = jn.stiffness;
        m->jnt_range[2 * j] = jn.range[0];
        m->jnt_range[2 * j + 1] = jn.range[1];
        m->jnt_margin[j] = jn.margin;
        std::memcpy(m->jnt_solref + 2 * j, jn.solref_lim, sizeof(double) * 2);
        std::memcpy(m->jnt_solimp + 5 * j, jn.solimp_lim, sizeof(double) * 5);
        // qpos0 / qpos_spring
        if (jn.type == B2MJ_JNT_FREE) {
          copy3(m->qpos0 + q, b.pos);
          copy4(m->qpos0 + q + 3, b.quat);
          std::memcpy(m->qpos_spring + q, m->qpos0 + q, sizeof(double) * 7);
        } else if (jn.type == B2MJ_JNT_BALL) {
          m->qpos0[q] = 1;
          m->qpos_spring[q] = 1;
        } else {
          m->qpos0[q] = jn.ref;
          m->qpos_spring[q] = jn.springref;
        }
        for (int k = 0; k < nvj; k++) {
          m->dof_bodyid[d + k] = i;
          m->dof_jntid[d + k] = j;
          m->dof_parentid[d + k] = dof_parent[d + k];
          m->dof_Madr[d + k] = dof_madr[d + k];
          m->dof_armature[d + k] = jn.armature;
          m->dof_damping[d + k] = jn.damping;
          m->dof_frictionloss[d + k] = jn.frictionloss;
          std::memcpy(m->dof_solref + 2 * (d + k), jn.solref_fri, sizeof(double) * 2);
          std::memcpy(m->dof_solimp + 5 * (d + k), jn.solimp_fri, sizeof(double) * 5);
        }
        j++; d += nvj; q += nqj;
      }
      for (auto& gm : b.geoms) {
        m->geom_type[g] = gm.type;
        m->geom_contype[g] = gm.contype;
        m->geom_conaffinity[g] = gm.conaffinity;
        m->geom_condim[g] = gm.condim;
        m->geom_bodyid[g] = i;
        m->geom_priority[g] = gm.priority;
        m->geom_solmix[g] = gm.solmix;
        std::memcpy(m->geom_solref + 2 * g, gm.solref, sizeof(double) * 2);
        std::memcpy(m->geom_solimp + 5 * g, gm.solimp, sizeof(double) * 5);
        copy3(m->geom_size + 3 * g, gm.size);
        copy3(m->geom_pos + 3 * g, gm.pos);
        copy4(m->geom_quat + 4 * g, gm.quat);
        copy3(m->geom_friction + 3 * g, gm.friction);
        m->geom_margin[g] = gm.margin;
        m->geom_gap[g] = gm.gap;
        for (int k = 0; k < 4; k++) m->geom_rgba[4 * g + k] = gm.rgba[k];
        m->geom_dataid[g] = gm.meshid;
        double rb = 0;
        switch (gm.type) {
          case B2MJ_GEOM_SPHERE: rb = gm.size[0]; break;
          case B2MJ_GEOM_CAPSULE: rb = gm.size[0] + gm.size[1]; break;
          case B2MJ_GEOM_CYLINDER: rb = std::sqrt(gm.size[0] * gm.size[0] + gm.size[1] * gm.size[1]); break;
          case B2MJ_GEOM_ELLIPSOID: rb = std::max(gm.size[0], std::max(gm.size[1], gm.size[2])); break;
          case B2MJ_GEOM_BOX: rb = norm3(gm.size); break;
          case B2MJ_GEOM_MESH: rb = gm.rbound; break;
          case B2MJ_GEOM_HFIELD: rb = gm.rbound; break;
          default: rb = 0;
        }
        m->geom_rbound[g] = rb;
        g++;
      }
      for (auto& st : b.sites) {
        m->site_bodyid[s] = i;
        m->site_type[s] = st.type;
        copy3(m->site_size + 3 * s, st.size);
        copy3(m->site_pos + 3 * s, st.pos);
        copy4(m->site_quat + 4 * s, st.quat);
        s++;
      }
    }
  }
  // free bodies: pose lives in qpos; the body's own offset is its qpos0 (kept in body_pos for reset of mocap-like use)
  // kinematic-tree ids per dof (a tree = a child of the world with dofs below it)
  {
    std::map<int, int> root2tree;
    for (int d = 0; d < nv; d++) {
      int r = m->body_rootid[m->dof_bodyid[d]];
      if (!root2tree.count(r)) {
        int t = (int)root2tree.size();
        root2tree[r] = t;
      }
      m->dof_treeid[d] = root2tree[r];
    }
    m->ntree = (int)root2tree.size();
  }
  // level sets
  {
    std::vector<int> cnt(nlevel, 0);
    for (int i = 0; i < nbody; i++) cnt[level[i]]++;
    int adr = 0;
    for (int l = 0; l < nlevel; l++) {
      m->level_bodyadr[l] = adr;
      m->level_bodynum[l] = cnt[l];
      adr += cnt[l];
      cnt[l] = 0;
    }
    for (int i = 0; i < nbody; i++) m->level_body[m->level_bodyadr[level[i]] + cnt[level[i]]++] = i;
  }

  // ---- tendons
  {
    int w = 0;
    for (size_t t = 0; t < tendon_nodes.size(); t++) {
      const XmlNode* n = tendon_nodes[t];
      AttrMap em = effective(ctx, n, "");
      A a{em, n};
      m->tendon_adr[t] = w;
      m->tendon_num[t] = (int)n->children.size();
      double rng[2] = {0, 0};
      bool has_range = a.vec("range", rng, 2, 2) == 2;
      int lim = a.tristate("limited", -1);
      if (lim == -1) lim = (ctx.autolimits && has_range) ? 1 : 0;
      m->tendon_limited[t] = lim;
      m->tendon_range[2 * t] = rng[0];
      m->tendon_range[2 * t + 1] = rng[1];
      m->tendon_margin[t] = a.num("margin", 0);
      if (a.has("armature") && a.num("armature", 0) != 0) fail(tendon_nodes[t], "tendon armature is not supported");
      m->tendon_stiffness[t] = a.num("stiffness", 0);
      m->tendon_damping[t] = a.num("damping", 0);
      m->tendon_frictionloss[t] = a.num("frictionloss", 0);
      double sr[2] = {0.02, 1}, si[5] = {0.9, 0.95, 0.001, 0.5, 2};
      std::memcpy(m->tendon_solref_lim + 2 * t, sr, sizeof(sr));
      std::memcpy(m->tendon_solimp_lim + 5 * t, si, sizeof(si));
      std::memcpy(m->tendon_solref_fri + 2 * t, sr, sizeof(sr));
      std::memcpy(m->tendon_solimp_fri + 5 * t, si, sizeof(si));
      a.vec("solreflimit", m->tendon_solref_lim + 2 * t, 2, 2);
      a.vec("solimplimit", m->tendon_solimp_lim + 5 * t, 5, 3);
      a.vec("solreffriction", m->tendon_solref_fri + 2 * t, 2, 2);
      a.vec("solimpfriction", m->tendon_solimp_fri + 5 * t, 5, 3);
      double sl[2] = {-1, -1};
      int nsl = a.vec("springlength", sl, 2);
      if (nsl == 1) sl[1] = sl[0];
      m->tendon_lengthspring[2 * t] = sl[0];
      m->tendon_lengthspring[2 * t + 1] = sl[1];
      if (n->tag == "spatial") {
        // path through sites, with pulleys splitting it into branches (mjWRAP_SITE = 3, mjWRAP_PULLEY = 2); wrapping
        // around sphere / cylinder geoms is not supported
        int nsite_in_branch = 0;
        for (auto& ch : n->children) {
          if (ch->tag == "site") {
            auto* sn = ch->attr("site");
            if (!sn) fail(ch.get(), "missing site attribute");
            m->wrap_type[w] = 3;
            m->wrap_objid[w] = lookup(B.site_id, ch.get(), "site", *sn);
            m->wrap_prm[w] = 0;
            nsite_in_branch++;
          } else if (ch->tag == "pulley") {
            if (nsite_in_branch == 1) fail(ch.get(), "a tendon branch needs at least two sites");
            auto* dv = ch->attr("divisor");
            if (!dv) fail(ch.get(), "pulley needs a divisor");
            m->wrap_type[w] = 2;
            m->wrap_objid[w] = -1;
            m->wrap_prm[w] = parse_nums(*dv).at(0);
            if (m->wrap_prm[w] <= 0) fail(ch.get(), "pulley divisor must be positive");
            nsite_in_branch = 0;
          } else {
            fail(ch.get(), "spatial tendon paths accept <site> and <pulley> (geom wrapping is not supported)");
          }
          w++;
        }
        if (nsite_in_branch < 2) fail(n, "a spatial tendon must end with a branch of at least two sites");
        continue;
      }
      for (auto& ch : n->children) {
        if (ch->tag != "joint") fail(ch.get(), "fixed tendon accepts only <joint>");
        auto* jn = ch->attr("joint");
        if (!jn) fail(ch.get(), "missing joint attribute");
        int jid = lookup(B.joint_id, ch.get(), "joint", *jn);
        if (m->jnt_type[jid] != B2MJ_JNT_HINGE && m->jnt_type[jid] != B2MJ_JNT_SLIDE)
          fail(ch.get(), "fixed tendon joints must be hinge or slide");
        m->wrap_type[w] = 1;  // mjWRAP_JOINT
        m->wrap_objid[w] = jid;
        auto* cf = ch->attr("coef");
        m->wrap_prm[w] = cf ? parse_nums(*cf).at(0) : 1.0;
        w++;
      }
    }
  }

  // ---- actuators
  {
    int na = 0;
    for (size_t i = 0; i < act_nodes.size(); i++) {
      const XmlNode* n = act_nodes[i];
      const std::string& tag = n->tag;
      if (group_of(tag) != "actuator")
        fail(n, "unsupported actuator type '" + tag + "' (muscle and adhesion actuators are out of scope)");
      AttrMap em = effective(ctx, n, "");
      A a{em, n};
      double* gain = m->actuator_gainprm + B2MJ_NGAIN * i;
      double* bias = m->actuator_biasprm + B2MJ_NBIAS * i;
      double* dyn = m->actuator_dynprm + B2MJ_NDYN * i;
      gain[0] = 1;
      dyn[0] = 1;
      m->actuator_dyntype[i] = B2MJ_DYN_NONE;
      m->actuator_gaintype[i] = B2MJ_GAIN_FIXED;
      m->actuator_biastype[i] = B2MJ_BIAS_NONE;
      if (tag == "position") {
        double kp = a.num("kp", 1), kv = a.num("kv", 0);
        gain[0] = kp;
        m->actuator_biastype[i] = B2MJ_BIAS_AFFINE;
        bias[1] = -kp;
        bias[2] = -kv;
      } else if (tag == "velocity") {
        double kv = a.num("kv", 1);
        gain[0] = kv;
        m->actuator_biastype[i] = B2MJ_BIAS_AFFINE;
        bias[2] = -kv;
      } else if (tag == "intvelocity") {  // position servo on the integrated control: act' = ctrl, force = kp (act - q)
        double kp = a.num("kp", 1);
        gain[0] = kp;
        m->actuator_dyntype[i] = B2MJ_DYN_INTEGRATOR;
        m->actuator_biastype[i] = B2MJ_BIAS_AFFINE;
        bias[1] = -kp;
      } else if (tag == "damper") {  // force = -kv * velocity * ctrl: affine gain on the velocity, ctrl >= 0 required
        double kv = a.num("kv", 0);
        if (kv < 0) fail(n, "damping coefficient cannot be negative");
        gain[0] = 0;
        gain[2] = -kv;
        m->actuator_gaintype[i] = B2MJ_GAIN_AFFINE;
      } else if (tag == "cylinder") {  // pneumatic / hydraulic cylinder: first-order pressure filter, force = area * act + bias
        dyn[0] = a.num("timeconst", 1);
        gain[0] = a.num("area", 1);
        if (a.has("diameter")) { const double d = a.num("diameter", 0); gain[0] = M_PI / 4 * d * d; }
        a.vec("bias", bias, 3, 3);
        m->actuator_dyntype[i] = B2MJ_DYN_FILTER;
        m->actuator_biastype[i] = B2MJ_BIAS_AFFINE;
      } else if (tag == "general") {
        std::string s = a.str("dyntype", "none");
        m->actuator_dyntype[i] = s == "none" ? B2MJ_DYN_NONE : s == "integrator" ? B2MJ_DYN_INTEGRATOR
                                 : s == "filter" ? B2MJ_DYN_FILTER : -1;
        if (m->actuator_dyntype[i] < 0) fail(n, "unsupported dyntype '" + s + "'");
        if (a.has("actearly") && a.str("actearly") == "true") fail(n, "actuator actearly is not supported");
        s = a.str("gaintype", "fixed");
        m->actuator_gaintype[i] = s == "fixed" ? B2MJ_GAIN_FIXED : s == "affine" ? B2MJ_GAIN_AFFINE : -1;
        if (m->actuator_gaintype[i] < 0) fail(n, "unsupported gaintype '" + s + "'");
        s = a.str("biastype", "none");
        m->actuator_biastype[i] = s == "none" ? B2MJ_BIAS_NONE : s == "affine" ? B2MJ_BIAS_AFFINE : -1;
        if (m->actuator_biastype[i] < 0) fail(n, "unsupported biastype '" + s + "'");
        a.vec("gainprm", gain, B2MJ_NGAIN);
        a.vec("biasprm", bias, B2MJ_NBIAS);
        a.vec("dynprm", dyn, B2MJ_NDYN);
      }
      double* gear = m->actuator_gear + 6 * i;
      gear[0] = 1;
      a.vec("gear", gear, 6);
      double cr[2] = {0, 0}, fr[2] = {0, 0}, ar[2] = {0, 0};
      bool hcr = a.vec("ctrlrange", cr, 2, 2) == 2, hfr = a.vec("forcerange", fr, 2, 2) == 2,
           har = a.vec("actrange", ar, 2, 2) == 2;
      int cl = a.tristate("ctrllimited", -1), fl = a.tristate("forcelimited", -1), al = a.tristate("actlimited", -1);
      m->actuator_ctrllimited[i] = cl == -1 ? (ctx.autolimits && hcr) : cl;
      m->actuator_forcelimited[i] = fl == -1 ? (ctx.autolimits && hfr) : fl;
      m->actuator_actlimited[i] = al == -1 ? (ctx.autolimits && har) : al;
      if (tag == "damper") {
        if (!m->actuator_ctrllimited[i]) fail(n, "damper actuators need a control range (ctrllimited)");
        if (cr[0] < 0 || cr[1] < 0) fail(n, "damper control range cannot be negative");
      }
      std::memcpy(m->actuator_ctrlrange + 2 * i, cr, sizeof(cr));
      std::memcpy(m->actuator_forcerange + 2 * i, fr, sizeof(fr));
      std::memcpy(m->actuator_actrange + 2 * i, ar, sizeof(ar));
      m->actuator_trnid[2 * i + 1] = -1;
      if (a.has("joint")) {
        int jid = lookup(B.joint_id, n, "joint", a.str("joint"));
        if (m->jnt_type[jid] != B2MJ_JNT_HINGE && m->jnt_type[jid] != B2MJ_JNT_SLIDE)
          fail(n, "actuated joints must be hinge or slide");
        m->actuator_trntype[i] = B2MJ_TRN_JOINT;
        m->actuator_trnid[2 * i] = jid;
      } else if (a.has("tendon")) {
        m->actuator_trntype[i] = B2MJ_TRN_TENDON;
        m->actuator_trnid[2 * i] = lookup(B.tendon_id, n, "tendon", a.str("tendon"));
      } else {
        fail(n, "actuator needs joint= or tendon= (other transmissions are out of scope)");
      }
      m->actuator_actadr[i] = m->actuator_dyntype[i] == B2MJ_DYN_NONE ? -1 : na++;
    }
    m->na = na;
  }

  // ---- keyframes (mjModel key_*): unspecified parts default to the reference pose / zero, as in the MuJoCo compiler
  if (m->nkey) {
    m->nkeya = m->nkey * m->na;
    std::free(m->key_act);
    m->key_act = (double*)std::calloc(m->nkeya ? m->nkeya : 1, sizeof(double));
    for (int k = 0; k < m->nkey; k++) {
      const XmlNode* n = key_nodes[k];
      AttrMap em;
      for (auto& kv : n->attrs) em[kv.first] = kv.second;
      A a{em, n};
      m->key_time[k] = a.num("time", 0);
      auto fill = [&](const char* attr, double* dst, int cnt) {
        if (!a.has(attr)) return;
        auto v = parse_nums(a.str(attr));
        if ((int)v.size() != cnt)
          fail(n, std::string("keyframe attribute '") + attr + "' needs " + std::to_string(cnt) + " numbers, got " +
                      std::to_string(v.size()));
        std::copy(v.begin(), v.end(), dst);
      };
      std::copy(m->qpos0, m->qpos0 + nq, m->key_qpos + (size_t)k * nq);
      for (int b = 0; b < nbody; b++) {
        const int id = m->body_mocapid[b];
        if (id < 0) continue;
        std::copy(m->body_pos + 3 * b, m->body_pos + 3 * b + 3, m->key_mpos + ((size_t)k * nmocap + id) * 3);
        std::copy(m->body_quat + 4 * b, m->body_quat + 4 * b + 4, m->key_mquat + ((size_t)k * nmocap + id) * 4);
      }
      fill("qpos", m->key_qpos + (size_t)k * nq, nq);
      fill("qvel", m->key_qvel + (size_t)k * nv, nv);
      fill("act", m->key_act + (size_t)k * m->na, m->na);
      fill("ctrl", m->key_ctrl + (size_t)k * m->nu, m->nu);
      fill("mpos", m->key_mpos + (size_t)k * 3 * nmocap, 3 * nmocap);
      fill("mquat", m->key_mquat + (size_t)k * 4 * nmocap, 4 * nmocap);
    }
  }

  // ---- sensors
  {
    int adr = 0;
    for (size_t i = 0; i < sens_nodes.size(); i++) {
      const XmlNode* n = sens_nodes[i];
      const SensorSpec* sp = nullptr;
      for (auto& k : kSensors)
        if (n->tag == k.tag) sp = &k;
      AttrMap em;
      for (auto& kv : n->attrs) em[kv.first] = kv.second;
      A a{em, n};
      m->sensor_type[i] = sp->type;
      m->sensor_dim[i] = sp->dim;
      m->sensor_needstage[i] = sp->stage;
      m->sensor_datatype[i] = sp->datatype;
      m->sensor_adr[i] = adr;
      adr += sp->dim;
      m->sensor_cutoff[i] = a.num("cutoff", 0);
      m->sensor_noise[i] = a.num("noise", 0);
      m->sensor_reftype[i] = B2MJ_OBJ_UNKNOWN;
      m->sensor_refid[i] = -1;
      m->sensor_objtype[i] = sp->objtype < 0 ? B2MJ_OBJ_UNKNOWN : sp->objtype;
      m->sensor_objid[i] = -1;
      auto resolve = [&](const std::string& type, const std::string& name, int* otype, int* oid) {
        if (type == "body") { *otype = B2MJ_OBJ_BODY; *oid = lookup(B.body_id, n, "body", name); }
        else if (type == "xbody") { *otype = B2MJ_OBJ_XBODY; *oid = lookup(B.body_id, n, "body", name); }
        else if (type == "geom") { *otype = B2MJ_OBJ_GEOM; *oid = lookup(B.geom_id, n, "geom", name); }
        else if (type == "site") { *otype = B2MJ_OBJ_SITE; *oid = lookup(B.site_id, n, "site", name); }
        else fail(n, "unsupported objtype '" + type + "'");
      };
      if (sp->objattr == nullptr) {
      } else if (sp->objtype < 0) {
        if (!a.has("objtype") || !a.has("objname")) fail(n, "frame sensor needs objtype and objname");
        resolve(a.str("objtype"), a.str("objname"), &m->sensor_objtype[i], &m->sensor_objid[i]);
        if (a.has("reftype") && a.has("refname"))
          resolve(a.str("reftype"), a.str("refname"), &m->sensor_reftype[i], &m->sensor_refid[i]);
      } else {
        if (!a.has(sp->objattr)) fail(n, std::string("sensor needs attribute '") + sp->objattr + "'");
        std::string nm = a.str(sp->objattr);
        switch (sp->objtype) {
          case B2MJ_OBJ_SITE: m->sensor_objid[i] = lookup(B.site_id, n, "site", nm); break;
          case B2MJ_OBJ_JOINT: m->sensor_objid[i] = lookup(B.joint_id, n, "joint", nm); break;
          case B2MJ_OBJ_TENDON: m->sensor_objid[i] = lookup(B.tendon_id, n, "tendon", nm); break;
          case B2MJ_OBJ_BODY: m->sensor_objid[i] = lookup(B.body_id, n, "body", nm); break;
          case B2MJ_OBJ_ACTUATOR: {
            int id = -1;
            for (size_t k = 0; k < act_nodes.size(); k++)
              if (act_nodes[k]->attr("name") && *act_nodes[k]->attr("name") == nm) id = (int)k;
            if (id < 0) fail(n, "unknown actuator '" + nm + "'");
            m->sensor_objid[i] = id;
            break;
          }
        }
        int jt = sp->objtype == B2MJ_OBJ_JOINT ? m->jnt_type[m->sensor_objid[i]] : -1;
        if ((sp->type == B2MJ_SENS_BALLQUAT || sp->type == B2MJ_SENS_BALLANGVEL) && jt != B2MJ_JNT_BALL)
          fail(n, "ball sensor needs a ball joint");
        if ((sp->type == B2MJ_SENS_JOINTPOS || sp->type == B2MJ_SENS_JOINTVEL || sp->type == B2MJ_SENS_JOINTACTFRC ||
             sp->type == B2MJ_SENS_JOINTLIMITPOS || sp->type == B2MJ_SENS_JOINTLIMITVEL ||
             sp->type == B2MJ_SENS_JOINTLIMITFRC) && jt != B2MJ_JNT_HINGE && jt != B2MJ_JNT_SLIDE)
          fail(n, "joint sensor needs a hinge or slide joint");
      }
    }
  }

  // ---- contact excludes
  for (size_t i = 0; i < excl_nodes.size(); i++) {
    const XmlNode* n = excl_nodes[i];
    auto *b1 = n->attr("body1"), *b2 = n->attr("body2");
    if (!b1 || !b2) fail(n, "exclude needs body1 and body2");
    int i1 = lookup(B.body_id, n, "body", *b1), i2 = lookup(B.body_id, n, "body", *b2);
    if (i1 > i2) std::swap(i1, i2);
    m->exclude_signature[i] = (i1 << 16) + i2;
  }

  // ---- equality constraints (relative poses need body frames at qpos0)
  std::vector<double> xpos0, xquat0;
  model_body_poses0(m, xpos0, xquat0);
  for (size_t i = 0; i < eq_nodes.size(); i++) {
    const XmlNode* n = eq_nodes[i];
    AttrMap em = effective(ctx, n, "");
    A a{em, n};
    double* data = m->eq_data + B2MJ_NEQDATA * i;
    double sr[2] = {0.02, 1}, si[5] = {0.9, 0.95, 0.001, 0.5, 2};
    a.vec("solref", sr, 2, 2);
    a.vec("solimp", si, 5, 3);
    std::memcpy(m->eq_solref + 2 * i, sr, sizeof(sr));
    std::memcpy(m->eq_solimp + 5 * i, si, sizeof(si));
    m->eq_active[i] = a.str("active", "true") == "true";
    m->eq_obj2id[i] = -1;
    if (n->tag == "connect" || n->tag == "weld") {
      if (!a.has("body1")) fail(n, "needs body1");
      int b1 = lookup(B.body_id, n, "body", a.str("body1"));
      int b2 = a.has("body2") ? lookup(B.body_id, n, "body", a.str("body2")) : 0;
      m->eq_obj1id[i] = b1;
      m->eq_obj2id[i] = b2;
      double R1[9], R2[9];
      quat2mat(R1, &xquat0[4 * b1]);
      quat2mat(R2, &xquat0[4 * b2]);
      if (n->tag == "connect") {
        m->eq_type[i] = B2MJ_EQ_CONNECT;
        if (a.vec("anchor", data, 3, 3) != 3) fail(n, "connect needs anchor");
        // anchor on body2 (local) coinciding with body1's anchor at qpos0
        double g[3], d[3];
        mulmatvec3(g, R1, data);
        for (int c = 0; c < 3; c++) d[c] = g[c] + xpos0[3 * b1 + c] - xpos0[3 * b2 + c];
        mulmattvec3(data + 3, R2, d);
      } else {
        m->eq_type[i] = B2MJ_EQ_WELD;
        a.vec("anchor", data, 3, 3);
        double rp[7] = {0, 1, 0, 0, 0, 0, 0};
        a.vec("relpose", rp, 7, 7);
        data[10] = a.num("torquescale", 1);
        double qn = rp[3] * rp[3] + rp[4] * rp[4] + rp[5] * rp[5] + rp[6] * rp[6];
        if (qn < 1e-20) {
          // unspecified: relative pose of body2 in body1 at qpos0, anchor-consistent
          double g[3], d[3];
          mulmatvec3(g, R2, data);  // anchor (in body2) -> world offset
          for (int c = 0; c < 3; c++) d[c] = g[c] + xpos0[3 * b2 + c] - xpos0[3 * b1 + c];
          mulmattvec3(data + 3, R1, d);
          double q1n[4];
          negquat(q1n, &xquat0[4 * b1]);
          mulquat(data + 6, q1n, &xquat0[4 * b2]);
          normalize4(data + 6);
        } else {
          copy3(data + 3, rp);
          copy4(data + 6, rp + 3);
          normalize4(data + 6);
        }
      }
    } else if (n->tag == "joint") {
      m->eq_type[i] = B2MJ_EQ_JOINT;
      if (!a.has("joint1")) fail(n, "needs joint1");
      m->eq_obj1id[i] = lookup(B.joint_id, n, "joint", a.str("joint1"));
      if (a.has("joint2")) m->eq_obj2id[i] = lookup(B.joint_id, n, "joint", a.str("joint2"));
      double pc[5] = {0, 1, 0, 0, 0};
      a.vec("polycoef", pc, 5);
      std::memcpy(data, pc, sizeof(pc));
      for (int k = 0; k < 2; k++) {
        int id = k ? m->eq_obj2id[i] : m->eq_obj1id[i];
        if (id >= 0 && m->jnt_type[id] != B2MJ_JNT_HINGE && m->jnt_type[id] != B2MJ_JNT_SLIDE)
          fail(n, "joint equality needs hinge or slide joints");
      }
    } else if (n->tag == "tendon") {
      m->eq_type[i] = B2MJ_EQ_TENDON;
      if (!a.has("tendon1")) fail(n, "needs tendon1");
      m->eq_obj1id[i] = lookup(B.tendon_id, n, "tendon", a.str("tendon1"));
      if (a.has("tendon2")) m->eq_obj2id[i] = lookup(B.tendon_id, n, "tendon", a.str("tendon2"));
      double pc[5] = {0, 1, 0, 0, 0};
      a.vec("polycoef", pc, 5);
      std::memcpy(data, pc, sizeof(pc));
    } else {
      fail(n, "unsupported equality type");
    }
  }

  // ---- explicit contact pairs: parameters not given are filled from the two geoms with the rules the collision
  //      driver applies to dynamic pairs (max condim / friction / margin / gap, solmix-weighted solref / solimp)
  for (size_t i = 0; i < pair_nodes.size(); i++) {
    const XmlNode* n = pair_nodes[i];
    AttrMap em = effective(ctx, n, "");
    A a{em, n};
    if (!a.has("geom1") || !a.has("geom2")) fail(n, "pair needs geom1 and geom2");
    const int g1 = lookup(B.geom_id, n, "geom", a.str("geom1")), g2 = lookup(B.geom_id, n, "geom", a.str("geom2"));
    if (m->geom_bodyid[g1] == m->geom_bodyid[g2]) fail(n, "pair geoms belong to the same body");
    m->pair_geom1[i] = g1;
    m->pair_geom2[i] = g2;
    m->pair_dim[i] = a.integer("condim", std::max(m->geom_condim[g1], m->geom_condim[g2]));
    if (m->pair_dim[i] != 1 && m->pair_dim[i] != 3 && m->pair_dim[i] != 4 && m->pair_dim[i] != 6) fail(n, "condim must be 1,3,4 or 6");
    const double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2];
    const double mix = (s1 >= B2MJ_MINVAL && s2 >= B2MJ_MINVAL) ? s1 / (s1 + s2) : (s1 < B2MJ_MINVAL && s2 < B2MJ_MINVAL) ? 0.5 : (s1 < B2MJ_MINVAL ? 0.0 : 1.0);
    const double *ra = m->geom_solref + 2 * g1, *rb = m->geom_solref + 2 * g2;
    for (int k = 0; k < 2; k++) m->pair_solref[2 * i + k] = (ra[0] > 0 && rb[0] > 0) ? mix * ra[k] + (1 - mix) * rb[k] : std::min(ra[k], rb[k]);
    for (int k = 0; k < 5; k++) m->pair_solimp[5 * i + k] = mix * m->geom_solimp[5 * g1 + k] + (1 - mix) * m->geom_solimp[5 * g2 + k];
    double f3[3];
    for (int k = 0; k < 3; k++) f3[k] = std::max(m->geom_friction[3 * g1 + k], m->geom_friction[3 * g2 + k]);
    double f5[5] = {f3[0], f3[0], f3[1], f3[2], f3[2]};
    a.vec("friction", f5, 5);
    std::copy(f5, f5 + 5, m->pair_friction + 5 * i);
    a.vec("solref", m->pair_solref + 2 * i, 2, 2);
    a.vec("solimp", m->pair_solimp + 5 * i, 5, 3);
    m->pair_margin[i] = a.num("margin", std::max(m->geom_margin[g1], m->geom_margin[g2]));
    m->pair_gap[i] = a.num("gap", std::max(m->geom_gap[g1], m->geom_gap[g2]));
  }

  // <flag override="enable"> (mjENBL_OVERRIDE): every contact takes margin o_margin, gap 0 and the o_solref / o_solimp
  // reference and impedance.  The mixing rules reproduce a value both geoms share, so writing the override into every
  // geom and explicit pair at compile time is the same as overriding every contact at run time; enableflags keeps the
  // bit for the record.
  if (override_contacts) {
    m->opt.enableflags |= B2MJ_ENBL_OVERRIDE;
    for (int g = 0; g < m->ngeom; g++) {
      m->geom_margin[g] = m->opt.o_margin;
      m->geom_gap[g] = 0;
      std::copy(m->opt.o_solref, m->opt.o_solref + 2, m->geom_solref + 2 * g);
      std::copy(m->opt.o_solimp, m->opt.o_solimp + 5, m->geom_solimp + 5 * g);
    }
    for (int i = 0; i < m->npair; i++) {
      m->pair_margin[i] = m->opt.o_margin;
      m->pair_gap[i] = 0;
      std::copy(m->opt.o_solref, m->opt.o_solref + 2, m->pair_solref + 2 * i);
      std::copy(m->opt.o_solimp, m->opt.o_solimp + 5, m->pair_solimp + 5 * i);
    }
  }
  // compiler boundmass / boundinertia (lower bounds for every body but the world) and settotalmass (all masses and
  // inertias rescaled so that the model weighs this much): mjCModel::Compile order, before the derived constants
  for (int i = 1; i < m->nbody; i++) {
    if (ctx.boundmass > 0) m->body_mass[i] = std::max(m->body_mass[i], ctx.boundmass);
    if (ctx.boundinertia > 0)
      for (int k = 0; k < 3; k++) m->body_inertia[3 * i + k] = std::max(m->body_inertia[3 * i + k], ctx.boundinertia);
  }
  if (ctx.settotalmass > 0) {
    double total = 0;
    for (int i = 1; i < m->nbody; i++) total += m->body_mass[i];
    if (total > 0) {
      const double scale = ctx.settotalmass / total;
      for (int i = 1; i < m->nbody; i++) {
        m->body_mass[i] *= scale;
        for (int k = 0; k < 3; k++) m->body_inertia[3 * i + k] *= scale;
      }
    }
  }
  std::string err;
  if (model_set_const(m, err)) throw CompileError{err};
  // joint springdamper = (time constant tau, damping ratio zeta): stiffness and damping of the mass-spring-damper formed
  // with the joint's effective inertia at qpos0, I = ndof / sum(dof_invweight0) (the free joint's translational and
  // rotational weights differ, hence the average) -- k = I / (tau^2 zeta^2), b = 2 I / tau, the same relation as solref;
  // overrides the joint's own stiffness / damping and runs after set_const (mjCModel::AutoSpringDamper)
  {
    int j = 0;
    for (int i = 0; i < m->nbody; i++)
      for (auto& jn : B.bodies[i].joints) {
        if (jn.springdamper[0] > 0 && jn.springdamper[1] > 0) {
          const int dof = m->jnt_dofadr[j];
          const int ndof = jn.type == B2MJ_JNT_FREE ? 6 : jn.type == B2MJ_JNT_BALL ? 3 : 1;
          double w = 0;
          for (int k = 0; k < ndof; k++) w += m->dof_invweight0[dof + k];
          const double I = ndof / std::max(1e-15, w), tau = jn.springdamper[0], zeta = jn.springdamper[1];
          m->jnt_stiffness[j] = I / std::max(1e-15, tau * tau * zeta * zeta);
          for (int k = 0; k < ndof; k++) m->dof_damping[dof + k] = 2 * I / std::max(1e-15, tau);
        }
        j++;
      }
  }
  // <statistic>: values given in the file override the computed ones (meaninertia scales the solver tolerances)
  for (auto& sec : root->children) {
    if (sec->tag != "statistic") continue;
    AttrMap em = effective(ctx, sec.get(), "");
    A a{em, sec.get()};
    m->stat.meaninertia = a.num("meaninertia", m->stat.meaninertia);
    m->stat.meanmass = a.num("meanmass", m->stat.meanmass);
    m->stat.meansize = a.num("meansize", m->stat.meansize);
    m->stat.extent = a.num("extent", m->stat.extent);
    a.vec("center", m->stat.center, 3, 3);
  }
  // tendon spring length -1 => use length at qpos0
  for (int t = 0; t < m->ntendon; t++)
    for (int k = 0; k < 2; k++)
      if (m->tendon_lengthspring[2 * t + k] < 0) m->tendon_lengthspring[2 * t + k] = m->tendon_length0[t];

  model_build_collision_pairs(m);
  if (nconmax_user >= 0) m->nconmax = std::min(m->nconmax, nconmax_user);
  // constraint-row capacity
  {
    int rows = 0;
    for (int i = 0; i < m->neq; i++)
      rows += m->eq_type[i] == B2MJ_EQ_CONNECT ? 3 : m->eq_type[i] == B2MJ_EQ_WELD ? 6 : 1;
    for (int d = 0; d < nv; d++) rows += m->dof_frictionloss[d] > 0;
    for (int t = 0; t < m->ntendon; t++) rows += (m->tendon_frictionloss[t] > 0) + 2 * (m->tendon_limited[t] != 0);
    for (int j = 0; j < njnt; j++)
      if (m->jnt_limited[j]) rows += (m->jnt_type[j] == B2MJ_JNT_BALL) ? 1 : 2;
    int maxdim = 1;
    for (int g = 0; g < ngeom; g++) maxdim = std::max(maxdim, m->geom_condim[g]);
    int per = m->opt.cone == B2MJ_CONE_PYRAMIDAL ? (maxdim == 1 ? 1 : 2 * (maxdim - 1)) : maxdim;
    rows += per * m->nconmax;
    m->njmax = rows;
    if (njmax_user >= 0) m->njmax = std::min(m->njmax, njmax_user);
  }
  guard.m = nullptr;
  return m;
}

}  // namespace

// candidate geom pairs: everything MuJoCo's broadphase+filters could ever pass to the narrowphase,
// in the order its driver emits contacts (ascending body-pair signature, then geom ids).
void model_build_collision_pairs(b2mjModel* m) {
  std::free(m->collpair_geom1); std::free(m->collpair_geom2);
  std::free(m->collpair_slotadr); std::free(m->collpair_maxcon); std::free(m->collpair_pairid);
  m->collpair_geom1 = m->collpair_geom2 = m->collpair_slotadr = m->collpair_maxcon = m->collpair_pairid = nullptr;
  struct Cand { int sig, g1, g2, maxcon, pairid; };
  std::vector<Cand> cand;
  auto maxcon_of = [&](int t1, int t2) {
    if (t1 == B2MJ_GEOM_PLANE && t2 == B2MJ_GEOM_CAPSULE) return 2;
    if (t1 == B2MJ_GEOM_PLANE && (t2 == B2MJ_GEOM_CYLINDER || t2 == B2MJ_GEOM_MESH || t2 == B2MJ_GEOM_BOX)) return 4;
    if (t1 == B2MJ_GEOM_CAPSULE && (t2 == B2MJ_GEOM_CAPSULE || t2 == B2MJ_GEOM_BOX)) return 2;
    if (t1 == B2MJ_GEOM_BOX && t2 == B2MJ_GEOM_BOX) return 8;
    // one contact per penetrated prism of the sub-grid under the geom; MuJoCo stops at mjMAXCONPAIR = 50, this
    // library at the 8 contacts a pair's lane can hold (kernels/pair_con.cuh)
    if (t1 == B2MJ_GEOM_HFIELD) return 8;
    return 1;
  };
  auto push = [&](int ga, int gb, int pairid) {
    int x = ga, y = gb;
    if (m->geom_type[x] > m->geom_type[y]) std::swap(x, y);  // narrowphase table is upper-triangular in type
    const int t1 = m->geom_type[x], t2 = m->geom_type[y];
    if (t1 == B2MJ_GEOM_PLANE && t2 == B2MJ_GEOM_PLANE) return;
    if (t2 == B2MJ_GEOM_HFIELD) return;  // plane-hfield, hfield-hfield: no entry in mjCOLLISIONFUNC
    const int b1 = std::min(m->geom_bodyid[ga], m->geom_bodyid[gb]), b2 = std::max(m->geom_bodyid[ga], m->geom_bodyid[gb]);
    cand.push_back({(b1 << 16) + b2, x, y, maxcon_of(t1, t2), pairid});
  };
  // explicit pairs (opt.collision all / predefined): no contype / conaffinity, parent or exclude filtering
  std::set<std::pair<int, int>> explicit_pairs;
  if (m->opt.collision != 2)
    for (int p = 0; p < m->npair; p++) {
      push(m->pair_geom1[p], m->pair_geom2[p], p);
      explicit_pairs.insert({std::min(m->pair_geom1[p], m->pair_geom2[p]), std::max(m->pair_geom1[p], m->pair_geom2[p])});
    }
  // dynamic pairs (opt.collision all / dynamic); a geom pair that is also listed explicitly is left to its explicit entry
  const bool filterparent = !(m->opt.disableflags & B2MJ_DSBL_FILTERPARENT);
  std::set<int> excl(m->exclude_signature, m->exclude_signature + m->nexclude);
  if (m->opt.collision != 1)
    for (int b1 = 0; b1 < m->nbody; b1++)
      for (int b2 = b1 + 1; b2 < m->nbody; b2++) {
        if (!m->body_geomnum[b1] || !m->body_geomnum[b2]) continue;
        int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
        if (w1 == w2) continue;
        int p1 = m->body_weldid[m->body_parentid[w1]], p2 = m->body_weldid[m->body_parentid[w2]];
        if (filterparent && w1 != 0 && w2 != 0 && (w1 == p2 || w2 == p1)) continue;
        if (excl.count((b1 << 16) + b2)) continue;
        for (int ga = m->body_geomadr[b1]; ga < m->body_geomadr[b1] + m->body_geomnum[b1]; ga++)
          for (int gb = m->body_geomadr[b2]; gb < m->body_geomadr[b2] + m->body_geomnum[b2]; gb++) {
            bool ok = (m->geom_contype[ga] & m->geom_conaffinity[gb]) || (m->geom_contype[gb] & m->geom_conaffinity[ga]);
            if (!ok || explicit_pairs.count({std::min(ga, gb), std::max(ga, gb)})) continue;
            push(ga, gb, -1);
          }
      }
  // contacts come out in body-pair signature order; within a body pair explicit pairs first (in pair order), then the
  // dynamic geom pairs in geom order
  std::stable_sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) {
    if (a.sig != b.sig) return a.sig < b.sig;
    return (a.pairid >= 0) > (b.pairid >= 0);
  });
  m->ncollpair = (int)cand.size();
  size_t n = cand.size() ? cand.size() : 1;
  m->collpair_geom1 = (int*)std::calloc(n, sizeof(int));
  m->collpair_geom2 = (int*)std::calloc(n, sizeof(int));
  m->collpair_slotadr = (int*)std::calloc(n, sizeof(int));
  m->collpair_maxcon = (int*)std::calloc(n, sizeof(int));
  m->collpair_pairid = (int*)std::calloc(n, sizeof(int));
  int slots = 0;
  for (size_t i = 0; i < cand.size(); i++) {
    m->collpair_geom1[i] = cand[i].g1;
    m->collpair_geom2[i] = cand[i].g2;
    m->collpair_maxcon[i] = cand[i].maxcon;
    m->collpair_pairid[i] = cand[i].pairid;
    m->collpair_slotadr[i] = slots;
    slots += cand[i].maxcon;
  }
  m->nconmax = slots;
}

}  // namespace b2mj

using namespace b2mj;

extern "C" {

int b2mj_model_from_xml_string(const char* xml, b2mjModel** out) {
  if (!xml || !out) {
    set_error("b2mj_model_from_xml_string: null argument");
    return B2MJ_EINVAL;
  }
  *out = nullptr;
  std::string err;
  auto root = xml_parse(xml, err);
  if (!root) {
    set_error("XML parse error: " + err);
    return B2MJ_EPARSE;
  }
  try {
    expand_includes(root.get(), 0);
    *out = compile(root.get());
  } catch (const CompileError& e) {
    set_error(e.msg);
    return B2MJ_EPARSE;
  } catch (const std::exception& e) {
    set_error(std::string("model compile failed: ") + e.what());
    return B2MJ_EPARSE;
  }
  return B2MJ_OK;
}

int b2mj_model_from_xml_file(const char* path, b2mjModel** out) {
  if (!path || !out) {
    set_error("b2mj_model_from_xml_file: null argument");
    return B2MJ_EINVAL;
  }
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    set_error(std::string("cannot open '") + path + "'");
    return B2MJ_EINVAL;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  // relative mesh files are resolved against the directory of the model file (plus compiler meshdir)
  const std::string p(path);
  const size_t sl = p.find_last_of('/');
  set_model_dir(sl == std::string::npos ? std::string(".") : p.substr(0, sl));
  const int rc = b2mj_model_from_xml_string(ss.str().c_str(), out);
  set_model_dir("");
  return rc;
}

}  // extern "C"
