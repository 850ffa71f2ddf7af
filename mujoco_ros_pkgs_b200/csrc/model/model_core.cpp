// model_core.cpp — b2mjModel ownership, reflection, names, field table.
#include "model_core.h"

#include <cstdio>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace b2mj {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const std::string& get_error() { return g_err; }

void model_default_option(b2mjOption* o) {
  // MuJoCo 2.3.7 defaults
  std::memset(o, 0, sizeof(*o));
  o->timestep = 0.002;
  o->impratio = 1;
  o->tolerance = 1e-8;
  o->ls_tolerance = 0.01;
  o->noslip_tolerance = 1e-6;
  o->mpr_tolerance = 1e-6;
  o->gravity[2] = -9.81;
  o->magnetic[1] = -0.5;
  o->o_solref[0] = 0.02;
  o->o_solref[1] = 1;
  o->o_solimp[0] = 0.9;
  o->o_solimp[1] = 0.95;
  o->o_solimp[2] = 0.001;
  o->o_solimp[3] = 0.5;
  o->o_solimp[4] = 2;
  o->integrator = B2MJ_INT_EULER;
  o->collision = 0;
  o->cone = B2MJ_CONE_PYRAMIDAL;
  o->jacobian = 2;
  o->solver = B2MJ_SOL_NEWTON;
  o->iterations = 100;
  o->ls_iterations = 50;
  o->noslip_iterations = 0;
  o->mpr_iterations = 50;
}

b2mjModel* model_new() {
  b2mjModel* m = (b2mjModel*)std::calloc(1, sizeof(b2mjModel));
  if (m) model_default_option(&m->opt);
  return m;
}

void model_alloc_arrays(b2mjModel* m) {
#define X(t, n, r, c)                                                         \
  if (!m->n) {                                                                \
    size_t cnt = (size_t)std::max(m->r, 0) * (size_t)(c);                     \
    m->n = (t*)std::calloc(cnt ? cnt : 1, sizeof(t));                         \
  }
  B2MJ_MODEL_ARRAYS(X)
#undef X
}

b2mjModel* model_clone(const b2mjModel* src) {
  b2mjModel* m = (b2mjModel*)std::calloc(1, sizeof(b2mjModel));
  if (!m) return nullptr;
#define X(n) m->n = src->n;
  B2MJ_MODEL_SIZES(X)
#undef X
  m->opt = src->opt;
  m->stat = src->stat;
  model_alloc_arrays(m);
#define X(t, n, r, c) \
  if (src->n && m->r > 0) std::memcpy(m->n, src->n, sizeof(t) * (size_t)m->r * (size_t)(c));
  B2MJ_MODEL_ARRAYS(X)
#undef X
  return m;
}

struct FieldInfo {
  const char* name;
  int is_int;
};

static const FieldInfo kFields[B2MJ_NFIELD] = {
    {"qpos", 0}, {"qvel", 0}, {"act", 0}, {"ctrl", 0}, {"qfrc_applied", 0}, {"xfrc_applied", 0},
    {"mocap_pos", 0}, {"mocap_quat", 0}, {"qacc_warmstart", 0}, {"time", 0},
    {"qacc", 0}, {"sensordata", 0}, {"act_dot", 0},
    {"xpos", 0}, {"xquat", 0}, {"xmat", 0}, {"xipos", 0}, {"ximat", 0}, {"xanchor", 0}, {"xaxis", 0},
    {"geom_xpos", 0}, {"geom_xmat", 0}, {"site_xpos", 0}, {"site_xmat", 0}, {"subtree_com", 0},
    {"cinert", 0}, {"cdof", 0}, {"crb", 0}, {"ten_length", 0}, {"ten_J", 0}, {"actuator_length", 0},
    {"actuator_moment", 0}, {"qM", 0}, {"qLD", 0}, {"qLDiagInv", 0}, {"qLDiagSqrtInv", 0},
    {"ten_velocity", 0}, {"actuator_velocity", 0}, {"cvel", 0}, {"cdof_dot", 0}, {"qfrc_bias", 0},
    {"qfrc_passive", 0},
    {"actuator_force", 0}, {"qfrc_actuator", 0}, {"qfrc_smooth", 0}, {"qacc_smooth", 0},
    {"qfrc_constraint", 0}, {"cacc", 0}, {"cfrc_int", 0}, {"cfrc_ext", 0},
    {"contact_dist", 0}, {"contact_pos", 0}, {"contact_frame", 0}, {"contact_includemargin", 0},
    {"contact_friction", 0}, {"contact_solref", 0}, {"contact_solimp", 0}, {"contact_mu", 0},
    {"contact_dim", 1}, {"contact_geom1", 1}, {"contact_geom2", 1}, {"contact_exclude", 1},
    {"contact_efc_address", 1},
    {"efc_type", 1}, {"efc_id", 1}, {"efc_J", 0}, {"efc_pos", 0}, {"efc_margin", 0},
    {"efc_frictionloss", 0}, {"efc_diagApprox", 0}, {"efc_KBIP", 0}, {"efc_D", 0}, {"efc_R", 0},
    {"efc_vel", 0}, {"efc_aref", 0}, {"efc_b", 0}, {"efc_force", 0}, {"efc_state", 1}, {"efc_AR", 0},
    {"ncon", 1}, {"nefc", 1}, {"solver_iter", 1}, {"warning", 1},
};

template <typename T> struct Kind;
template <> struct Kind<double> { static const int v = 0; };
template <> struct Kind<int> { static const int v = 1; };
template <> struct Kind<char> { static const int v = 2; };

}  // namespace b2mj

using namespace b2mj;

extern "C" {

const char* b2mj_last_error(void) { return get_error().c_str(); }
int b2mj_version(void) { return B2MJ_VERSION; }

int b2mj_field_size(const b2mjModel* m, b2mj_field f, int* is_int) {
  if (!m || f < 0 || f >= B2MJ_NFIELD) {
    set_error("b2mj_field_size: bad argument");
    return B2MJ_EINVAL;
  }
  if (is_int) *is_int = kFields[f].is_int;
  const int nb = m->nbody, nv = m->nv, nc = m->nconmax, nj = m->njmax;
  switch (f) {
    case B2MJ_F_QPOS: return m->nq;
    case B2MJ_F_QVEL: case B2MJ_F_QFRC_APPLIED: case B2MJ_F_QACC_WARMSTART: case B2MJ_F_QACC:
    case B2MJ_F_QLDIAGINV: case B2MJ_F_QLDIAGSQRTINV: case B2MJ_F_QFRC_BIAS: case B2MJ_F_QFRC_PASSIVE:
    case B2MJ_F_QFRC_ACTUATOR: case B2MJ_F_QFRC_SMOOTH: case B2MJ_F_QACC_SMOOTH:
    case B2MJ_F_QFRC_CONSTRAINT: return nv;
    case B2MJ_F_ACT: case B2MJ_F_ACT_DOT: return m->na;
    case B2MJ_F_CTRL: case B2MJ_F_ACTUATOR_LENGTH: case B2MJ_F_ACTUATOR_VELOCITY:
    case B2MJ_F_ACTUATOR_FORCE: return m->nu;
    case B2MJ_F_XFRC_APPLIED: case B2MJ_F_CVEL: case B2MJ_F_CACC: case B2MJ_F_CFRC_INT:
    case B2MJ_F_CFRC_EXT: return 6 * nb;
    case B2MJ_F_MOCAP_POS: return 3 * m->nmocap;
    case B2MJ_F_MOCAP_QUAT: return 4 * m->nmocap;
    case B2MJ_F_TIME: return 1;
    case B2MJ_F_SENSORDATA: return m->nsensordata;
    case B2MJ_F_XPOS: case B2MJ_F_XIPOS: case B2MJ_F_SUBTREE_COM: return 3 * nb;
    case B2MJ_F_XQUAT: return 4 * nb;
    case B2MJ_F_XMAT: case B2MJ_F_XIMAT: return 9 * nb;
    case B2MJ_F_XANCHOR: case B2MJ_F_XAXIS: return 3 * m->njnt;
    case B2MJ_F_GEOM_XPOS: return 3 * m->ngeom;
    case B2MJ_F_GEOM_XMAT: return 9 * m->ngeom;
    case B2MJ_F_SITE_XPOS: return 3 * m->nsite;
    case B2MJ_F_SITE_XMAT: return 9 * m->nsite;
    case B2MJ_F_CINERT: case B2MJ_F_CRB: return 10 * nb;
    case B2MJ_F_CDOF: case B2MJ_F_CDOF_DOT: return 6 * nv;
    case B2MJ_F_TEN_LENGTH: case B2MJ_F_TEN_VELOCITY: return m->ntendon;
    case B2MJ_F_TEN_J: return m->ntendon * nv;
    case B2MJ_F_ACTUATOR_MOMENT: return m->nu * nv;
    case B2MJ_F_QM: case B2MJ_F_QLD: return m->nM;
    case B2MJ_F_CONTACT_DIST: case B2MJ_F_CONTACT_INCLUDEMARGIN: case B2MJ_F_CONTACT_MU:
    case B2MJ_F_CONTACT_DIM: case B2MJ_F_CONTACT_GEOM1: case B2MJ_F_CONTACT_GEOM2:
    case B2MJ_F_CONTACT_EXCLUDE: case B2MJ_F_CONTACT_EFC_ADDRESS: return nc;
    case B2MJ_F_CONTACT_POS: return 3 * nc;
    case B2MJ_F_CONTACT_FRAME: return 9 * nc;
    case B2MJ_F_CONTACT_FRICTION: case B2MJ_F_CONTACT_SOLIMP: return 5 * nc;
    case B2MJ_F_CONTACT_SOLREF: return 2 * nc;
    case B2MJ_F_EFC_TYPE: case B2MJ_F_EFC_ID: case B2MJ_F_EFC_POS: case B2MJ_F_EFC_MARGIN:
    case B2MJ_F_EFC_FRICTIONLOSS: case B2MJ_F_EFC_DIAGAPPROX: case B2MJ_F_EFC_D: case B2MJ_F_EFC_R:
    case B2MJ_F_EFC_VEL: case B2MJ_F_EFC_AREF: case B2MJ_F_EFC_B: case B2MJ_F_EFC_FORCE:
    case B2MJ_F_EFC_STATE: return nj;
    case B2MJ_F_EFC_J: return nj * nv;
    case B2MJ_F_EFC_KBIP: return 4 * nj;
    case B2MJ_F_EFC_AR: return (m->opt.solver == B2MJ_SOL_PGS || m->opt.noslip_iterations > 0) ? nj * nj : 0;
    case B2MJ_F_NCON: case B2MJ_F_NEFC: case B2MJ_F_SOLVER_ITER: return 1;
    case B2MJ_F_WARNING: return B2MJ_NWARNING;
    default: break;
  }
  return B2MJ_EINVAL;
}

const char* b2mj_field_name(b2mj_field f) {
  if (f < 0 || f >= B2MJ_NFIELD) return nullptr;
  return kFields[f].name;
}

int b2mj_field_by_name(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < B2MJ_NFIELD; i++)
    if (!std::strcmp(kFields[i].name, name)) return i;
  return -1;
}

void b2mj_model_free(b2mjModel* m) {
  if (!m) return;
#define X(t, n, r, c) std::free(m->n);
  B2MJ_MODEL_ARRAYS(X)
#undef X
  std::free(m);
}

// ---- binary model files (the reference loads ".mjb" through mj_loadModel, mujoco_env.cpp:771-911) ----
// layout: "B2MJB\0" | int32 version | int32 nsizes | int32 narrays | int32 sizeof(option) | int32 sizeof(statistic) |
//         nsizes x int32 | option | statistic | per array: int64 byte count, bytes
static const char kBinMagic[6] = {'B', '2', 'M', 'J', 'B', 0};

int b2mj_model_save_binary(const b2mjModel* m, const char* path) {
  if (!m || !path) { set_error("b2mj_model_save_binary: null argument"); return B2MJ_EINVAL; }
  FILE* f = std::fopen(path, "wb");
  if (!f) { set_error(std::string("cannot open '") + path + "' for writing"); return B2MJ_EINVAL; }
  bool ok = std::fwrite(kBinMagic, 1, 6, f) == 6;
  auto put32 = [&](int v) { ok = ok && std::fwrite(&v, sizeof(int), 1, f) == 1; };
  put32(B2MJ_VERSION);
  put32(b2mj_model_nsizes());
  put32(b2mj_model_narrays());
  put32((int)sizeof(b2mjOption));
  put32((int)sizeof(b2mjStatistic));
#define X(n) put32(m->n);
  B2MJ_MODEL_SIZES(X)
#undef X
  ok = ok && std::fwrite(&m->opt, sizeof(b2mjOption), 1, f) == 1;
  ok = ok && std::fwrite(&m->stat, sizeof(b2mjStatistic), 1, f) == 1;
#define X(t, n, r, c)                                                                   \
  {                                                                                     \
    const long long bytes = (long long)sizeof(t) * std::max(m->r, 0) * (long long)(c);  \
    ok = ok && std::fwrite(&bytes, sizeof(bytes), 1, f) == 1;                           \
    if (bytes) ok = ok && std::fwrite(m->n, 1, (size_t)bytes, f) == (size_t)bytes;      \
  }
  B2MJ_MODEL_ARRAYS(X)
#undef X
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) { set_error(std::string("write error on '") + path + "'"); return B2MJ_EINVAL; }
  return 0;
}

int b2mj_model_load_binary(const char* path, b2mjModel** out) {
  if (!path || !out) { set_error("b2mj_model_load_binary: null argument"); return B2MJ_EINVAL; }
  FILE* f = std::fopen(path, "rb");
  if (!f) { set_error(std::string("cannot open '") + path + "'"); return B2MJ_EINVAL; }
  auto bad = [&](const std::string& why, b2mjModel* m) {
    std::fclose(f);
    if (m) b2mj_model_free(m);
    set_error(std::string("'") + path + "': " + why);
    return B2MJ_EINVAL;
  };
  char magic[6];
  int hdr[5];
  if (std::fread(magic, 1, 6, f) != 6 || std::memcmp(magic, kBinMagic, 6)) return bad("not a b2mj binary model", nullptr);
  if (std::fread(hdr, sizeof(int), 5, f) != 5) return bad("truncated header", nullptr);
  if (hdr[0] != B2MJ_VERSION || hdr[1] != b2mj_model_nsizes() || hdr[2] != b2mj_model_narrays() ||
      hdr[3] != (int)sizeof(b2mjOption) || hdr[4] != (int)sizeof(b2mjStatistic))
    return bad("written by an incompatible library version (re-compile the model from its XML)", nullptr);
  b2mjModel* m = model_new();
  if (!m) return bad("out of memory", nullptr);
#define X(n) if (std::fread(&m->n, sizeof(int), 1, f) != 1 || m->n < 0) return bad("bad size field '" #n "'", m);
  B2MJ_MODEL_SIZES(X)
#undef X
  if (std::fread(&m->opt, sizeof(b2mjOption), 1, f) != 1 || std::fread(&m->stat, sizeof(b2mjStatistic), 1, f) != 1)
    return bad("truncated option block", m);
  model_alloc_arrays(m);
#define X(t, n, r, c)                                                                                        \
  {                                                                                                          \
    long long bytes = -1;                                                                                    \
    const long long want = (long long)sizeof(t) * m->r * (long long)(c);                                     \
    if (std::fread(&bytes, sizeof(bytes), 1, f) != 1 || bytes != want) return bad("bad array '" #n "'", m);  \
    if (bytes && std::fread(m->n, 1, (size_t)bytes, f) != (size_t)bytes) return bad("truncated array '" #n "'", m); \
  }
  B2MJ_MODEL_ARRAYS(X)
#undef X
  std::fclose(f);
  *out = m;
  return 0;
}

int b2mj_model_from_file(const char* path, b2mjModel** out) {
  if (!path || !out) { set_error("b2mj_model_from_file: null argument"); return B2MJ_EINVAL; }
  const std::string p(path);
  const std::string ext = p.size() >= 6 ? p.substr(p.size() - 6) : "";
  if (ext == ".b2mjb") return b2mj_model_load_binary(path, out);
  return b2mj_model_from_xml_file(path, out);
}

static const int* name_adr_table(const b2mjModel* m, int objtype, int* count) {
  switch (objtype) {
    case B2MJ_OBJ_BODY: case B2MJ_OBJ_XBODY: *count = m->nbody; return m->name_bodyadr;
    case B2MJ_OBJ_JOINT: *count = m->njnt; return m->name_jntadr;
    case B2MJ_OBJ_GEOM: *count = m->ngeom; return m->name_geomadr;
    case B2MJ_OBJ_SITE: *count = m->nsite; return m->name_siteadr;
    case B2MJ_OBJ_TENDON: *count = m->ntendon; return m->name_tendonadr;
    case B2MJ_OBJ_ACTUATOR: *count = m->nu; return m->name_actuatoradr;
    case B2MJ_OBJ_SENSOR: *count = m->nsensor; return m->name_sensoradr;
    case B2MJ_OBJ_EQUALITY: *count = m->neq; return m->name_eqadr;
    case B2MJ_OBJ_KEY: *count = m->nkey; return m->name_keyadr;
    case B2MJ_OBJ_PAIR: *count = m->npair; return m->name_pairadr;
    default: *count = 0; return nullptr;
  }
}

int b2mj_name2id(const b2mjModel* m, int objtype, const char* name) {
  if (!m || !name) return -1;
  int n = 0;
  const int* adr = name_adr_table(m, objtype, &n);
  if (!adr) return -1;
  for (int i = 0; i < n; i++)
    if (!std::strcmp(m->names + adr[i], name)) return i;
  return -1;
}

const char* b2mj_id2name(const b2mjModel* m, int objtype, int id) {
  if (!m) return nullptr;
  int n = 0;
  const int* adr = name_adr_table(m, objtype, &n);
  if (!adr || id < 0 || id >= n) return nullptr;
  return m->names + adr[id];
}

int b2mj_model_narrays(void) {
  int n = 0;
#define X(t, nm, r, c) n++;
  B2MJ_MODEL_ARRAYS(X)
#undef X
  return n;
}

int b2mj_model_array_info(const b2mjModel* m, int idx, const char** name, int* elem_kind, int* rows,
                          int* cols, void** ptr) {
  if (!m) return B2MJ_EINVAL;
  int i = 0;
#define X(t, nm, r, c)                        \
  if (i++ == idx) {                           \
    if (name) *name = #nm;                    \
    if (elem_kind) *elem_kind = Kind<t>::v;   \
    if (rows) *rows = m->r;                   \
    if (cols) *cols = (c);                    \
    if (ptr) *ptr = (void*)m->nm;             \
    return 0;                                 \
  }
  B2MJ_MODEL_ARRAYS(X)
#undef X
  return B2MJ_EINVAL;
}

int b2mj_model_nsizes(void) {
  int n = 0;
#define X(nm) n++;
  B2MJ_MODEL_SIZES(X)
#undef X
  return n;
}

int b2mj_model_size_info(const b2mjModel* m, int idx, const char** name, int* value) {
  if (!m) return B2MJ_EINVAL;
  int i = 0;
#define X(nm)                    \
  if (i++ == idx) {              \
    if (name) *name = #nm;       \
    if (value) *value = m->nm;   \
    return 0;                    \
  }
  B2MJ_MODEL_SIZES(X)
#undef X
  return B2MJ_EINVAL;
}

int b2mj_model_set_const(b2mjModel* m) {
  if (!m) return B2MJ_EINVAL;
  std::string err;
  int rc = model_set_const(m, err);
  if (rc) set_error(err);
  return rc;
}

}  // extern "C"
