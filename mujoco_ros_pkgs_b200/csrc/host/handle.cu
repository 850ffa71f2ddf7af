// handle.cu — the batched-environment handle behind the C-ABI (include/b2mj.h).
//
// Host side of the boundary: owns the device copy of the model, the HBM state records and arenas,
// decides the shared-memory layout / launch shape, and launches the fused step kernel.  Mirrors what
// MujocoEnv owns around its mj_step call (reference mujoco_ros/src/mujoco_env.cpp: model_/data_
// :747-748, mj_makeData :872, mj_resetData :252, mj_forward :329/:621, mj_step :498/:552/:593).
// No CPU fallback: every entry point that needs the GPU fails with B2MJ_ECUDA / B2MJ_ENODEVICE.
#include <cuda_runtime.h>
#include <cmath>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b2mj.h"
#include "handle.h"
#include "kernels/step_launch.h"
#include "model/model_core.h"

using namespace b2mj;
using namespace b2k;

namespace b2mj {

#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                          \
      return B2MJ_ECUDA;                                                                      \
    }                                                                                         \
  } while (0)

static bool is_record_field(int f) {
  switch (f) {
    case B2MJ_F_QPOS: case B2MJ_F_QVEL: case B2MJ_F_ACT: case B2MJ_F_CTRL: case B2MJ_F_QFRC_APPLIED:
    case B2MJ_F_QACC_WARMSTART: case B2MJ_F_TIME: case B2MJ_F_QACC: case B2MJ_F_SENSORDATA: case B2MJ_F_ACT_DOT:
      return true;
    default: return false;
  }
}

// ---- kernels local to the handle ----
__global__ void reset_kernel(double* rec, const double* rec_init, int pitch, int nenv, const unsigned char* mask,
                             int* warning, int* stats, double* xfrc, int nxfrc, double* mocap, const double* mocap_init,
                             int nmocap7) {
  const int env = blockIdx.x;
  if (env >= nenv || (mask && !mask[env])) return;
  for (int k = threadIdx.x; k < pitch; k += blockDim.x) rec[(size_t)env * pitch + k] = rec_init[k];
  for (int k = threadIdx.x; k < B2MJ_NWARNING; k += blockDim.x) warning[(size_t)env * B2MJ_NWARNING + k] = 0;
  for (int k = threadIdx.x; k < 4; k += blockDim.x) stats[(size_t)env * 4 + k] = 0;
  if (xfrc) for (int k = threadIdx.x; k < nxfrc; k += blockDim.x) xfrc[(size_t)env * nxfrc + k] = 0;
  if (mocap) for (int k = threadIdx.x; k < nmocap7; k += blockDim.x) mocap[(size_t)env * nmocap7 + k] = mocap_init[k];
}

// device-to-device field write: dst[env*dpitch + k] = src[env*spitch + k] (32-bit words)
__global__ void scatter_field_kernel(unsigned* dst, int dpitch, const unsigned* src, int spitch, int words, int nenv) {
  const long long total = (long long)nenv * words;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int env = (int)(i / words), k = (int)(i - (long long)env * words);
    dst[(size_t)env * dpitch + k] = src[(size_t)env * spitch + k];
  }
}

static int field_count(const b2mjModel* m, int f, int* is_int) {
  if (f == B2MJ_F_EFC_AR) {  // the GPU solver is matrix-free: AR is never materialised
    if (is_int) *is_int = 0;
    return 0;
  }
  return b2mj_field_size(m, (b2mj_field)f, is_int);
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// bytes of the per-variant region of a model blob: every B2MJ_MODEL_ARRAYS array + gravity[3] + {meaninertia}
static size_t variant_region_bytes(const b2mjModel* m) {
  size_t total = 0;
#define X(t, n, r, c) total += al256(sizeof(t) * (size_t)std::max(m->r, 0) * (size_t)(c) + 16);
  B2MJ_MODEL_ARRAYS(X)
#undef X
  total += al256(3 * sizeof(double) + 16) + al256(sizeof(double) + 16);
  return total;
}
// pack one model's arrays at host (same order / offsets for every variant) and point d's arrays at dev_base
static void pack_variant(const b2mjModel* m, unsigned char* host, DevModel* d, unsigned char* dev_base) {
  size_t off = 0;
#define X(t, n, r, c)                                                               \
  {                                                                                 \
    size_t bytes = sizeof(t) * (size_t)std::max(m->r, 0) * (size_t)(c);             \
    if (bytes) std::memcpy(host + off, m->n, bytes);                                \
    if (d) d->n = reinterpret_cast<const t*>(dev_base + off);                       \
    off += al256(bytes + 16);                                                       \
  }
  B2MJ_MODEL_ARRAYS(X)
#undef X
  std::memcpy(host + off, m->opt.gravity, 3 * sizeof(double));
  if (d) d->env_gravity = reinterpret_cast<const double*>(dev_base + off);
  off += al256(3 * sizeof(double) + 16);
  std::memcpy(host + off, &m->stat.meaninertia, sizeof(double));
  if (d) d->env_scalars = reinterpret_cast<const double*>(dev_base + off);
}

// upload model arrays into one device blob and fill the DevModel pointers
static int upload_model(Handle* h) {
  const b2mjModel* m = h->model;
  DevModel& d = h->dm;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t total = variant_region_bytes(m);
  // ---- derived topology tables (dof-chain masks, subtree masks, scan jump tables, sparse-M indices) ----
  const int nbody = m->nbody, nv = m->nv, nM = m->nM;
  d.nmaskword = std::max(1, (nv + 31) / 32);
  d.nbodyword = std::max(1, (nbody + 31) / 32);
  std::vector<unsigned> mask((size_t)nbody * d.nmaskword, 0u);
  for (int b = 1; b < nbody; b++) {
    int bb = b;
    while (bb && !m->body_dofnum[bb]) bb = m->body_parentid[bb];
    if (!bb) continue;
    for (int k = m->body_dofadr[bb] + m->body_dofnum[bb] - 1; k >= 0; k = m->dof_parentid[k])
      mask[(size_t)b * d.nmaskword + (k >> 5)] |= 1u << (k & 31);
  }
  std::vector<unsigned> submask((size_t)nbody * d.nbodyword, 0u);
  int maxlevel = 0;
  for (int i = 0; i < nbody; i++) {
    maxlevel = std::max(maxlevel, m->body_level[i]);
    for (int b = i;; b = m->body_parentid[b]) {
      submask[(size_t)b * d.nbodyword + (i >> 5)] |= 1u << (i & 31);
      if (b == 0) break;
    }
  }
  d.njump = 0;
  while ((1 << d.njump) < maxlevel) d.njump++;
  std::vector<int> jump((size_t)std::max(1, d.njump) * nbody, 0);
  for (int r = 0; r < d.njump; r++)
    for (int i = 0; i < nbody; i++) {
      // ancestor 2^r levels up; world (0) carries the identity transform, so "none" == 0
      int a = i;
      for (int s = 0; s < (1 << r) && a > 0; s++) a = m->body_parentid[a];
      jump[(size_t)r * nbody + i] = (m->body_level[i] > (1 << r)) ? a : 0;
    }
  std::vector<int> M_row(std::max(1, nM)), M_col(std::max(1, nM)), M_ancadr(std::max(1, nM)), nanc(std::max(1, nv), 0);
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j], adr++) {
      M_row[adr] = i;
      M_col[adr] = j;
      M_ancadr[adr] = m->dof_Madr[j];
      if (j != i) nanc[i]++;
    }
  }
  // dofs whose motion is accumulated into cvel BEFORE a dof's cdof_dot is taken (mj_comVel order):
  // all chain dofs of earlier joints; for a free joint's rotational dofs also its translational ones
  std::vector<unsigned> premask((size_t)std::max(1, nv) * d.nmaskword, 0u);
  for (int k = 0; k < nv; k++) {
    const int jid = m->dof_jntid[k];
    const int jt = m->jnt_type[jid], jda = m->jnt_dofadr[jid];
    int first = jda;                                   // first dof of this dof's velocity group
    if (jt == B2MJ_JNT_FREE && k >= jda + 3) first = jda + 3;
    for (int j = m->dof_parentid[k]; j >= 0; j = m->dof_parentid[j])
      if (j < first) premask[(size_t)k * d.nmaskword + (j >> 5)] |= 1u << (j & 31);
  }
  int maxdepth = 0;
  for (int i = 0; i < nv; i++) maxdepth = std::max(maxdepth, nanc[i]);
  d.ndoflevel = nv ? maxdepth + 1 : 0;
  std::vector<int> lvl_adr(d.ndoflevel + 1, 0), lvl_dof(std::max(1, nv));
  for (int i = 0; i < nv; i++) lvl_adr[nanc[i] + 1]++;
  for (int l = 0; l < d.ndoflevel; l++) lvl_adr[l + 1] += lvl_adr[l];
  {
    std::vector<int> fill(lvl_adr.begin(), lvl_adr.end());
    for (int i = 0; i < nv; i++) lvl_dof[fill[nanc[i]]++] = i;
  }
  std::vector<int> desc_adr(nv + 1, 0), desc_dof(std::max(1, nM - nv)), desc_t(std::max(1, nM - nv));
  for (int t = 0; t < nM; t++) if (M_row[t] != M_col[t]) desc_adr[M_col[t] + 1]++;
  for (int k = 0; k < nv; k++) desc_adr[k + 1] += desc_adr[k];
  {
    std::vector<int> fill(desc_adr.begin(), desc_adr.end());
    for (int t = 0; t < nM; t++)
      if (M_row[t] != M_col[t]) { const int k = M_col[t]; desc_dof[fill[k]] = M_row[t]; desc_t[fill[k]] = t; fill[k]++; }
  }
  struct Extra { const void* src; size_t bytes; const void** slot; };
  const Extra extras[] = {
      {mask.data(), mask.size() * sizeof(unsigned), (const void**)&d.body_dofmask},
      {submask.data(), submask.size() * sizeof(unsigned), (const void**)&d.body_submask},
      {jump.data(), jump.size() * sizeof(int), (const void**)&d.body_jump},
      {M_row.data(), M_row.size() * sizeof(int), (const void**)&d.M_row},
      {M_col.data(), M_col.size() * sizeof(int), (const void**)&d.M_col},
      {M_ancadr.data(), M_ancadr.size() * sizeof(int), (const void**)&d.M_ancadr},
      {nanc.data(), nanc.size() * sizeof(int), (const void**)&d.dof_nanc},
      {premask.data(), premask.size() * sizeof(unsigned), (const void**)&d.dof_premask},
      {lvl_adr.data(), lvl_adr.size() * sizeof(int), (const void**)&d.doflevel_adr},
      {lvl_dof.data(), lvl_dof.size() * sizeof(int), (const void**)&d.doflevel_dof},
      {desc_adr.data(), desc_adr.size() * sizeof(int), (const void**)&d.dof_descadr},
      {desc_dof.data(), desc_dof.size() * sizeof(int), (const void**)&d.dof_desc_dof},
      {desc_t.data(), desc_t.size() * sizeof(int), (const void**)&d.dof_desc_adr},
  };
  for (const Extra& x : extras) total += al(x.bytes + 16);
  if (h->model_blob && h->model_blob_bytes < total) {
    cudaFree(h->model_blob);
    h->model_blob = nullptr;
  }
  if (!h->model_blob) {
    CUDA_OK(cudaMalloc(&h->model_blob, total));
    h->model_blob_bytes = total;
  }
  std::vector<unsigned char> host(total, 0);
  pack_variant(m, host.data(), &d, (unsigned char*)h->model_blob);
  d.env_model_stride = 0;
  size_t off = variant_region_bytes(m);
  for (const Extra& x : extras) {
    if (x.bytes) std::memcpy(host.data() + off, x.src, x.bytes);
    *x.slot = (unsigned char*)h->model_blob + off;
    off += al(x.bytes + 16);
  }
  CUDA_OK(cudaMemcpy(h->model_blob, host.data(), total, cudaMemcpyHostToDevice));
#define X(n) d.n = m->n;
  B2MJ_MODEL_SIZES(X)
#undef X
  d.opt = m->opt;
  d.meaninertia = m->stat.meaninertia;
  d.any_damping = 0;
  for (int i = 0; i < m->nv; i++) d.any_damping |= m->dof_damping[i] > 0;
  d.dense_small = (m->nv > 0 && m->nv <= 16 && !getenv("B2MJ_NO_DENSE")) ? 1 : 0;
  d.need_rnepost = d.need_subtreevel = 0;
  for (int i = 0; i < m->nsensor; i++) {
    const int t = m->sensor_type[i];
    if (t == B2MJ_SENS_ACCELEROMETER || t == B2MJ_SENS_FORCE || t == B2MJ_SENS_TORQUE || t == B2MJ_SENS_FRAMELINACC ||
        t == B2MJ_SENS_FRAMEANGACC)
      d.need_rnepost = 1;
    if (t == B2MJ_SENS_SUBTREELINVEL || t == B2MJ_SENS_SUBTREEANGMOM) d.need_subtreevel = 1;
  }
  return 0;
}

// largest contact dimension the model can produce (geom condim, explicit pair dim): sizes the per-contact cone Hessian
static int model_max_condim(const b2mjModel* m) {
  int mx = 1;
  for (int g = 0; g < m->ngeom; g++) mx = std::max(mx, m->geom_condim[g]);
  for (int p = 0; p < m->npair; p++) mx = std::max(mx, m->pair_dim[p]);
  return std::min(mx, 6);
}

static int check_supported(const b2mjModel* m) {
  for (int p = 0; p < m->ncollpair; p++) {
    const int t1 = m->geom_type[m->collpair_geom1[p]], t2 = m->geom_type[m->collpair_geom2[p]];
    if (t2 == B2MJ_GEOM_HFIELD || (t1 == B2MJ_GEOM_HFIELD && m->geom_dataid[m->collpair_geom1[p]] < 0)) {
      set_error("height field pair without a narrowphase function (plane-hfield, hfield-hfield, or an hfield geom without data)");
      return B2MJ_EUNSUPPORTED;
    }
  }
  if (m->opt.integrator != B2MJ_INT_EULER && m->opt.integrator != B2MJ_INT_RK4 && m->opt.integrator != B2MJ_INT_IMPLICIT &&
      m->opt.integrator != B2MJ_INT_IMPLICITFAST) {
    set_error("unknown integrator");
    return B2MJ_EUNSUPPORTED;
  }
  if (m->opt.solver != B2MJ_SOL_PGS && m->opt.solver != B2MJ_SOL_NEWTON && m->opt.solver != B2MJ_SOL_CG) {
    set_error("unknown solver");
    return B2MJ_EUNSUPPORTED;
  }
  if ((m->opt.density > 0 || m->opt.viscosity > 0) &&
      (m->opt.integrator == B2MJ_INT_IMPLICIT || m->opt.integrator == B2MJ_INT_IMPLICITFAST)) {
    set_error("fluid forces with an implicit integrator need the fluid velocity derivatives (mjd_inertiaBoxFluid): not implemented");
    return B2MJ_EUNSUPPORTED;
  }
  if (m->opt.solver != B2MJ_SOL_PGS && m->nv > B2K_NEWTON_MAX_NV) {
    // the Newton / CG triangular solves keep the solution vector in a 4-register-per-lane window (cholSolve_warp)
    set_error("Newton / CG solvers support at most " + std::to_string(B2K_NEWTON_MAX_NV) + " dofs per env (model has " +
              std::to_string(m->nv) + ")");
    return B2MJ_EUNSUPPORTED;
  }
  return 0;
}

// arena + record layout and the launch shape
static int make_layout(Handle* h) {
  const b2mjModel* m = h->model;
  DevModel& d = h->dm;
  const int nv = m->nv, nq = m->nq, na = m->na, nu = m->nu;
  auto even = [](int x) { return (x + 1) & ~1; };
  // record
  int o = 0;
  d.rec_A_begin = o;
  d.rec_ctrl = o; o += nu;
  d.rec_qfrc_applied = o; o += nv;
  o = even(o);
  d.rec_B_begin = o;
  d.rec_qpos = o; o += nq;
  d.rec_qvel = o; o += nv;
  d.rec_act = o; o += na;
  d.rec_warm = o; o += nv;
  d.rec_time = o; o += 1;
  o = even(o);
  d.rec_C_begin = o;
  d.rec_qacc = o; o += nv;
  d.rec_sensordata = o; o += m->nsensordata;
  d.rec_act_dot = o; o += na;
  o = even(o);
  d.rec_end = o;
  d.rec_pitch = o;
  // field sizes
  for (int f = 0; f < B2MJ_NFIELD; f++) {
    int is_int = 0;
    d.fsize[f] = std::max(0, field_count(m, f, &is_int));
    d.fis_int[f] = (unsigned char)is_int;
  }
  const bool pgs = m->opt.solver == B2MJ_SOL_PGS, newton = m->opt.solver == B2MJ_SOL_NEWTON;
  const bool rk4 = m->opt.integrator == B2MJ_INT_RK4;
  int* xs = d.xsize;
  for (int i = 0; i < XF_COUNT; i++) xs[i] = 0;
  xs[XF_QLOC] = 0;
  xs[XF_QH] = m->nM;
  xs[XF_QHDIAGINV] = nv;
  const bool dual = pgs || m->opt.noslip_iterations > 0;  // mj_isDual: the noslip pass runs on the dual problem
  xs[XF_EFC_MINVJT] = dual ? m->njmax * nv : 0;
  xs[XF_EFC_ARDIAG] = dual ? m->njmax : 0;
  for (int i = XF_VEC0; i <= XF_VEC5; i++) xs[i] = nv;
  xs[XF_EFC_JAREF] = m->njmax;
  xs[XF_EFC_JV] = pgs ? 0 : m->njmax;
  xs[XF_EFC_QUAD] = pgs ? 0 : 3 * m->njmax;
  // team mode (kernels/team.cuh): wide Newton models get one env per CTA and 8 warps; H then has an odd leading dimension
  // (the noslip pass is single-warp code over dense J rows: no team mode with it)
  d.team_warps = (newton && nv >= B2K_TEAM_MIN_NV && m->opt.noslip_iterations <= 0 && !getenv("B2MJ_NO_TEAM")) ? B2K_TEAM_WARPS : 1;
  if (d.team_warps > 1 && getenv("B2MJ_TEAM_WARPS")) d.team_warps = std::max(2, std::min(16, atoi(getenv("B2MJ_TEAM_WARPS"))));
  d.ldh = nv | 1;  // odd: lanes walking a column (or owning a row each) hit distinct shared-memory banks
  xs[XF_NEWTON_H] = !newton ? 0 : d.team_warps > 1 ? nv * (nv + 1) / 2 : nv * d.ldh;  // team mode: packed lower triangle
  xs[XF_JCOLS] = d.team_warps > 1 ? (m->njmax * 17 + 7) / 8 : 0;
  // cone Hessian blocks: dim x dim per contact, dim <= the largest condim of the model (9 doubles for condim 3, not the
  // 36 of condim 6: C5's 160 contact slots shrink from 46 KB to 11.5 KB and come on chip)
  d.conh_stride = model_max_condim(m) * model_max_condim(m);
  xs[XF_CONTACT_H] = (newton && m->opt.cone == B2MJ_CONE_ELLIPTIC) ? d.conh_stride * m->nconmax : 0;
  xs[XF_SUBTREE_LINVEL] = d.need_subtreevel ? 3 * m->nbody : 0;
  xs[XF_SUBTREE_ANGMOM] = d.need_subtreevel ? 3 * m->nbody : 0;
  xs[XF_BODYVEL] = d.need_subtreevel ? 6 * m->nbody : 0;
  xs[XF_RK_X0] = rk4 ? nq + nv + na + 1 : 0;  // + the step's start time (split RK4 steps span several launches)
  xs[XF_RK_XF] = rk4 ? 4 * nv : 0;
  xs[XF_RK_F] = rk4 ? 4 * (nv + na) : 0;
  xs[XF_RK_DX] = rk4 ? 2 * nv + na : 0;
  xs[XF_SCRATCH] = std::max(14 * m->nbody, 6 * nv);
  xs[XF_QW] = m->nM;
  xs[XF_QHW] = m->nM;
  xs[XF_MINV] = d.dense_small ? nv * nv : 0;
  // reserved whether or not the model has damping now: b2mj_model_update may switch damping on later and the arena
  // layout of a live handle cannot change
  xs[XF_HINV] = (d.dense_small && !rk4) ? nv * nv : 0;
  xs[XF_PRIMAL] = pgs ? 0 : 8 * nv;
  xs[XF_EFC_AR] = pgs ? m->njmax * (m->njmax + 4) : 0;
  xs[XF_EFC_AR_S] = pgs ? std::min(m->njmax * (m->njmax + 4), 384) : 0;  // nefc <= 17 stays on chip
  // triangle pair table (2 bytes per entry): sparse L'DL of qM / qH and the Newton Cholesky index their trailing blocks
  // through it; team mode has its own factorisation
  xs[XF_TRI] = (nv <= 255 && d.team_warps == 1 && (!d.dense_small || newton)) ? (nv * (nv + 1) / 2 * 2 + 7) / 8 : 0;
  const bool implicit_full = m->opt.integrator == B2MJ_INT_IMPLICIT;
  xs[XF_IMPL_LU] = implicit_full ? nv * nv : 0;
  xs[XF_IMPL_D] = implicit_full ? 6 * m->nbody * nv : 0;
  // xfrc_applied / mocap live in their own HBM arrays (read only when the surface is enabled)
  d.fsize[B2MJ_F_XFRC_APPLIED] = 0;


  // shared placement: record image first, then hot fields; big constraint arrays are demoted to the
  // global arena (L2) when the per-env footprint would starve occupancy
  std::vector<char> cold(B2MJ_NFIELD, 0), xcold(XF_COUNT, 0);
  xcold[XF_EFC_AR] = 1;
  xcold[XF_IMPL_LU] = xcold[XF_IMPL_D] = 1;  // once-per-step matrices of the implicit integrator
  // API-only fields (read back only through an arena dump) never occupy shared memory
  cold[B2MJ_F_XIMAT] = 1;
  if (!d.need_rnepost) cold[B2MJ_F_CACC] = cold[B2MJ_F_CFRC_INT] = cold[B2MJ_F_CFRC_EXT] = 1;
  if (d.dense_small) {
    cold[B2MJ_F_QLD] = cold[B2MJ_F_QLDIAGINV] = cold[B2MJ_F_QLDIAGSQRTINV] = 1;
    xcold[XF_QH] = xcold[XF_QHDIAGINV] = xcold[XF_QW] = xcold[XF_QHW] = 1;
  }
  auto smem_bytes_env = [&]() {
    size_t dbl = d.rec_end, ints = 0;
    for (int f = 0; f < B2MJ_NFIELD; f++) {
      if (cold[f] || is_record_field(f)) continue;
      if (d.fis_int[f]) ints += d.fsize[f];
      else dbl += d.fsize[f];
    }
    for (int i = 0; i < XF_COUNT; i++)
      if (!xcold[i]) dbl += xs[i];
    return ((dbl * 8 + ((ints + 1) & ~(size_t)1) * 4) + 15) & ~(size_t)15;
  };
  const size_t kSmPerSM = 228 * 1024, kCtaReserve = 1024, kMaxCta = 227 * 1024;
  size_t target = h->smem_target_bytes ? h->smem_target_bytes : 16 * 1024;  // >= 14 envs per SM
  if (const char* env = getenv("B2MJ_SMEM_TARGET_KB")) target = (size_t)atoi(env) * 1024;
  // demotion candidates: everything sized by njmax / nconmax (the constraint / contact working set) plus
  // the dense solver matrices, largest first.  The small AR window (XF_EFC_AR_S) and the per-env vectors
  // always stay in shared memory.
  struct Cand { int is_x, id; size_t bytes; };
  std::vector<Cand> cands;
  for (int f = 0; f < B2MJ_NFIELD; f++) {
    const bool efc = f >= B2MJ_F_EFC_TYPE && f <= B2MJ_F_EFC_AR, con = f >= B2MJ_F_CONTACT_DIST && f <= B2MJ_F_CONTACT_EFC_ADDRESS;
    if ((efc || con) && d.fsize[f]) cands.push_back({0, f, (size_t)d.fsize[f] * (d.fis_int[f] ? 4 : 8)});
  }
  for (int x : {XF_NEWTON_H, XF_EFC_MINVJT, XF_EFC_QUAD, XF_CONTACT_H, XF_EFC_ARDIAG, XF_EFC_JAREF, XF_EFC_JV})
    if (xs[x] && !(x == XF_NEWTON_H && d.team_warps > 1))  // the team factorises H in shared memory: never demoted
      cands.push_back({1, x, (size_t)xs[x] * 8});
  std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.bytes > b.bytes; });
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const int reg_envs = 512 / B2K_G;  // 128 registers per thread
  // envs an SM keeps resident when CTAs hold W envs each: asked of the occupancy calculator (registers, allocation
  // granularity, per-CTA reserve), not estimated
  auto resident = [&](int W) -> int {
    const int tw = d.team_warps;
    const size_t cta = (size_t)W * (smem_bytes_env() + (tw > 1 ? 64 : 16));
    if (W < 1 || (tw > 1 && W > 1) || W * B2K_G * tw > B2K_MAX_THREADS || cta > kMaxCta - kCtaReserve) return 0;
    int ctas = 0;
    if (b2k_occupancy(W * B2K_G * tw, cta, &ctas) != 0) {
      ctas = (int)(kSmPerSM / (cta + kCtaReserve));
      ctas = std::min(ctas, reg_envs / W);
    }
    return std::min(ctas, 32) * W;
  };
  auto envs_per_sm = [&](int* bestW_out) {
    int bestW = 0, bestEnv = 0;
    const int forceW = getenv("B2MJ_WARPS_PER_CTA") ? atoi(getenv("B2MJ_WARPS_PER_CTA")) : h->force_warps_per_cta;
    for (int W = 1; W <= B2K_STEP_THREADS / B2K_G; W++) {
      if (forceW && W != forceW) continue;
      // among equals prefer the smallest CTA: a finished env frees its shared memory for the next one
      // immediately instead of waiting for its CTA mates
      const int n = resident(W);
      if (n > bestEnv) { bestEnv = n; bestW = W; }
    }
    if (bestW_out) *bestW_out = bestW;
    return bestEnv;
  };
  // the lock-stepped rollout CTA (half of an SM's envs) must stay resident twice per SM as well
  auto rollout_shape_ok = [&](int E) { return E < 4 || (E % 2) || resident(E / 2) >= E; };
  // Placement policy (round 2).  All candidates start in the HBM/L2 arena; they are promoted to shared memory hottest
  // first while the batch keeps its residency.  Residency target: with only the fixed fields on chip an SM holds E0
  // envs, the batch then needs w = ceil(nenv / (E0 * SMs)) waves; E = ceil(nenv / (w * SMs)) <= E0 is the SMALLEST
  // residency that still runs in w waves, and every byte of shared memory beyond E envs per SM is free to hold solver
  // state.  For the BASELINE batch sizes: C2 (4096 envs) E0 = E = 14 -> the round-1 layout; C3 (1024 envs) E0 = 6,
  // w = 2, E = 4 -> 57 KB per env; C4 (2048 envs) E0 = E = 4 -> 57 KB.  In round 1 every solver array of the Newton
  // configs (H, Jaref, Jv, quad, D, state, ... even the 4.6 KB Hessian of the hand) lived in L2: each dependent access
  // in the Cholesky / line search paid ~400 cycles (solve = 69 % of a C3 / C4 step, 94 % of C5).
  {
    (void)target;
    for (const Cand& c : cands) { if (c.is_x) xcold[c.id] = 1; else cold[c.id] = 1; }
    int E = 1;
    {
      const int E0 = std::max(1, envs_per_sm(nullptr));
      const int w = (h->nenv + E0 * sms - 1) / (E0 * sms);
      E = std::max(1, std::min(E0, (h->nenv + w * sms - 1) / (w * sms)));
      if (const char* env = getenv("B2MJ_ENVS_PER_SM")) E = std::max(1, atoi(env));
    }
    // hottest first: what the solver touches in every iteration, then once-per-step arrays, largest last
    std::vector<std::pair<int, int>> order;  // (is_x, id)
    if (newton || m->opt.solver == B2MJ_SOL_CG) {
      for (int x : {XF_NEWTON_H, XF_EFC_JAREF, XF_EFC_JV, XF_EFC_QUAD}) order.push_back({1, x});
      for (int f : {B2MJ_F_EFC_D, B2MJ_F_EFC_STATE, B2MJ_F_EFC_FORCE, B2MJ_F_EFC_AREF, B2MJ_F_EFC_TYPE, B2MJ_F_EFC_ID,
                    B2MJ_F_EFC_R, B2MJ_F_EFC_FRICTIONLOSS, B2MJ_F_CONTACT_DIM, B2MJ_F_CONTACT_EFC_ADDRESS,
                    B2MJ_F_CONTACT_MU, B2MJ_F_CONTACT_FRICTION})
        order.push_back({0, f});
      order.push_back({1, XF_CONTACT_H});
      for (int f : {B2MJ_F_EFC_B, B2MJ_F_EFC_J, B2MJ_F_EFC_VEL, B2MJ_F_EFC_POS, B2MJ_F_EFC_MARGIN}) order.push_back({0, f});
    } else {
      for (int f : {B2MJ_F_EFC_FORCE, B2MJ_F_EFC_B, B2MJ_F_EFC_TYPE, B2MJ_F_EFC_ID}) order.push_back({0, f});
      order.push_back({1, XF_EFC_ARDIAG});
      for (int f : {B2MJ_F_CONTACT_DIM, B2MJ_F_CONTACT_EFC_ADDRESS, B2MJ_F_CONTACT_GEOM1, B2MJ_F_CONTACT_GEOM2,
                    B2MJ_F_CONTACT_EXCLUDE, B2MJ_F_EFC_R, B2MJ_F_EFC_D, B2MJ_F_EFC_AREF, B2MJ_F_EFC_FRICTIONLOSS,
                    B2MJ_F_CONTACT_FRICTION})
        order.push_back({0, f});
      order.push_back({1, XF_EFC_JAREF});
      order.push_back({0, B2MJ_F_EFC_J});
      order.push_back({1, XF_EFC_MINVJT});
    }
    const bool no_promote = getenv("B2MJ_NO_PROMOTE") != nullptr;
    std::vector<Cand> rest(cands.begin(), cands.end());
    std::stable_sort(rest.begin(), rest.end(), [](const Cand& a, const Cand& b) { return a.bytes < b.bytes; });
    std::vector<std::pair<int, int>> rest_order;
    for (const Cand& c : rest) rest_order.push_back({c.is_x, c.id});
    // place(E): everything back to the L2 arena, then promotion at residency E; returns the rows of the efc_J window
    auto place = [&](int Etarget) -> int {
    const int E = Etarget;
    for (const Cand& c : cands) { if (c.is_x) xcold[c.id] = 1; else cold[c.id] = 1; }
    xs[XF_JWIN] = 0;
    xs[XF_JVALS] = 0;
    int promote_left = getenv("B2MJ_PROMOTE_MAX") ? atoi(getenv("B2MJ_PROMOTE_MAX")) : 1 << 30;  // experiment knob
    auto promote = [&](const std::vector<std::pair<int, int>>& list) {
    for (const auto& o : list) {
      char& flag = o.first ? xcold[o.second] : cold[o.second];
      const int sz = o.first ? xs[o.second] : d.fsize[o.second];
      if (!flag || !sz || no_promote) continue;
      if (o.first && o.second == XF_EFC_AR) continue;  // the full AR always stays in the L2 arena
      if (promote_left <= 0) break;
      flag = 0;
      if (envs_per_sm(nullptr) < E || !rollout_shape_ok(E)) flag = 1;  // does not fit at this residency: stays in L2
      else promote_left--;
    }
    };
    // team mode sizes its mirror of the Jacobian entries (below) right after the per-iteration arrays, ahead of the
    // once-per-step ones at the tail of `order`
    std::vector<std::pair<int, int>> hot_order, warm_order;
    for (const auto& o : order) {
      const bool warm = !o.first && (o.second == B2MJ_F_EFC_B || o.second == B2MJ_F_EFC_J || o.second == B2MJ_F_EFC_VEL ||
                                     o.second == B2MJ_F_EFC_POS || o.second == B2MJ_F_EFC_MARGIN);
      (warm ? warm_order : hot_order).push_back(o);
    }
    const bool jvals_first = d.team_warps > 1 && !no_promote && !getenv("B2MJ_NO_JVALS");
    if (jvals_first) promote(hot_order); else promote(order);
    // primal solvers: whatever shared memory is still free at this residency becomes a window for the ACTIVE rows of
    // efc_J (stages_constraint.cuh::solveJ), when the full njmax-row Jacobian itself stayed in L2
    d.jwin_rows = 0;
    if (!pgs && d.team_warps == 1 && cold[B2MJ_F_EFC_J] && !no_promote && !getenv("B2MJ_NO_JWIN")) {
      int lo = 0, hi = std::min(m->njmax, 4096);
      while (lo < hi) {  // largest row count that keeps the residency
        const int mid = (lo + hi + 1) / 2;
        xs[XF_JWIN] = 2 + mid * nv;
        if (envs_per_sm(nullptr) >= E && rollout_shape_ok(E)) lo = mid; else hi = mid - 1;
      }
      // (rows beyond what a step ever uses would only crowd out the cold arrays below: cap at 64)
      d.jwin_rows = lo >= 4 ? std::min(lo, 64) : 0;
      xs[XF_JWIN] = d.jwin_rows ? 2 + d.jwin_rows * nv : 0;
    }
    // team mode: the non-zero entries of efc_J rows (at most 16 per row, the column lists of XF_JCOLS) mirrored in
    // shared memory for as many rows as fit -- the J' D J build, J v and J' f then read LDS instead of paying an L2 round
    // trip per constraint row on a serial chain (team.cuh)
    d.jvals_rows = 0;
    xs[XF_JVALS] = 0;
    if (jvals_first) {
      int lo = 0, hi = m->njmax;
      while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        xs[XF_JVALS] = mid * 16;
        if (envs_per_sm(nullptr) >= E) lo = mid; else hi = mid - 1;
      }
      d.jvals_rows = lo >= 32 ? lo : 0;
      xs[XF_JVALS] = d.jvals_rows * 16;
      promote(warm_order);
    }
    // the rest, smallest first
    promote(rest_order);
    return d.jwin_rows;
    };
    place(E);
    // The wave count is a crude model of a launch whose envs differ 3x in cost.  Measured on C4 (2048 humanoid envs):
    // at 5 envs/SM only the Hessian fits on chip and every access of the line search / update to Jaref, Jv, quad, D,
    // state, force is an L2 round trip: 1.20 M env-steps/s per step; at 4 envs/SM the whole hot set is in shared memory:
    // 1.26 M (fused rollout 1.41 M against 1.21 M); 3 envs/SM 1.06 M.  So from five envs per SM on, one or two envs
    // of residency are given up when that is what brings the solver's per-iteration working set on chip
    // (profiles/r2c_shape_sweep.txt).
    auto hot_on_chip = [&]() {
      for (int x : {XF_NEWTON_H, XF_EFC_JAREF, XF_EFC_JV, XF_EFC_QUAD})
        if (xs[x] && xcold[x]) return false;
      for (int f : {B2MJ_F_EFC_D, B2MJ_F_EFC_STATE, B2MJ_F_EFC_FORCE, B2MJ_F_EFC_AREF, B2MJ_F_EFC_TYPE, B2MJ_F_EFC_ID})
        if (d.fsize[f] && cold[f]) return false;
      return true;
    };
    if (!pgs && d.team_warps == 1 && E >= 5 && !getenv("B2MJ_ENVS_PER_SM") && !getenv("B2MJ_NO_HOTSET_TRADE") && !hot_on_chip()) {
      // (up to two envs: 1536 hand envs would run 6 per SM with nothing but H on chip, 0.398 M env-steps/s; 5 per SM
      // still leaves type / id / friction in L2, 0.421 M; 4 per SM 0.445 M)
      bool ok = false;
      for (int Et = E - 1; Et >= std::max(4, E - 2) && !ok; Et--) {
        place(Et);
        ok = hot_on_chip();
      }
      if (!ok) place(E);
    }
  }
  // global (full) arena offsets (after every size is final: the efc_J window above is sized last)
  int gd = 0, gi = 0;
  for (int f = 0; f < B2MJ_NFIELD; f++) {
    if (d.fis_int[f]) { d.off_g[f] = gi; gi += d.fsize[f]; }
    else { d.off_g[f] = gd; gd += d.fsize[f]; }
  }
  for (int i = 0; i < XF_COUNT; i++) { d.xoff_g[i] = gd; gd += xs[i]; }
  d.arena_g_doubles = even(gd);
  d.arena_g_ints = even(gi);

  // A second tier that also demoted write-once/read-once kinematic fields (geom frames, crb, cinert, cvel, ...)
  // was measured and dropped (profiles/r1_layout_sweep.txt): it buys throughput only for contact-free batches far
  // beyond one wave (18M -> 26M env-steps/s at 16k envs) and costs ~8% at the BASELINE batch because every access
  // to a demotable field goes through a generic pointer instead of an LDS address.
  if (smem_bytes_env() + 16 > kMaxCta - kCtaReserve) {
    set_error("model does not fit the per-warp shared-memory arena (" + std::to_string(smem_bytes_env()) +
              " bytes per env even with the constraint working set in HBM)");
    return B2MJ_EUNSUPPORTED;
  }
  int sdo = d.rec_end, sio = 0;
  for (int f = 0; f < B2MJ_NFIELD; f++) d.off_s[f] = -1;
  d.off_s[B2MJ_F_CTRL] = d.rec_ctrl; d.off_s[B2MJ_F_QFRC_APPLIED] = d.rec_qfrc_applied;
  d.off_s[B2MJ_F_QPOS] = d.rec_qpos; d.off_s[B2MJ_F_QVEL] = d.rec_qvel; d.off_s[B2MJ_F_ACT] = d.rec_act;
  d.off_s[B2MJ_F_QACC_WARMSTART] = d.rec_warm; d.off_s[B2MJ_F_TIME] = d.rec_time; d.off_s[B2MJ_F_QACC] = d.rec_qacc;
  d.off_s[B2MJ_F_SENSORDATA] = d.rec_sensordata; d.off_s[B2MJ_F_ACT_DOT] = d.rec_act_dot;
  for (int f = 0; f < B2MJ_NFIELD; f++) {
    if (is_record_field(f) || cold[f]) continue;
    if (d.fis_int[f]) { d.off_s[f] = sio; sio += d.fsize[f]; }
    else { d.off_s[f] = sdo; sdo += d.fsize[f]; }
  }
  for (int i = 0; i < XF_COUNT; i++) {
    if (xcold[i]) { d.xoff_s[i] = -1; continue; }
    d.xoff_s[i] = sdo;
    sdo += xs[i];
  }
  d.arena_s_doubles = even(sdo);
  d.arena_s_ints = even(sio);
  // PGS AR overlay (stages_solver.cuh::arPtr): the contiguous run of kinematic frame fields xpos .. crb, allowed when no
  // acc-stage sensor reads frames after the solve (touch, accelerometer, force, torque, frame accelerations do)
  d.ar_ovl_off = 0;
  d.ar_ovl_doubles = 0;
  if (pgs && !getenv("B2MJ_NO_AR_OVERLAY")) {
    // the implicit integrator differentiates the RNE pass after the solve: it reads cinert / cdof, which the overlay reuses
    bool ok = m->opt.integrator != B2MJ_INT_IMPLICIT;
    for (int i = 0; i < m->nsensor && ok; i++) {
      const int t = m->sensor_type[i];
      if (m->sensor_needstage[i] == B2MJ_STAGE_ACC && t != B2MJ_SENS_ACTUATORFRC && t != B2MJ_SENS_JOINTACTFRC &&
          t != B2MJ_SENS_JOINTLIMITFRC && t != B2MJ_SENS_TENDONLIMITFRC)
        ok = false;
    }
    // longest run of consecutively placed hot fields starting at xpos
    int begin = d.off_s[B2MJ_F_XPOS], end = begin;
    for (int f = B2MJ_F_XPOS; ok && begin >= 0 && f <= B2MJ_F_CRB; f++) {
      if (d.off_s[f] < 0 || d.fis_int[f] || !d.fsize[f]) continue;
      if (d.off_s[f] != end) break;
      end += d.fsize[f];
    }
    if (ok && begin >= 0 && end > begin) {
      d.ar_ovl_off = begin;
      d.ar_ovl_doubles = end - begin;
    }
  }
  if (getenv("B2MJ_PRINT_LAYOUT")) {
    static const char* xnames[] = {"QLOC", "QH", "QHDIAGINV", "EFC_MINVJT", "EFC_ARDIAG", "VEC0", "VEC1", "VEC2", "VEC3", "VEC4",
                                   "VEC5", "EFC_JAREF", "EFC_JV", "EFC_QUAD", "NEWTON_H", "CONTACT_H", "SUBTREE_LINVEL",
                                   "SUBTREE_ANGMOM", "BODYVEL", "RK_X0", "RK_XF", "RK_F", "RK_DX", "SCRATCH", "QW", "QHW",
                                   "EFC_AR", "MINV", "HINV", "PRIMAL", "EFC_AR_S", "JWIN", "JCOLS", "TRI", "IMPL_LU", "IMPL_D", "JVALS"};
    fprintf(stderr, "[b2mj layout] record %d doubles; shared arena %d doubles + %d ints per env\n", d.rec_end, d.arena_s_doubles,
            d.arena_s_ints);
    for (int f = 0; f < B2MJ_NFIELD; f++)
      if (d.fsize[f] && !is_record_field(f))
        fprintf(stderr, "  %-24s %6d %s %s\n", b2mj_field_name((b2mj_field)f), d.fsize[f], d.fis_int[f] ? "i32" : "f64",
                d.off_s[f] >= 0 ? "smem" : "HBM");
    for (int i = 0; i < XF_COUNT; i++)
      if (xs[i]) fprintf(stderr, "  x:%-22s %6d f64 %s\n", xnames[i], xs[i], d.xoff_s[i] >= 0 ? "smem" : "HBM");
  }
  const size_t env_bytes = (((size_t)d.arena_s_doubles * 8 + (size_t)d.arena_s_ints * 4) + 15) & ~(size_t)15;
  int bestW = 1;
  const int bestEnv = envs_per_sm(&bestW);
  if (bestEnv == 0) {
    set_error("model does not fit the shared-memory arena even with all optional arrays in HBM");
    return B2MJ_EUNSUPPORTED;
  }
  h->warps_per_cta = bestW;
  h->resident_envs = bestEnv * sms;
  h->smem_bytes = (size_t)bestW * (env_bytes + (d.team_warps > 1 ? 64 : 16));
  // Rollout shape: half of an SM's resident envs per CTA, stages lock-stepped with a CTA barrier.  Warps that
  // run the same stage together share instruction fetches -- the fused rollout is fetch bound (ncu: 12.7
  // no-instruction stall cycles per issue when the warps drift apart).  Measured on C2 at 4096 envs (CTA width x
  // stage barrier): 2 -> 11.6M, 5 -> 12.3M, 6 -> 11.2M, 7 -> 13.6M, 8 -> 9.4M, 14 -> 12.2M env-steps/s.
  h->rollout_warps_per_cta = 0;
  {
    const int Wr = bestEnv / 2;
    if (d.team_warps == 1 && Wr > bestW && Wr * B2K_G <= B2K_MAX_THREADS && resident(Wr) >= bestEnv)
      h->rollout_warps_per_cta = Wr;
  }
  if (const char* env = getenv("B2MJ_ROLLOUT_WARPS_PER_CTA")) h->rollout_warps_per_cta = atoi(env);
  h->arena_in_smem = 1;
  for (int f = 0; f < B2MJ_NFIELD; f++) if (!is_record_field(f) && d.fsize[f] && d.off_s[f] < 0) h->arena_in_smem = 0;
  for (int i = 0; i < XF_COUNT; i++) if (xs[i] && d.xoff_s[i] < 0) h->arena_in_smem = 0;
  return 0;
}

static int alloc_state(Handle* h) {
  const b2mjModel* m = h->model;
  DevModel& d = h->dm;
  const size_t n = (size_t)h->nenv;
  CUDA_OK(cudaMalloc(&h->rec, n * d.rec_pitch * sizeof(double)));
  CUDA_OK(cudaMalloc(&h->rec_init, (size_t)d.rec_pitch * sizeof(double)));
  CUDA_OK(cudaMalloc(&h->garena_d, std::max<size_t>(1, n * d.arena_g_doubles) * sizeof(double)));
  CUDA_OK(cudaMalloc(&h->garena_i, std::max<size_t>(1, n * d.arena_g_ints) * sizeof(int)));
  CUDA_OK(cudaMemset(h->garena_d, 0, std::max<size_t>(1, n * d.arena_g_doubles) * sizeof(double)));
  CUDA_OK(cudaMemset(h->garena_i, 0, std::max<size_t>(1, n * d.arena_g_ints) * sizeof(int)));
  CUDA_OK(cudaMalloc(&h->warning, n * B2MJ_NWARNING * sizeof(int)));
  CUDA_OK(cudaMalloc(&h->stats, n * 4 * sizeof(int)));
  CUDA_OK(cudaMalloc(&h->xfrc, n * 6 * m->nbody * sizeof(double)));
  CUDA_OK(cudaMemset(h->xfrc, 0, n * 6 * m->nbody * sizeof(double)));
  if (m->nmocap) {
    CUDA_OK(cudaMalloc(&h->mocap, n * 7 * m->nmocap * sizeof(double)));
    CUDA_OK(cudaMalloc(&h->mocap_init, (size_t)7 * m->nmocap * sizeof(double)));
  }
  CUDA_OK(cudaMalloc(&h->mask_dev, n));
  return 0;
}

static int upload_init_templates(Handle* h) {
  const b2mjModel* m = h->model;
  DevModel& d = h->dm;
  std::vector<double> rec(d.rec_pitch, 0.0);
  for (int i = 0; i < m->nq; i++) rec[d.rec_qpos + i] = m->qpos0[i];
  CUDA_OK(cudaMemcpy(h->rec_init, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (m->nmocap) {
    std::vector<double> mc(7 * m->nmocap, 0.0);
    for (int b = 0; b < m->nbody; b++) {
      const int id = m->body_mocapid[b];
      if (id < 0) continue;
      for (int k = 0; k < 3; k++) mc[3 * id + k] = m->body_pos[3 * b + k];
      for (int k = 0; k < 4; k++) mc[3 * m->nmocap + 4 * id + k] = m->body_quat[4 * b + k];
    }
    CUDA_OK(cudaMemcpy(h->mocap_init, mc.data(), mc.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  return 0;
}

int handle_launch(Handle* h, int mode, int nsteps, const double* ctrl_seq, double* traj_qpos, double* traj_qvel,
                  double* traj_sensor, int chunk) {
  LaunchArgs a;
  a.rec = h->rec;
  a.garena_d = h->garena_d;
  a.garena_i = h->garena_i;
  a.xfrc = h->dm.has_xfrc ? h->xfrc : nullptr;
  a.mocap = h->mocap;
  a.warning = h->warning;
  a.stats = h->stats;
  a.nenv = h->nenv;
  a.nsteps = nsteps;
  a.mode = mode;
  a.dump = h->keep_intermediates;
  a.rk_stage = h->rk_stage;
  a.prof = h->prof;
  a.ctrl_seq = ctrl_seq;
  a.traj_qpos = traj_qpos;
  a.traj_qvel = traj_qvel;
  a.traj_sensor = traj_sensor;
  a.sched = nullptr;
  a.chunk = 0;
  if (chunk > 0 && nsteps > chunk && !h->keep_intermediates) {
    if (!h->sched) CUDA_OK(cudaMalloc(&h->sched, (size_t)(1 + h->nenv) * sizeof(int)));
    CUDA_OK(cudaMemsetAsync(h->sched, 0, (size_t)(1 + h->nenv) * sizeof(int), h->stream));
    a.sched = h->sched;
    a.chunk = chunk;
  }
  // Launch shape.  The static fused rollout uses its wide lock-stepped shape.  A per-step launch of up to two waves starts
  // every warp in step anyway and keeps the small CTAs (a finished env frees its slot at once; the wide shape measured
  // -4 % at 4096 envs).  Beyond that the small CTAs of the later waves start whenever a slot frees, the fourteen warps of
  // an SM drift into fourteen different stages and the kernel turns instruction-fetch bound: per-env residency doubles
  // (213 k -> 462 k cycles at 16384 envs) and throughput FALLS with the batch (11.8 M at 4096 envs, 7.0 M at 65536).
  // With the lock-stepped shape it keeps rising: 13.2 M at 8192, 15.2 M at 65536 (profiles/r2c_batch_sweep.txt).
  int W = h->warps_per_cta;
  a.sync_stages = 0;
  static const char* step_wide_env = getenv("B2MJ_STEP_WIDE");
  const bool step_wide = step_wide_env ? atoi(step_wide_env) != 0 : (double)h->nenv > 2.5 * std::max(1, h->resident_envs);
  if (mode == MODE_STEP && (nsteps > 1 || step_wide) && !a.sched && !h->keep_intermediates && h->rollout_warps_per_cta > 0) {
    W = h->rollout_warps_per_cta;
    a.sync_stages = 1;
  }
  // Newton / CG models: the CTA mates of a per-step or rollout launch meet at every stage boundary as well -- their
  // code path is the longest in the kernel and two or three warps per SM running it out of step miss the instruction
  // cache on nearly every line (C3: per-step +3.7 %, fused rollout +16 %; C2 / PGS per-step launches measured -1 %)
  if (W >= 2 && h->model->opt.solver != B2MJ_SOL_PGS && h->dm.team_warps == 1 && !h->keep_intermediates) a.sync_stages = 1;
  if (const char* env = getenv("B2MJ_STAGE_SYNC")) a.sync_stages = atoi(env) ? 1 : 0;
  const size_t smem = h->smem_bytes / h->warps_per_cta * W;
  h->last_warps_per_cta = W;
  static const bool reorder = !getenv("B2MJ_NO_REORDER");
  static const int order_legacy = getenv("B2MJ_ORDER_LEGACY") ? 1 : 0;
  static const bool order_sync = order_legacy || getenv("B2MJ_ORDER_SYNC");
  a.perm = (reorder && h->perm_valid && !a.sched) ? h->perm : nullptr;
  a.cost = nullptr;
  // Single-step launches of a batch wider than one wave refresh the launch order OFF the critical path: the order kernel
  // of launch k runs on a side stream under launch k + 1 (one CTA; it takes the first slot a finished env frees), and
  // launch k + 1 uses the order of launch k - 1.  Contact states persist over steps, so a prediction that is one step
  // older orders as well, and the ~10 us order kernel no longer sits between two steps (C2: 353 us per step).
  // Worth it where the order kernel is a visible share of the step -- small models (nv <= 16, the dense-inverse class):
  // C2 per-step 11.55 M -> 11.78 M env-steps/s.  Where a step takes milliseconds the one-step-older prediction costs
  // more than the 10 us it hides (C3 450 k -> 416 k, C4 -1 %, C5 -2 %): those keep the synchronous refresh.
  static const char* async_env = getenv("B2MJ_ORDER_ASYNC");
  const bool async_model = async_env ? atoi(async_env) != 0 : h->dm.dense_small != 0;
  const bool order_async = reorder && !order_sync && async_model && nsteps == 1 && !a.sched && (mode == MODE_STEP || mode == MODE_STEP_END) &&
                           h->nenv > std::max(h->resident_envs, 32);
  if (order_async) {
    if (!h->order_stream) {
      CUDA_OK(cudaStreamCreateWithFlags(&h->order_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++) {
        CUDA_OK(cudaEventCreateWithFlags(&h->order_step_done[i], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&h->order_done[i], cudaEventDisableTiming));
      }
      CUDA_OK(cudaMalloc(&h->perm_async, (size_t)2 * h->nenv * sizeof(int)));
      CUDA_OK(cudaMalloc(&h->cost, (size_t)2 * h->nenv * sizeof(int)));
    }
    const uint64_t n = h->order_n;  // this launch is followed by order kernel n + 1
    if (n >= 2) {
      // order kernel n - 1 (finished under the previous launch) wrote perm_async[(n - 1) & 1] and was the last reader of
      // the cost buffer this launch overwrites
      CUDA_OK(cudaStreamWaitEvent(h->stream, h->order_done[(n - 1) & 1], 0));
      a.perm = h->perm_async + (size_t)((n - 1) & 1) * h->nenv;
    }  // (the first two launches keep whatever order a synchronous refresh left, or the identity)
    a.cost = h->cost + (size_t)((n + 1) & 1) * h->nenv;
  }
  a.env_model = nullptr;
  a.pub = h->launch_pub;
  a.pub_seq = h->launch_pub_seq;
  int rc;
  if (h->n_env_models > 0) {  // per-env model variants: the kernel build that reads every model array per env
    a.env_model = h->env_model_idx;
    h->dm_env.has_xfrc = h->dm.has_xfrc;
    rc = b2k_em_launch_step(&h->dm_env, &a, W, smem, h->stream);
  } else {
    rc = b2k_launch_step(&h->dm, &a, W, smem, h->stream);
  }
  if (rc != 0) {
    set_error(std::string("step kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    return B2MJ_ECUDA;
  }
  h->launches++;
  // heaviest-first launch order for the next launches, from the cost this one recorded per env (only worth a kernel
  // when the batch spans more than one wave: C5 has 512 envs and 148 slots)
  if (order_async) {
    const uint64_t j = ++h->order_n;
    CUDA_OK(cudaEventRecord(h->order_step_done[j & 1], h->stream));
    CUDA_OK(cudaStreamWaitEvent(h->order_stream, h->order_step_done[j & 1], 0));
    const int orc = b2k_launch_order(h->stats, h->cost + (size_t)(j & 1) * h->nenv, h->nenv,
                                     h->perm_async + (size_t)(j & 1) * h->nenv, 0, h->order_stream);
    if (orc != 0) {
      set_error(std::string("order kernel launch failed: ") + cudaGetErrorString((cudaError_t)orc));
      return B2MJ_ECUDA;
    }
    CUDA_OK(cudaEventRecord(h->order_done[j & 1], h->order_stream));
    h->launches++;
  } else if (reorder && mode != MODE_FORWARD && mode != MODE_STEP_BEGIN &&
             (order_legacy ? h->nenv >= 1024 : h->nenv > std::max(h->resident_envs, 32))) {
    // synchronous refresh (rollouts, B2MJ_ORDER_SYNC / B2MJ_ORDER_LEGACY).  After a multi-step rollout the residency is a
    // sum over steps and the lock-stepped rollout CTAs do better with heavy and light envs mixed: the coarse rows x
    // iterations classes of the last step stay in use there (13.2 M against 12.9 M env-steps/s on the C2 rollout)
    if (!h->perm) CUDA_OK(cudaMalloc(&h->perm, (size_t)h->nenv * sizeof(int)));
    const int orc = b2k_launch_order(h->stats, nullptr, h->nenv, h->perm, (order_legacy || nsteps > 1) ? 1 : 0, h->stream);
    if (orc != 0) {
      set_error(std::string("order kernel launch failed: ") + cudaGetErrorString((cudaError_t)orc));
      return B2MJ_ECUDA;
    }
    h->perm_valid = 1;
    h->launches++;
  }
  h->dump_valid = (h->keep_intermediates || mode == MODE_STEP_BEGIN);
  if (mode != MODE_STEP_END) h->rk_stage = 0;
  return 0;
}

}  // namespace b2mj

extern "C" {

int b2mj_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int b2mj_create(const b2mjModel* m, int nenv, int device, b2mj_handle** out) {
  if (!m || !out || nenv <= 0) {
    set_error("b2mj_create: bad argument");
    return B2MJ_EINVAL;
  }
  *out = nullptr;
  if (int rc = check_supported(m)) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: the batched step has no CPU fallback");
    return B2MJ_ENODEVICE;
  }
  if (device < 0 || device >= ndev) {
    set_error("b2mj_create: device index out of range");
    return B2MJ_EINVAL;
  }
  CUDA_OK(cudaSetDevice(device));
  Handle* h = new Handle();
  h->device = device;
  h->nenv = nenv;
  h->model = model_clone(m);
  int rc = upload_model(h);
  if (!rc) rc = make_layout(h);
  if (!rc) rc = alloc_state(h);
  if (!rc) rc = upload_init_templates(h);
  if (rc) {
    b2mj_destroy(reinterpret_cast<b2mj_handle*>(h));
    return rc;
  }
  *out = reinterpret_cast<b2mj_handle*>(h);
  return b2mj_reset(*out, nullptr);
}

void b2mj_destroy(b2mj_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->model_blob); cudaFree(h->rec); cudaFree(h->rec_init); cudaFree(h->garena_d); cudaFree(h->garena_i);
  cudaFree(h->warning); cudaFree(h->stats); cudaFree(h->xfrc); cudaFree(h->mocap); cudaFree(h->mocap_init);
  cudaFree(h->mask_dev);
  cudaFree(h->rec_key); cudaFree(h->mocap_key);
  cudaFree(h->prof);
  cudaFree(h->sched);
  cudaFree(h->perm);
  if (h->order_stream) {
    cudaStreamSynchronize(h->order_stream);
    cudaStreamDestroy(h->order_stream);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(h->order_step_done[i]); cudaEventDestroy(h->order_done[i]); }
  }
  cudaFree(h->perm_async);
  cudaFree(h->cost);
  cudaFree(h->publish_slab);
  cudaFree(h->env_blob);
  cudaFree(h->env_model_idx);
  handle_free_plugins(h);
  handle_free_fused_publish(h);
  b2mj_model_free(h->model);
  delete h;
}

int b2mj_nenv(const b2mj_handle* hh) { return hh ? reinterpret_cast<const Handle*>(hh)->nenv : B2MJ_EINVAL; }
const b2mjModel* b2mj_model(const b2mj_handle* hh) { return hh ? reinterpret_cast<const Handle*>(hh)->model : nullptr; }

int b2mj_set_stream(b2mj_handle* hh, void* stream) {
  if (!hh) return B2MJ_EINVAL;
  reinterpret_cast<Handle*>(hh)->stream = (cudaStream_t)stream;
  return 0;
}

int b2mj_reset(b2mj_handle* hh, const uint8_t* env_mask) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  const unsigned char* dmask = nullptr;
  if (env_mask) {
    CUDA_OK(cudaMemcpyAsync(h->mask_dev, env_mask, h->nenv, cudaMemcpyHostToDevice, h->stream));
    dmask = h->mask_dev;
  }
  const b2mjModel* m = h->model;
  reset_kernel<<<h->nenv, 64, 0, h->stream>>>(h->rec, h->rec_init, h->dm.rec_pitch, h->nenv, dmask, h->warning, h->stats,
                                              h->xfrc, 6 * m->nbody, h->mocap, h->mocap_init, 7 * m->nmocap);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  h->dump_valid = 0;
  h->in_split_step = 0;  // a reset between step_begin and step_end abandons the split step
  if (!env_mask) h->dm.has_xfrc = 0;
  handle_reset_plugins(h, env_mask);
  return 0;
}

int b2mj_reset_keyframe(b2mj_handle* hh, int key, const uint8_t* env_mask) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  const b2mjModel* m = h->model;
  if (key < 0 || key >= m->nkey) {
    set_error("b2mj_reset_keyframe: the model has " + std::to_string(m->nkey) + " keyframes, asked for " + std::to_string(key));
    return B2MJ_EINVAL;
  }
  CUDA_OK(cudaSetDevice(h->device));
  const DevModel& d = h->dm;
  // same path as b2mj_reset with the keyframe as the template record (mj_resetData, then the key's state on top)
  std::vector<double> rec(d.rec_pitch, 0.0);
  for (int i = 0; i < m->nq; i++) rec[d.rec_qpos + i] = m->key_qpos[(size_t)key * m->nq + i];
  for (int i = 0; i < m->nv; i++) rec[d.rec_qvel + i] = m->key_qvel[(size_t)key * m->nv + i];
  for (int i = 0; i < m->na; i++) rec[d.rec_act + i] = m->key_act[(size_t)key * m->na + i];
  for (int i = 0; i < m->nu; i++) rec[d.rec_ctrl + i] = m->key_ctrl[(size_t)key * m->nu + i];
  rec[d.rec_time] = m->key_time[key];
  if (!h->rec_key) CUDA_OK(cudaMalloc(&h->rec_key, (size_t)d.rec_pitch * sizeof(double)));
  CUDA_OK(cudaMemcpyAsync(h->rec_key, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (m->nmocap) {
    std::vector<double> mc(7 * m->nmocap);
    for (int i = 0; i < 3 * m->nmocap; i++) mc[i] = m->key_mpos[(size_t)key * 3 * m->nmocap + i];
    for (int i = 0; i < 4 * m->nmocap; i++) mc[3 * m->nmocap + i] = m->key_mquat[(size_t)key * 4 * m->nmocap + i];
    if (!h->mocap_key) CUDA_OK(cudaMalloc(&h->mocap_key, mc.size() * sizeof(double)));
    CUDA_OK(cudaMemcpyAsync(h->mocap_key, mc.data(), mc.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));  // the staging vectors go out of scope
  const unsigned char* dmask = nullptr;
  if (env_mask) {
    CUDA_OK(cudaMemcpyAsync(h->mask_dev, env_mask, h->nenv, cudaMemcpyHostToDevice, h->stream));
    dmask = h->mask_dev;
  }
  reset_kernel<<<h->nenv, 64, 0, h->stream>>>(h->rec, h->rec_key, d.rec_pitch, h->nenv, dmask, h->warning, h->stats, h->xfrc,
                                              6 * m->nbody, h->mocap, m->nmocap ? h->mocap_key : h->mocap_init, 7 * m->nmocap);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  h->dump_valid = 0;
  h->in_split_step = 0;
  if (!env_mask) h->dm.has_xfrc = 0;
  handle_reset_plugins(h, env_mask);
  return 0;
}

int b2mj_forward(b2mj_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  return handle_launch(h, MODE_FORWARD, 1);
}

int b2mj_step(b2mj_handle* hh, int nsteps) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  if (nsteps <= 0) {
    set_error("b2mj_step: nsteps must be positive");  // MujocoEnv::step returns false for n<=0 (mujoco_env.cpp:925)
    return B2MJ_EINVAL;
  }
  CUDA_OK(cudaSetDevice(h->device));
  h->in_split_step = 0;
  return handle_launch(h, MODE_STEP, nsteps);
}

int b2mj_rollout(b2mj_handle* hh, int nsteps, const double* dev_ctrl, double* dev_qpos_out, double* dev_qvel_out,
                 double* dev_sensor_out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  if (nsteps <= 0) {
    set_error("b2mj_rollout: nsteps must be positive");
    return B2MJ_EINVAL;
  }
  CUDA_OK(cudaSetDevice(h->device));
  h->in_split_step = 0;
  // Scheduling.  Static (chunk 0): one env per warp for the whole rollout; all warps start in step and share
  // instruction fetches (ncu: 15 -> 2 no-instruction stall cycles per issue against the ticketed grid) -- best when
  // the batch fills whole waves (C2 at 4096 envs: 11.6M vs 8.9M env-steps/s).  Ticketed: (env, chunk) work items over
  // a persistent grid, which wins when the last wave would be mostly empty.  B2MJ_ROLLOUT_CHUNK overrides.
  int chunk = 0;
  {
    int per_sm = 0, sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    if (h->dm.team_warps == 1 && b2k_occupancy(h->warps_per_cta * B2K_G, h->smem_bytes, &per_sm) == 0 && per_sm > 0) {
      const double slots = (double)per_sm * h->warps_per_cta * sms;
      const double waves = h->nenv / slots;
      const double idle = std::ceil(waves) - waves;  // empty fraction of the last wave
      // ... as a share of the whole launch: the ticketed grid gives up the lock-stepped instruction fetch (-25 %), so it
      // only pays when the empty tail is a large part of the launch (1.4 waves), never for a long one (65536 envs = 31.6
      // waves ran ticketed under the round-2 rule `idle > 0.35` and lost a third: 9.0 M against 14 M env-steps/s)
      if (waves > 1.0 && idle / std::ceil(waves) > 0.2) chunk = 64;
    }
  }
  if (const char* env = getenv("B2MJ_ROLLOUT_CHUNK")) chunk = atoi(env);
  return handle_launch(h, MODE_STEP, nsteps, dev_ctrl, dev_qpos_out, dev_qvel_out, dev_sensor_out, chunk);
}

int b2mj_step_begin(b2mj_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  h->rk_stage = 0;
  int rc = handle_launch(h, MODE_STEP_BEGIN, 1);
  if (!rc) h->in_split_step = 1;
  return rc;
}

int b2mj_step_end(b2mj_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  if (!h->in_split_step) {
    set_error("b2mj_step_end without b2mj_step_begin");
    return B2MJ_ESTATE;
  }
  CUDA_OK(cudaSetDevice(h->device));
  if (h->model->opt.integrator == B2MJ_INT_RK4) {
    // mj_RungeKutta runs four forward passes and the hooks fire in each: one launch per sub-step, B2MJ_AGAIN until
    // the fourth has been integrated
    const int stage = h->rk_stage;
    const int rc = handle_launch(h, MODE_STEP_END, 1);
    if (rc) { h->in_split_step = 0; return rc; }
    if (stage < 3) { h->rk_stage = stage + 1; h->dump_valid = 1; return B2MJ_AGAIN; }
    h->rk_stage = 0;
    h->in_split_step = 0;
    return 0;
  }
  h->in_split_step = 0;
  return handle_launch(h, MODE_STEP_END, 1);
}

int b2mj_sync(b2mj_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int b2mj_set_keep_intermediates(b2mj_handle* hh, int on) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  h->keep_intermediates = on ? 1 : 0;
  return 0;
}

static int rec_offset(const DevModel& d, int f) {
  switch (f) {
    case B2MJ_F_QPOS: return d.rec_qpos;
    case B2MJ_F_QVEL: return d.rec_qvel;
    case B2MJ_F_ACT: return d.rec_act;
    case B2MJ_F_CTRL: return d.rec_ctrl;
    case B2MJ_F_QFRC_APPLIED: return d.rec_qfrc_applied;
    case B2MJ_F_QACC_WARMSTART: return d.rec_warm;
    case B2MJ_F_TIME: return d.rec_time;
    case B2MJ_F_QACC: return d.rec_qacc;
    case B2MJ_F_SENSORDATA: return d.rec_sensordata;
    case B2MJ_F_ACT_DOT: return d.rec_act_dot;
    default: return -1;
  }
}

// resolve a field to (device base pointer, env pitch in bytes, element size, count)
static int locate(Handle* h, int f, bool for_write, unsigned char** base, size_t* pitch, size_t* esize, int* count) {
  if (f < 0 || f >= B2MJ_NFIELD) { set_error("unknown field"); return B2MJ_EINVAL; }
  const b2mjModel* m = h->model;
  const DevModel& d = h->dm;
  int is_int = 0;
  int n = b2mj_field_size(m, (b2mj_field)f, &is_int);
  *esize = is_int ? sizeof(int) : sizeof(double);
  *count = n;
  const int ro = rec_offset(d, f);
  if (ro >= 0) { *base = (unsigned char*)(h->rec + ro); *pitch = (size_t)d.rec_pitch * sizeof(double); return 0; }
  if (f == B2MJ_F_XFRC_APPLIED) { *base = (unsigned char*)h->xfrc; *pitch = (size_t)6 * m->nbody * sizeof(double); return 0; }
  if (f == B2MJ_F_MOCAP_POS) { *base = (unsigned char*)h->mocap; *pitch = (size_t)7 * m->nmocap * sizeof(double); return 0; }
  if (f == B2MJ_F_MOCAP_QUAT) { *base = (unsigned char*)(h->mocap + 3 * m->nmocap); *pitch = (size_t)7 * m->nmocap * sizeof(double); return 0; }
  if (for_write && f == B2MJ_F_QFRC_PASSIVE && h->in_split_step) {
    // passive hook of a split step (mjcb_passive, mujoco_env.h:247-251): plugins ADD to qfrc_passive; the arena
    // image written by step_begin is what step_end reloads
    *base = (unsigned char*)(h->garena_d + d.off_g[f]);
    *pitch = (size_t)d.arena_g_doubles * sizeof(double);
    return 0;
  }
  if (for_write) { set_error(std::string("field '") + b2mj_field_name((b2mj_field)f) + "' is computed, not settable"); return B2MJ_EINVAL; }
  if (f == B2MJ_F_NCON || f == B2MJ_F_NEFC || f == B2MJ_F_SOLVER_ITER) {
    *base = (unsigned char*)(h->stats + (f == B2MJ_F_NCON ? 0 : f == B2MJ_F_NEFC ? 1 : 2));
    *pitch = 4 * sizeof(int);
    return 0;
  }
  if (f == B2MJ_F_WARNING) { *base = (unsigned char*)h->warning; *pitch = B2MJ_NWARNING * sizeof(int); return 0; }
  if (f == B2MJ_F_EFC_AR) { set_error("efc_AR is never materialised: the CUDA PGS is matrix-free"); return B2MJ_EUNSUPPORTED; }
  if (!h->dump_valid) {
    set_error(std::string("field '") + b2mj_field_name((b2mj_field)f) +
              "' is an intermediate: call b2mj_set_keep_intermediates(h,1) before the step/forward");
    return B2MJ_ESTATE;
  }
  if (is_int) { *base = (unsigned char*)(h->garena_i + d.off_g[f]); *pitch = (size_t)d.arena_g_ints * sizeof(int); }
  else { *base = (unsigned char*)(h->garena_d + d.off_g[f]); *pitch = (size_t)d.arena_g_doubles * sizeof(double); }
  return 0;
}

int b2mj_get(b2mj_handle* hh, b2mj_field f, void* host_dst, size_t bytes) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !host_dst) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  unsigned char* base; size_t pitch, es; int n;
  if (int rc = locate(h, f, false, &base, &pitch, &es, &n)) return rc;
  if (bytes != (size_t)h->nenv * n * es) { set_error("b2mj_get: byte count mismatch"); return B2MJ_EINVAL; }
  if (n == 0) return 0;
  // ordered on the handle's stream (full-speed DMA when host_dst is pinned), then waited for
  CUDA_OK(cudaMemcpy2DAsync(host_dst, n * es, base, pitch, n * es, h->nenv, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int b2mj_set(b2mj_handle* hh, b2mj_field f, const void* host_src, size_t bytes) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !host_src) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  unsigned char* base; size_t pitch, es; int n;
  if (int rc = locate(h, f, true, &base, &pitch, &es, &n)) return rc;
  if (bytes != (size_t)h->nenv * n * es) { set_error("b2mj_set: byte count mismatch"); return B2MJ_EINVAL; }
  if (n == 0) return 0;
  CUDA_OK(cudaMemcpy2DAsync(base, pitch, host_src, n * es, n * es, h->nenv, cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (f == B2MJ_F_XFRC_APPLIED) h->dm.has_xfrc = 1;
  return 0;
}

int b2mj_step_host(b2mj_handle* hh, int nsteps, const double* host_ctrl, double* host_qpos, double* host_qvel,
                   double* host_sensordata) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  if (nsteps <= 0) {
    set_error("b2mj_step_host: nsteps must be positive");
    return B2MJ_EINVAL;
  }
  CUDA_OK(cudaSetDevice(h->device));
  const b2mjModel* m = h->model;
  unsigned char* base; size_t pitch, es; int n;
  // Zero-copy exchange for single steps with pinned (page-locked, mapped) host buffers: the step kernel reads each env's
  // controls straight from host memory when it picks the env up and writes qpos / qvel / sensordata to host memory when
  // the env's step is done, so the transfers ride under the launch (whose length is set by its slowest env) instead
  // of bracketing it as four strided copies.  Pageable buffers, multi-step calls and arena dumps take the copy path.
  auto mapped = [](const void* p) -> void* {
    if (!p) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (at.type == cudaMemoryTypeHost) ? at.devicePointer : nullptr;
  };
  static const bool no_zero_copy = getenv("B2MJ_NO_ZERO_COPY") != nullptr;
  const bool want_ctrl = host_ctrl && m->nu, want_s = host_sensordata && m->nsensordata;
  void* z_ctrl = want_ctrl ? mapped(host_ctrl) : nullptr;
  void* z_qpos = mapped(host_qpos);
  void* z_qvel = mapped(host_qvel);
  void* z_sens = want_s ? mapped(host_sensordata) : nullptr;
  const bool zero_copy = !no_zero_copy && nsteps == 1 && !h->keep_intermediates && (!want_ctrl || z_ctrl) &&
                         (!host_qpos || z_qpos) && (!host_qvel || z_qvel) && (!want_s || z_sens) &&
                         (want_ctrl || host_qpos || host_qvel || want_s);
  h->in_split_step = 0;
  if (zero_copy) {
    if (int rc = handle_launch(h, MODE_STEP, 1, static_cast<const double*>(z_ctrl), static_cast<double*>(z_qpos),
                               static_cast<double*>(z_qvel), static_cast<double*>(z_sens)))
      return rc;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  if (want_ctrl) {
    if (int rc = locate(h, B2MJ_F_CTRL, true, &base, &pitch, &es, &n)) return rc;
    CUDA_OK(cudaMemcpy2DAsync(base, pitch, host_ctrl, n * es, n * es, h->nenv, cudaMemcpyHostToDevice, h->stream));
  }
  if (int rc = handle_launch(h, MODE_STEP, nsteps)) return rc;
  const struct { b2mj_field f; double* dst; } outs[3] = {
      {B2MJ_F_QPOS, host_qpos}, {B2MJ_F_QVEL, host_qvel}, {B2MJ_F_SENSORDATA, host_sensordata}};
  for (const auto& o : outs) {
    if (!o.dst) continue;
    if (int rc = locate(h, o.f, false, &base, &pitch, &es, &n)) return rc;
    if (n == 0) continue;
    CUDA_OK(cudaMemcpy2DAsync(o.dst, n * es, base, pitch, n * es, h->nenv, cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return 0;
}

int b2mj_set_device(b2mj_handle* hh, b2mj_field f, const void* dev_src, size_t src_pitch_elems) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !dev_src) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  unsigned char* base; size_t pitch, es; int n;
  if (int rc = locate(h, f, true, &base, &pitch, &es, &n)) return rc;
  if (n == 0) return 0;
  if (src_pitch_elems == 0) src_pitch_elems = (size_t)n;
  if (src_pitch_elems < (size_t)n) { set_error("b2mj_set_device: source pitch smaller than the field"); return B2MJ_EINVAL; }
  // both element sizes are multiples of 4 bytes: scatter as 32-bit words
  const int words = (int)(n * es / 4);
  const long long total = (long long)h->nenv * words;
  const int threads = 256;
  const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148 * 16);
  scatter_field_kernel<<<blocks, threads, 0, h->stream>>>(reinterpret_cast<unsigned*>(base), (int)(pitch / 4),
                                                          reinterpret_cast<const unsigned*>(dev_src),
                                                          (int)(src_pitch_elems * es / 4), words, h->nenv);
  CUDA_OK(cudaGetLastError());
  h->launches++;
  if (f == B2MJ_F_XFRC_APPLIED) h->dm.has_xfrc = 1;
  return 0;
}

int b2mj_device_ptr(b2mj_handle* hh, b2mj_field f, void** dev_ptr, size_t* pitch_elems) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !dev_ptr) return B2MJ_EINVAL;
  unsigned char* base; size_t pitch, es; int n;
  const int ro = rec_offset(h->dm, f);
  if (ro < 0 && f != B2MJ_F_XFRC_APPLIED && f != B2MJ_F_MOCAP_POS && f != B2MJ_F_MOCAP_QUAT && f != B2MJ_F_WARNING &&
      f != B2MJ_F_NCON && f != B2MJ_F_NEFC && f != B2MJ_F_SOLVER_ITER) {
    set_error("b2mj_device_ptr: only resident state / input / output fields have stable device pointers");
    return B2MJ_EINVAL;
  }
  if (int rc = locate(h, f, false, &base, &pitch, &es, &n)) return rc;
  *dev_ptr = base;
  if (pitch_elems) *pitch_elems = pitch / es;
  if (f == B2MJ_F_XFRC_APPLIED) h->dm.has_xfrc = 1;  // the caller may write through the pointer
  return 0;
}

int b2mj_model_update(b2mj_handle* hh, const b2mjModel* m) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !m) return B2MJ_EINVAL;
  const b2mjModel* o = h->model;
#define X(n) if (m->n != o->n) { set_error("b2mj_model_update: size field '" #n "' changed; create a new handle"); return B2MJ_EINVAL; }
  B2MJ_MODEL_SIZES(X)
#undef X
  if (m->opt.solver != o->opt.solver || m->opt.integrator != o->opt.integrator || m->opt.cone != o->opt.cone) {
    set_error("b2mj_model_update: solver / integrator / cone changes alter the arena layout; create a new handle");
    return B2MJ_EINVAL;
  }
  if (int rc = check_supported(m)) return rc;
  if (model_max_condim(m) * model_max_condim(m) > h->dm.conh_stride) {
    set_error("b2mj_model_update: the edit raises the largest condim of the model (per-contact arrays are sized by it); create a new handle");
    return B2MJ_EINVAL;
  }
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (h->n_env_models) b2mj_set_env_models(hh, nullptr, 0, nullptr);  // a broadcast update replaces per-env variants
  b2mjModel* clone = model_clone(m);
  b2mj_model_free(h->model);
  h->model = clone;
  const int has_xfrc = h->dm.has_xfrc;
  const int was_rnepost = h->dm.need_rnepost, was_subtreevel = h->dm.need_subtreevel, was_dense = h->dm.dense_small;
  if (int rc = upload_model(h)) return rc;
  h->dm.has_xfrc = has_xfrc;
  if (h->dm.need_rnepost != was_rnepost || h->dm.need_subtreevel != was_subtreevel || h->dm.dense_small != was_dense) {
    set_error("b2mj_model_update: the edit changes which per-env arrays exist (sensor types); create a new handle");
    return B2MJ_EINVAL;
  }
  return upload_init_templates(h);
}

int b2mj_register_collision_function(b2mj_handle* hh, int t1, int t2, int collfn) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || t1 < 0 || t2 < 0 || t1 > 7 || t2 > 7 || collfn < B2MJ_COLLFN_DEFAULT || collfn > B2MJ_COLLFN_BOUNDING_SPHERES) {
    set_error("b2mj_register_collision_function: bad geom type or function id");
    return B2MJ_EINVAL;
  }
  if (t1 > t2) std::swap(t1, t2);
  if (collfn == B2MJ_COLLFN_BOUNDING_SPHERES && t2 == B2MJ_GEOM_PLANE) {
    set_error("b2mj_register_collision_function: a plane has no bounding sphere");
    return B2MJ_EINVAL;
  }
  h->dm.collfunc[t1 * 8 + t2] = (unsigned char)collfn;
  h->dm_env.collfunc[t1 * 8 + t2] = (unsigned char)collfn;
  return 0;
}

int b2mj_reset_collision_functions(b2mj_handle* hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  std::memset(h->dm.collfunc, 0, sizeof(h->dm.collfunc));
  std::memset(h->dm_env.collfunc, 0, sizeof(h->dm_env.collfunc));
  return 0;
}

int b2mj_set_env_models(b2mj_handle* hh, const b2mjModel* const* models, int nmodels, const int* env_model) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || nmodels < 0 || (nmodels > 0 && (!models || !env_model))) {
    set_error("b2mj_set_env_models: bad argument");
    return B2MJ_EINVAL;
  }
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  cudaFree(h->env_blob);
  cudaFree(h->env_model_idx);
  h->env_blob = nullptr;
  h->env_model_idx = nullptr;
  h->n_env_models = 0;
  if (nmodels == 0) return 0;
  const b2mjModel* o = h->model;
  for (int v = 0; v < nmodels; v++) {
    const b2mjModel* m = models[v];
    if (!m) { set_error("b2mj_set_env_models: null model"); return B2MJ_EINVAL; }
#define X(n) if (m->n != o->n) { set_error("b2mj_set_env_models: variant " + std::to_string(v) + " differs in size field '" #n "'"); return B2MJ_EINVAL; }
    B2MJ_MODEL_SIZES(X)
#undef X
    if (m->opt.solver != o->opt.solver || m->opt.integrator != o->opt.integrator || m->opt.cone != o->opt.cone ||
        m->opt.timestep != o->opt.timestep || m->opt.disableflags != o->opt.disableflags) {
      set_error("b2mj_set_env_models: variants must share solver / integrator / cone / timestep / flags");
      return B2MJ_EINVAL;
    }
    if (int rc = check_supported(m)) return rc;
    if (model_max_condim(m) * model_max_condim(m) > h->dm.conh_stride) {
      set_error("b2mj_set_env_models: variant " + std::to_string(v) + " raises the largest condim of the model; create the handle from the variant with the largest condim");
      return B2MJ_EINVAL;
    }
    // topology (parents, addresses, types, pair table) and the reset pose are shared: only parameters may differ
    const bool same_topology =
        !std::memcmp(m->body_parentid, o->body_parentid, sizeof(int) * o->nbody) &&
        !std::memcmp(m->jnt_type, o->jnt_type, sizeof(int) * o->njnt) && !std::memcmp(m->dof_parentid, o->dof_parentid, sizeof(int) * o->nv) &&
        !std::memcmp(m->geom_type, o->geom_type, sizeof(int) * o->ngeom) && !std::memcmp(m->geom_bodyid, o->geom_bodyid, sizeof(int) * o->ngeom) &&
        !std::memcmp(m->collpair_geom1, o->collpair_geom1, sizeof(int) * o->ncollpair) &&
        !std::memcmp(m->collpair_geom2, o->collpair_geom2, sizeof(int) * o->ncollpair) &&
        !std::memcmp(m->qpos0, o->qpos0, sizeof(double) * o->nq);
    if (!same_topology) {
      set_error("b2mj_set_env_models: variant " + std::to_string(v) + " changes topology or qpos0; only parameters may differ");
      return B2MJ_EINVAL;
    }
    int damp = 0;
    for (int i = 0; i < m->nv; i++) damp |= m->dof_damping[i] > 0;
    (void)damp;  // XF_HINV is reserved whether or not a variant has damping (make_layout)
  }
  std::vector<int> idx(env_model, env_model + h->nenv);
  for (int e = 0; e < h->nenv; e++)
    if (idx[e] < 0 || idx[e] >= nmodels) { set_error("b2mj_set_env_models: env_model index out of range"); return B2MJ_EINVAL; }
  const size_t stride = variant_region_bytes(o);
  std::vector<unsigned char> host(stride * nmodels, 0);
  CUDA_OK(cudaMalloc(&h->env_blob, host.size()));
  h->dm_env = h->dm;
  for (int v = 0; v < nmodels; v++)
    pack_variant(models[v], host.data() + stride * v, v == 0 ? &h->dm_env : nullptr, (unsigned char*)h->env_blob);
  h->dm_env.env_model_stride = (long long)stride;
  h->dm_env.any_damping = 0;
  for (int v = 0; v < nmodels; v++)
    for (int i = 0; i < o->nv; i++) h->dm_env.any_damping |= models[v]->dof_damping[i] > 0;
  CUDA_OK(cudaMemcpy(h->env_blob, host.data(), host.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&h->env_model_idx, sizeof(int) * h->nenv));
  CUDA_OK(cudaMemcpy(h->env_model_idx, idx.data(), sizeof(int) * h->nenv, cudaMemcpyHostToDevice));
  h->n_env_models = nmodels;
  return 0;
}

int b2mj_stage_profile(b2mj_handle* hh, int enable, uint64_t* cycles, int ncycles) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (cycles && h->prof) {
    unsigned long long host[PROF_COUNT];
    CUDA_OK(cudaMemcpy(host, h->prof, sizeof(host), cudaMemcpyDeviceToHost));
    for (int i = 0; i < ncycles; i++) cycles[i] = i < PROF_COUNT ? host[i] : 0;
  } else if (cycles) {
    for (int i = 0; i < ncycles; i++) cycles[i] = 0;
  }
  if (enable && !h->prof) CUDA_OK(cudaMalloc(&h->prof, PROF_COUNT * sizeof(unsigned long long)));
  if (enable) CUDA_OK(cudaMemset(h->prof, 0, PROF_COUNT * sizeof(unsigned long long)));
  if (!enable && h->prof) { cudaFree(h->prof); h->prof = nullptr; }
  return PROF_COUNT;
}

int b2mj_env_cycles(b2mj_handle* hh, int* host_kcycles) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !host_kcycles) return B2MJ_EINVAL;
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaMemcpy2D(host_kcycles, sizeof(int), h->stats + 3, 4 * sizeof(int), sizeof(int), h->nenv, cudaMemcpyDeviceToHost));
  return 0;
}

const char* b2mj_stage_name(int stage) {
  static const char* names[PROF_COUNT] = {"load", "kinematics", "comPos", "tendon_transmission", "crb_factorM", "collision",
                                          "makeConstraint", "projectConstraint", "sensorPos", "velocity_head", "comVel",
                                          "passive", "referenceConstraint", "rne_bias", "sensorVel", "actuation",
                                          "acceleration", "solve", "sensorAcc", "integrate", "store"};
  return stage >= 0 && stage < PROF_COUNT ? names[stage] : nullptr;
}

int b2mj_launch_info(b2mj_handle* hh, b2mjLaunchInfo* out) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !out) return B2MJ_EINVAL;
  const DevModel& d = h->dm;
  // the shape of the most recent launch (per-step launches of long batches and fused rollouts use the wide lock-stepped
  // CTA), else the per-step shape chosen at create time
  const int W = h->last_warps_per_cta > 0 ? h->last_warps_per_cta : h->warps_per_cta;
  out->warps_per_cta = W;
  out->ctas = (h->nenv + W - 1) / W;
  out->smem_bytes_per_cta = (int)(h->smem_bytes / h->warps_per_cta * W);
  out->arena_doubles_per_env = d.arena_g_doubles;
  out->arena_in_smem = h->arena_in_smem;
  // bytes the fused step moves per env: load A+B, store B+C
  out->state_record_bytes = ((d.rec_C_begin - d.rec_A_begin) + (d.rec_end - d.rec_B_begin)) * 8;
  int regs = 0;
  b2k_step_kernel_attrs(&regs, nullptr, nullptr);
  out->regs_per_thread = regs;
  out->launches = h->launches;
  return 0;
}

}  // extern "C"
