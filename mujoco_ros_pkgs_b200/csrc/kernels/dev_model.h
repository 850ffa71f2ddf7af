// dev_model.h — device-side view of a compiled model + the per-env arena layout.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//  * model constants: one blob per handle, shared by all envs (read through the read-only path);
//  * state record: rec[nenv][rec_pitch] doubles, per env
//        [ ctrl | qfrc_applied | pad ][ qpos | qvel | act | qacc_warmstart | time | pad ][ qacc | sensordata | act_dot | pad ]
//          \__________ segment A (inputs) ______/\______________ segment B (state) ___________/\_______ segment C (outputs) ___/
//    A+B is one contiguous 16B-aligned span (one bulk load per env), B+C likewise (one bulk store);
//  * arena: every other mjData field of one env (b2mj_field order).  The "hot" part is shared-memory
//    resident for the duration of a launch, the "cold" part (big constraint arrays when they do not
//    fit) lives in garena[nenv][arena_g_doubles] in HBM/L2.  With keep_intermediates the hot part is
//    dumped to garena at the end of the launch so b2mj_get can read any field.
#pragma once
#include <stdint.h>

#include "b2mj.h"

// lanes per env: one env per warp.  A 16-lane variant (two envs per warp) was measured in round 1 (profiles/
// r1_layout_sweep.txt): it wins only while the batch is contact free and loses 8-13 % on the BASELINE workload because
// the two envs of a warp serialise each other's solver iterations.  It is not a supported configuration: the
// owner-computes PGS, the register Gauss-Jordan and the Cholesky register windows assume 32 lanes.
#ifdef B2K_G
#error "B2K_G is not a build knob any more: the step kernel is written for one env per 32-lane warp"
#endif
#define B2K_G 32

// The step kernel translation unit is compiled twice: once as is (namespace b2k: every env shares one model) and once
// with -DB2K_PER_ENV_MODEL (namespace b2k_em): there every model array is read through a per-env byte offset, so that
// each env may run its own VARIANT of the model (domain randomisation / the reference's mutating services applied to
// one env: set_body_state mass, set_geom_properties, set_gravity, set_equality_constraint_parameters --
// mujoco_ros/src/callbacks.cpp:210-370, 462-738).  Same sizes and topology, different values.  The shared-model build
// pays nothing for it: MArr below is then a bare pointer.
#ifdef B2K_PER_ENV_MODEL
#define b2k b2k_em
#endif

namespace b2k {

#if defined(__CUDACC__) && defined(B2K_PER_ENV_MODEL)
__device__ __forceinline__ long long env_model_offset();  // env_ctx.cuh: byte offset of this warp's model variant
#endif

// model array as the device sees it (pointer-sized in both builds, so DevModel has ONE layout for the host)
template <class T>
struct MArr {
  const T* p;
  __host__ __device__ MArr& operator=(const T* q) { p = q; return *this; }
  __host__ __device__ __forceinline__ const T* get() const {
#if defined(__CUDA_ARCH__) && defined(B2K_PER_ENV_MODEL)
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(p) + env_model_offset());
#else
    return p;
#endif
  }
  __host__ __device__ __forceinline__ operator const T*() const { return get(); }
  __host__ __device__ __forceinline__ const T& operator[](int i) const { return get()[i]; }
  __host__ __device__ __forceinline__ const T& operator[](unsigned i) const { return get()[i]; }
  __host__ __device__ __forceinline__ const T& operator[](long long i) const { return get()[i]; }
  __host__ __device__ __forceinline__ const T& operator[](size_t i) const { return get()[i]; }
  __host__ __device__ __forceinline__ const T* operator+(int i) const { return get() + i; }
  __host__ __device__ __forceinline__ const T* operator+(size_t i) const { return get() + i; }
};

// extra per-env scratch arrays that are not b2mj_field entries
enum XField {
  XF_QLOC = 0,      // 4*njnt  local joint quaternions (hinge: axis-angle; ball: qpos)
  XF_QH,            // nM      M + h*diag(damping), factorised
  XF_QHDIAGINV,     // nv
  XF_EFC_MINVJT,    // njmax*nv rows of inv(M) J'   (matrix-free PGS / Newton helpers)
  XF_EFC_ARDIAG,    // njmax   diagonal of AR = J inv(M) J' + R
  XF_VEC0,          // nv scratch vectors
  XF_VEC1, XF_VEC2, XF_VEC3, XF_VEC4, XF_VEC5,
  XF_EFC_JAREF,     // njmax
  XF_EFC_JV,        // njmax
  XF_EFC_QUAD,      // 3*njmax
  XF_NEWTON_H,      // nv*ldh; team mode: packed lower triangle, nv (nv + 1) / 2
  XF_CONTACT_H,     // 36*nconmax
  XF_SUBTREE_LINVEL,// 3*nbody
  XF_SUBTREE_ANGMOM,// 3*nbody
  XF_BODYVEL,       // 6*nbody
  XF_RK_X0,         // nq+nv+na   RK4 saved state
  XF_RK_XF,         // 4*nv       RK4 stage velocities
  XF_RK_F,          // 4*(nv+na)  RK4 stage accelerations / act_dot
  XF_RK_DX,         // 2*nv+na
  XF_SCRATCH,       // stage-local scratch: max(14*nbody kinematic scan ping-pong, 6*nv crb*cdof, 6*nbody RNE forces)
  XF_QW,            // nM   off-diagonal entries of inv(L) for qLD (sparse layout of qM)
  XF_QHW,           // nM   same for qH
  XF_EFC_AR,        // njmax*njmax  dense AR = J inv(M) J' + R (PGS); always in the HBM/L2 arena
  XF_MINV,          // nv*nv  dense inv(qM) (small models: nv <= 16)
  XF_HINV,          // nv*nv  dense inv(qM + h diag(damping))
  XF_PRIMAL,        // 8*nv  Newton / CG work vectors (Ma, Mv, grad, Mgrad, search, gradold, Mgradold, invdiag)
  XF_EFC_AR_S,      // shared-memory home of AR when nefc*nefc fits (the common case)
  XF_JWIN,          // primal solvers: shared-memory window for the ACTIVE rows of efc_J: [0] = valid flag, rows from +2
  XF_JCOLS,         // team mode: per constraint row {nnz, columns} bytes (team.cuh), njmax * 17 bytes
  XF_TRI,           // pair table of a row-major lower triangle of order nv, packed a | b << 8 in 16 bits (cholPairTable)
  XF_IMPL_LU,       // implicit integrator: nv*nv dense M - h qDeriv and its LU factors (always in the HBM/L2 arena)
  XF_IMPL_D,        // implicit integrator: 6*nbody*nv body-force derivatives d cfrc / d qvel (always in the HBM/L2 arena)
  XF_JVALS,         // team mode: the non-zero entries of the first jvals_rows rows of efc_J, [row][16] aligned with XF_JCOLS
  XF_COUNT
};

struct DevModel {
#define B2K_X_SIZE(n) int n;
  B2MJ_MODEL_SIZES(B2K_X_SIZE)
#undef B2K_X_SIZE
  int nmaskword;  // words per body of body_dofmask
  b2mjOption opt;
  double meaninertia;
#define B2K_X_ARR(t, n, r, c) MArr<t> n;
  B2MJ_MODEL_ARRAYS(B2K_X_ARR)
#undef B2K_X_ARR
  MArr<double> env_gravity;   // [3] opt.gravity of the model variant
  MArr<double> env_scalars;   // [1] stat.meaninertia of the model variant
  long long env_model_stride; // bytes between model variants in the blob, 0 = one shared model
  const unsigned* body_dofmask;  // [nbody][nmaskword]: bit k set if dof k is on the chain from the body to its root
  // ---- derived topology tables (host-built, handle.cu::upload_model) for the wide-parallel stages ----
  int nbodyword;                 // words per body of body_submask
  int njump;                     // pointer-jumping rounds of the kinematic scan = ceil(log2(max body level))
  int ndoflevel;                 // number of dof depth levels (max #ancestors + 1)
  const unsigned* body_submask;  // [nbody][nbodyword]: bit i set if body i is in the subtree of the body (incl. itself)
  const int* body_jump;          // [njump][nbody]: ancestor 2^r levels up, 0 (world) if there is none
  const int* M_row;              // [nM] dof i of sparse inertia entry t  (entry t = M(i, j), j = i or an ancestor of i)
  const int* M_col;              // [nM] dof j
  const int* M_ancadr;           // [nM] dof_Madr[j]: start of the ancestor's own row
  const int* dof_nanc;           // [nv] number of ancestors of the dof (row length - 1)
  const unsigned* dof_premask;   // [nv][nmaskword]: dofs whose velocity is accumulated before this dof's joint (cdof_dot)
  const int* doflevel_adr;       // [ndoflevel+1] offsets into doflevel_dof
  const int* doflevel_dof;       // [nv] dofs sorted by depth
  const int* dof_descadr;        // [nv+1] CSR of strict descendants of a dof
  const int* dof_desc_dof;       // [nM-nv] descendant dof m
  const int* dof_desc_adr;       // [nM-nv] sparse address of entry (m, k)

  // ---- arena layout (element offsets; doubles for f64 fields, ints for i32 fields) ----
  int off_g[B2MJ_NFIELD];   // offset in the full per-env arena (always valid)
  int off_s[B2MJ_NFIELD];   // offset in the shared-memory arena, -1 if the field is cold (global only)
  int xoff_g[XF_COUNT];
  int xoff_s[XF_COUNT];
  int fsize[B2MJ_NFIELD];   // per-env element count of each field
  unsigned char fis_int[B2MJ_NFIELD];
  int xsize[XF_COUNT];
  int ar_ovl_off, ar_ovl_doubles;     // PGS: AR may overlay [xpos .. crb] of the shared arena (0 doubles = not allowed)
  int arena_g_doubles, arena_g_ints;  // full arena sizes per env
  int arena_s_doubles, arena_s_ints;  // shared-memory resident sizes per env
  // ---- state record layout (doubles) ----
  int rec_pitch;
  int rec_ctrl, rec_qfrc_applied, rec_qpos, rec_qvel, rec_act, rec_warm, rec_time, rec_qacc, rec_sensordata, rec_act_dot;
  int rec_A_begin, rec_B_begin, rec_C_begin, rec_end;  // 2-double aligned segment starts
  // flags
  int has_xfrc;            // xfrc_applied / mocap arrays are read (plugin-visible surface enabled)
  int need_rnepost;        // some sensor needs cacc / cfrc_int
  int need_subtreevel;
  int any_damping;         // Euler implicit damping active
  int dense_small;         // nv <= 16: inertia handled as dense nv x nv matrices (explicit inverses, no index tables)
  unsigned char collfunc[64];  // narrowphase override per geom-type pair [t1 * 8 + t2]: B2MJ_COLLFN_* (0 = built in)
  int team_warps;          // warps per env: 1, or 8 for wide Newton models (team.cuh): one env per CTA, helpers on call
  int jwin_rows;           // rows the efc_J window holds (0 = no window)
  int jvals_rows;          // team mode: rows of efc_J whose non-zero entries are mirrored in shared memory (XF_JVALS)
  int ldh;                 // leading dimension of the Newton Hessian (odd in team mode: conflict-free column walks)
  int conh_stride;         // doubles per contact in XF_CONTACT_H: (largest condim of the model)^2, not 36 (make_layout)
};

// fused publish (b2mj_step_publish): where the finished env's row goes in every rank's gathered slab
#define B2K_PUB_MAX_RANKS 16
#define B2K_PUB_MAX_FIELDS 8
struct PubArgs {
  int nranks, rank, count, nfields;
  int foff[B2K_PUB_MAX_FIELDS], fcnt[B2K_PUB_MAX_FIELDS];  // record offsets / lengths of the published fields
  double* slab[B2K_PUB_MAX_RANKS];                          // [world][nenv][count] in rank r's memory (peer mapped)
  int* flags[B2K_PUB_MAX_RANKS];                            // [world] sequence flags in rank r's memory (peer mapped)
  unsigned* done;                                           // local counter of envs whose rows are out (self-resetting)
};

struct LaunchArgs {
  double* rec;             // [nenv][rec_pitch]
  double* garena_d;        // [nenv][arena_g_doubles]
  int* garena_i;           // [nenv][arena_g_ints]
  const double* xfrc;      // [nenv][6*nbody] or null
  const double* mocap;     // [nenv][7*nmocap] (pos then quat) or null
  int* warning;            // [nenv][B2MJ_NWARNING] cumulative
  int* stats;              // [nenv][4]: ncon, nefc, solver_iter, reserved
  int nenv;
  int nsteps;
  int mode;                // 0 step, 1 forward only, 2 step_begin (to control hook), 3 step_end
  int dump;                // copy the shared arena to garena at the end
  int rk_stage;            // MODE_STEP_END of an RK4 model: which sub-step's second half this launch runs (0..3)
  // fused rollout (b2mj_rollout): per-step control stream in, per-step trajectory out (all optional)
  const double* ctrl_seq;  // [nsteps][nenv][nu]
  double* traj_qpos;       // [nsteps][nenv][nq]
  double* traj_qvel;       // [nsteps][nenv][nv]
  double* traj_sensor;     // [nsteps][nenv][nsensordata]
  // persistent rollout scheduling: warps draw (env, chunk-of-steps) tickets from a global counter so that
  // slow envs (long solver runs) do not leave the rest of the GPU idle; sched[0] = next ticket,
  // sched[1 + env] = chunks of that env already completed.  null = one env per warp, all steps.
  int* sched;
  int chunk;               // steps per ticket
  int sync_stages;         // CTA-wide lockstep at stage boundaries (instruction / constant cache locality)
  const int* perm;         // [nenv] launch slot -> env, heaviest envs first (b2k_order_kernel), or null = identity
  int* cost;               // [nenv] this launch's per-env residency (1024-cycle units) for the asynchronous order refresh, or null
  const int* env_model;    // [nenv] model variant of each env (per-env-model build only), or null
  const PubArgs* pub;      // fused publish targets (device memory), or null
  int pub_seq;             // sequence number this launch raises in every rank's flag array once all rows are out
  unsigned long long* prof; // [PROF_COUNT] per-stage SM-cycle totals over all envs, or null (b2mj_stage_profile)
};

// stage ids of the optional cycle profile
enum ProfStage {
  PROF_LOAD = 0, PROF_KINEMATICS, PROF_COMPOS, PROF_TENDON, PROF_CRB_FACTOR, PROF_COLLISION, PROF_MAKECONSTRAINT,
  PROF_PROJECT, PROF_SENSORPOS, PROF_VELHEAD, PROF_COMVEL, PROF_PASSIVE, PROF_REFCONSTRAINT, PROF_RNE, PROF_SENSORVEL,
  PROF_ACTUATION, PROF_ACCELERATION, PROF_SOLVE, PROF_SENSORACC, PROF_INTEGRATE, PROF_STORE, PROF_COUNT
};

enum { MODE_STEP = 0, MODE_FORWARD = 1, MODE_STEP_BEGIN = 2, MODE_STEP_END = 3 };

}  // namespace b2k
