// step_kernel.cu — the fused batched physics step for sm_100a: one env per warp, the env's mjData
// arena resident in shared memory, state record moved HBM<->SMEM with 1-D TMA bulk copies.
//
// Replaces the `mj_step(model_.get(), data_.get())` call of the reference's physics loop
// (mujoco_ros/src/mujoco_env.cpp:498,552,593) for a whole batch of independent envs; also
// mj_forward (:329,:621) and the two halves around the control hook (mjcb_control placement,
// mujoco_env.h:242-246).  Pipeline order follows SURVEY.md Appendix A.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "dev_model.h"
#include "env_ctx.cuh"
#include "stages_collision.cuh"
#include "stages_constraint.cuh"
#include "stages_sensor.cuh"
#include "stages_implicit.cuh"
#include "stages_smooth.cuh"
#include "stages_solver.cuh"
#include "step_launch.h"

namespace b2k {

__device__ __forceinline__ bool isBad(double x) { return isnan(x) || x > B2MJ_MAXVAL || x < -B2MJ_MAXVAL; }

// mj_resetData on the resident state of one env
__device__ __noinline__ void resetEnv(const Env e, int* warning, int which) {
  const DevModel& m = c_dm;
  double* qpos = e.D(B2MJ_F_QPOS);
  FORL(i, m.nq) qpos[i] = m.qpos0[i];
  double* qvel = e.D(B2MJ_F_QVEL);
  double* warm = e.D(B2MJ_F_QACC_WARMSTART);
  double* qap = e.D(B2MJ_F_QFRC_APPLIED);
  double* qacc = e.D(B2MJ_F_QACC);
  FORL(i, m.nv) { qvel[i] = 0; warm[i] = 0; qap[i] = 0; qacc[i] = 0; }
  if (m.na) { double* act = e.D(B2MJ_F_ACT); FORL(i, m.na) act[i] = 0; }
  if (m.nu) { double* ctrl = e.D(B2MJ_F_CTRL); FORL(i, m.nu) ctrl[i] = 0; }
  if (m.nsensordata) { double* sd = e.D(B2MJ_F_SENSORDATA); FORL(i, m.nsensordata) sd[i] = 0; }
  if (e.lane == 0) {
    e.D(B2MJ_F_TIME)[0] = 0;
    const int w = warning[which] + 1;
    for (int k = 0; k < B2MJ_NWARNING; k++) warning[k] = 0;
    warning[which] = w;
  }
  WSYNC();
}

struct StepCtx {
  int ncon, nefc, iters;
  int nsync;         // threads taking part in the stage barriers (0 = no barriers in this pass)
  long long t_prev;  // stage profile (only touched when LaunchArgs::prof is set)
};

// Stage boundary.  All warps of the CTA run the same model, so keeping them in lockstep at stage
// granularity makes them share instruction-cache lines and model-constant cache lines (the megakernel is
// fetch-latency bound, not issue bound).  The barrier is a locality device, not a correctness one; it
// counts only the warps that own an env (the last CTA may be partial).
#define STAGE_SYNC() \
  if (sc.nsync) asm volatile("bar.sync 1, %0;" ::"r"(sc.nsync) : "memory")

#define PROF_MARK(id)                                                              \
  STAGE_SYNC();                                                                    \
  if (a.prof) {                                                                    \
    const long long _now = clock64();                                              \
    if (e.lane == 0) atomicAdd(a.prof + (id), (unsigned long long)(_now - sc.t_prev)); \
    sc.t_prev = _now;                                                              \
  }

// mj_forwardSkip split at the control hook
__device__ __noinline__ void forwardPass(const Env e, const LaunchArgs& a, int env, StepCtx& sc, bool skipsensor, bool first_half,
                            bool second_half) {
  const DevModel& m = c_dm;
  int* warning = a.warning + (size_t)env * B2MJ_NWARNING;
  const double* xfrc = (m.has_xfrc && a.xfrc) ? a.xfrc + (size_t)env * 6 * m.nbody : nullptr;
  if (first_half) {
    stage_kinematics(e); PROF_MARK(PROF_KINEMATICS)
    stage_comPos(e); PROF_MARK(PROF_COMPOS)
    stage_tendon_transmission(e); PROF_MARK(PROF_TENDON)
    stage_crb_factor(e, e.dump != 0); PROF_MARK(PROF_CRB_FACTOR)
    sc.ncon = stage_collision(e, warning); PROF_MARK(PROF_COLLISION)
    sc.nefc = stage_makeConstraint(e, sc.ncon, warning); PROF_MARK(PROF_MAKECONSTRAINT)
    PROF_MARK(PROF_PROJECT)  // mj_projectConstraint runs at the head of the solve stage (stage_fwdConstraint)
    if (!skipsensor) stage_sensorPos(e, sc.nefc);
    PROF_MARK(PROF_SENSORPOS)
    stage_velocity_head(e); PROF_MARK(PROF_VELHEAD)
    stage_comVel(e); PROF_MARK(PROF_COMVEL)
    stage_passive(e); PROF_MARK(PROF_PASSIVE)
    stage_referenceConstraint(e, sc.nefc); PROF_MARK(PROF_REFCONSTRAINT)
    stage_rne_bias(e); PROF_MARK(PROF_RNE)
    if (!skipsensor) stage_sensorVel(e, sc.nefc);
    PROF_MARK(PROF_SENSORVEL)
  }
  if (second_half) {
    stage_actuation(e, warning); PROF_MARK(PROF_ACTUATION)
    stage_acceleration(e, xfrc); PROF_MARK(PROF_ACCELERATION)
    sc.iters = stage_fwdConstraint(e, sc.nefc, sc.ncon); PROF_MARK(PROF_SOLVE)
    if (!skipsensor) stage_sensorAcc(e, sc.nefc, sc.ncon, xfrc);
    PROF_MARK(PROF_SENSORACC)
  }
}

// mj_RungeKutta(4), in resumable pieces: the fused step runs them back to back around full forward passes; the split
// step (b2mj_step_begin / b2mj_step_end with the host's control hook in between) runs one piece per launch, so that
// mjcb_control / mjcb_passive fire in EVERY forward pass of a step as they do in the reference (plugin_utils.h:89-105:
// "called ... in every sub-step of RK4").  State that must survive between launches lives in the arena (XF_RK_*),
// which the split modes dump to / reload from HBM; X0's last slot carries the step's start time.
struct RkPtrs {
  double *qpos, *qvel, *qacc, *act, *act_dot, *timep, *X0, *Xf, *F, *dX;
  int nq, nv, na;
};
__device__ __forceinline__ RkPtrs rkPtrs(const Env e) {
  const DevModel& m = c_dm;
  RkPtrs r;
  r.nq = m.nq; r.nv = m.nv; r.na = m.na;
  r.qpos = e.D(B2MJ_F_QPOS); r.qvel = e.D(B2MJ_F_QVEL); r.qacc = e.D(B2MJ_F_QACC);
  r.act = m.na ? e.D(B2MJ_F_ACT) : nullptr;
  r.act_dot = m.na ? e.D(B2MJ_F_ACT_DOT) : nullptr;
  r.timep = e.D(B2MJ_F_TIME);
  r.X0 = e.X(XF_RK_X0);   // nq + nv + na + 1 (start time)
  r.Xf = e.X(XF_RK_XF);   // 4 * nv   stage velocities
  r.F = e.X(XF_RK_F);     // 4 * (nv + na) stage accelerations / act_dot
  r.dX = e.X(XF_RK_DX);   // 2*nv + na
  return r;
}
// after the first forward pass: save the start state and stage 0
__device__ __noinline__ void rk_init(const Env e) {
  const RkPtrs r = rkPtrs(e);
  const int nq = r.nq, nv = r.nv, na = r.na;
  FORL(i, nq) r.X0[i] = r.qpos[i];
  FORL(i, nv) { r.X0[nq + i] = r.qvel[i]; r.Xf[i] = r.qvel[i]; r.F[i] = r.qacc[i]; }
  FORL(i, na) { r.X0[nq + nv + i] = r.act[i]; r.F[nv + i] = r.act_dot[i]; }
  if (e.lane == 0) r.X0[nq + nv + na] = r.timep[0];
  WSYNC();
}
// state of stage s (1..3) from the start state and the stages recorded so far
__device__ __noinline__ void rk_setup_stage(const Env e, int s) {
  const RkPtrs r = rkPtrs(e);
  const int nq = r.nq, nv = r.nv, na = r.na;
  const double h = c_dm.opt.timestep;
  const double A[9] = {0.5, 0, 0, 0, 0.5, 0, 0, 0, 1};
  FORL(k, nv) {
    double dv = 0, da = 0;
    for (int j = 0; j < 3; j++) {
      const double c = A[(s - 1) * 3 + j];
      if (c == 0) continue;
      dv += c * r.Xf[j * nv + k];
      da += c * r.F[j * (nv + na) + k];
    }
    r.dX[k] = dv;
    r.dX[nv + k] = da;
  }
  FORL(k, na) {
    double d = 0;
    for (int j = 0; j < 3; j++) {
      const double c = A[(s - 1) * 3 + j];
      if (c != 0) d += c * r.F[j * (nv + na) + nv + k];
    }
    r.dX[2 * nv + k] = d;
  }
  FORL(i, nq) r.qpos[i] = r.X0[i];
  WSYNC();
  integratePos_warp(e, r.qpos, r.dX, h);
  FORL(k, nv) r.qvel[k] = r.X0[nq + k] + h * r.dX[nv + k];
  FORL(k, na) r.act[k] = r.X0[nq + nv + k] + h * r.dX[2 * nv + k];
  if (e.lane == 0) r.timep[0] = r.X0[nq + nv + na] + (s == 3 ? 1.0 : 0.5) * h;
  WSYNC();
}
// after the forward pass of stage s: record its velocity / acceleration
__device__ __noinline__ void rk_record(const Env e, int s) {
  const RkPtrs r = rkPtrs(e);
  const int nv = r.nv, na = r.na;
  FORL(k, nv) { r.Xf[s * nv + k] = r.qvel[k]; r.F[s * (nv + na) + k] = r.qacc[k]; }
  FORL(k, na) r.F[s * (nv + na) + nv + k] = r.act_dot[k];
  WSYNC();
}
// Butcher combination and the final advance from the start state
__device__ __noinline__ void rk_finish(const Env e) {
  const RkPtrs r = rkPtrs(e);
  const int nq = r.nq, nv = r.nv, na = r.na;
  const double Bw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
  FORL(k, nv) {
    double dv = 0, da = 0;
    for (int j = 0; j < 4; j++) { dv += Bw[j] * r.Xf[j * nv + k]; da += Bw[j] * r.F[j * (nv + na) + k]; }
    r.dX[k] = dv;
    r.dX[nv + k] = da;
  }
  FORL(k, na) {
    double d = 0;
    for (int j = 0; j < 4; j++) d += Bw[j] * r.F[j * (nv + na) + nv + k];
    r.dX[2 * nv + k] = d;
  }
  FORL(i, nq) r.qpos[i] = r.X0[i];
  FORL(i, nv) r.qvel[i] = r.X0[nq + i];
  FORL(i, na) r.act[i] = r.X0[nq + nv + i];
  if (e.lane == 0) r.timep[0] = r.X0[nq + nv + na];
  WSYNC();
  advance_warp(e, r.dX + 2 * nv, r.dX + nv, r.dX);
}

__device__ __noinline__ void stage_rk4(const Env e, const LaunchArgs& a, int env, StepCtx& sc) {
  rk_init(e);
  for (int s = 1; s < 4; s++) {
    rk_setup_stage(e, s);
    forwardPass(e, a, env, sc, true, true, true);
    rk_record(e, s);
  }
  rk_finish(e);
}

__global__ void __launch_bounds__(B2K_MAX_THREADS, B2K_MIN_CTAS) b2k_step_kernel(const __grid_constant__ LaunchArgs a) {
  const DevModel& m = c_dm;
  unsigned char* const smem_raw = b2k_smem;
  // "warp" below = env slot of the CTA.  Normally one warp per env; in team mode (team.cuh) team_warps warps share a
  // slot: warp 0 of the team runs the step, the others serve team calls.
  const int tw = m.team_warps;
  const int warp = (threadIdx.x / B2K_G) / tw, lane = threadIdx.x % B2K_G, nwarp = (blockDim.x / B2K_G) / tw;
  const bool team_helper = ((threadIdx.x / B2K_G) % tw) != 0;
  const unsigned gmask = 0xffffffffu;

  // shared layout: [nwarp slot headers (mbarrier; + team control block in team mode)][nwarp env blocks: doubles | ints]
  const size_t env_bytes = (((size_t)m.arena_s_doubles * 8 + (size_t)m.arena_s_ints * 4) + 15) & ~(size_t)15;
  const int hdr = tw > 1 ? B2K_TEAM_HDR : 16;
  unsigned char* base = smem_raw + hdr * nwarp + (size_t)warp * env_bytes;
  double* sd = reinterpret_cast<double*>(base);
  int* si = reinterpret_cast<int*>(base + (size_t)m.arena_s_doubles * 8);
  const unsigned bar = smem_u32(smem_raw + hdr * warp);
  unsigned parity = 0;
  bool bar_ready = false;
  const int nchunks = a.sched ? (a.nsteps + a.chunk - 1) / a.chunk : 1;

  for (;;) {
    // ---- pick the work item: (env, [step0, step1)) ----
    int env, step0, step1, chunk_id = 0;
    if (a.sched) {
      int t = 0;
      if (lane == 0) t = atomicAdd(a.sched, 1);
      t = __shfl_sync(gmask, t, 0, B2K_G);
      if (t >= a.nenv * nchunks) break;
      chunk_id = t / a.nenv;           // chunk-major: all envs advance together
      env = t - chunk_id * a.nenv;
      step0 = chunk_id * a.chunk;
      step1 = min(step0 + a.chunk, a.nsteps);
      if (lane == 0) {
        const volatile int* done = a.sched + 1 + env;
        while (*done < chunk_id) __nanosleep(256);
        __threadfence();
      }
      __syncwarp(gmask);
    } else {
      env = blockIdx.x * nwarp + warp;
      if (env >= a.nenv) return;  // whole warp exits; no CTA-wide barrier is used below
      if (a.perm) env = a.perm[env];  // heaviest envs occupy the first launch slots
      step0 = 0;
      step1 = (a.mode == MODE_STEP) ? a.nsteps : 1;
    }
#ifdef B2K_PER_ENV_MODEL
    // this env's model variant: byte offset into the variant blob, parked in the spare half of the mbarrier slot
    if (lane == 0 && !team_helper)
      *reinterpret_cast<long long*>(smem_raw + hdr * warp + 8) = a.env_model ? (long long)a.env_model[env] * m.env_model_stride : 0;
    __syncwarp(gmask);
#endif
    double* gd = a.garena_d + (size_t)env * m.arena_g_doubles;
    int* gi = a.garena_i + (size_t)env * m.arena_g_ints;
    Env e{(unsigned)(base - smem_raw), (unsigned)(base - smem_raw) + 8u * (unsigned)m.arena_s_doubles, gd, gi, lane,
          gmask, (a.dump != 0 || a.mode == MODE_STEP_BEGIN || a.mode == MODE_STEP_END) ? 1 : 0};
    if (team_helper) {  // team mode: this warp only serves the main warp's team calls for this env
      team_worker(e);
      return;
    }
    int* warning = a.warning + (size_t)env * B2MJ_NWARNING;
    double* rec = a.rec + (size_t)env * m.rec_pitch;

    StepCtx sc;
    sc.ncon = 0; sc.nefc = 0; sc.iters = 0;
    sc.t_prev = a.prof ? clock64() : 0;
    const long long t_item0 = clock64();
    const int cta_envs = min(nwarp, a.nenv - (int)blockIdx.x * nwarp);
    const int nsync_main = (nwarp > 1 && a.sync_stages && !a.sched) ? cta_envs * B2K_G : 0;
    sc.nsync = 0;

    // ---- resume a split step: bring the arena back from HBM ----
    if (a.mode == MODE_STEP_END) {
      for (int f = 0; f < B2MJ_NFIELD; f++) {
        const int os = m.off_s[f];
        if (os < 0) continue;
        if (m.fis_int[f]) { FORL(k, m.fsize[f]) si[os + k] = gi[m.off_g[f] + k]; }
        else { FORL(k, m.fsize[f]) sd[os + k] = gd[m.off_g[f] + k]; }
      }
      for (int f = 0; f < XF_COUNT; f++) {
        const int os = m.xoff_s[f];
        if (os < 0) continue;
        FORL(k, m.xsize[f]) sd[os + k] = gd[m.xoff_g[f] + k];
      }
      WSYNC();
    }

    triTableBuild(e);
    // ---- state record: HBM -> SMEM with one TMA bulk copy (segments A+B are contiguous) ----
    const unsigned load_bytes = (unsigned)(m.rec_C_begin - m.rec_A_begin) * 8u;
    if (!bar_ready) {
      if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      bar_ready = true;
      WSYNC();
    }
    if (lane == 0) {
      mbar_expect_tx(bar, load_bytes);
      bulk_g2s(smem_u32(sd + m.rec_A_begin), rec + m.rec_A_begin, load_bytes, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    if (m.nmocap && a.mocap) {
      const double* src = a.mocap + (size_t)env * 7 * m.nmocap;
      double* mp = e.D(B2MJ_F_MOCAP_POS);
      double* mq = e.D(B2MJ_F_MOCAP_QUAT);
      FORL(k, 3 * m.nmocap) mp[k] = src[k];
      FORL(k, 4 * m.nmocap) mq[k] = src[3 * m.nmocap + k];
    }
    WSYNC();

    PROF_MARK(PROF_LOAD)
    if (a.mode == MODE_STEP_END) {
      sc.ncon = e.I(B2MJ_F_NCON)[0];
      sc.nefc = e.I(B2MJ_F_NEFC)[0];
    }

    B2K_NOUNROLL for (int step = step0; step < step1; step++) {
      if (a.ctrl_seq) {  // fused rollout: this step's controls straight from the device-resident stream
        const double* src = a.ctrl_seq + ((size_t)step * a.nenv + env) * m.nu;
        double* ctrl = e.D(B2MJ_F_CTRL);
        FORL(i, m.nu) ctrl[i] = src[i];
        WSYNC();
      }
      if (a.mode != MODE_STEP_END && a.mode != MODE_FORWARD) {
        // mj_checkPos / mj_checkVel
        int bad = 0;
        { const double* q = e.D(B2MJ_F_QPOS); FORL(i, m.nq) bad |= isBad(q[i]); }
        if (__any_sync(e.mask, bad)) resetEnv(e, warning, B2MJ_WARN_BADQPOS);
        bad = 0;
        { const double* v = e.D(B2MJ_F_QVEL); FORL(i, m.nv) bad |= isBad(v[i]); }
        if (__any_sync(e.mask, bad)) resetEnv(e, warning, B2MJ_WARN_BADQVEL);
      }
      const bool first = a.mode != MODE_STEP_END, second = a.mode != MODE_STEP_BEGIN;
      const bool rk_split = a.mode == MODE_STEP_END && m.opt.integrator == B2MJ_INT_RK4;
      sc.nsync = nsync_main;
      // (split RK4: stages 1..3 are sensor-free forward passes, like the fused mj_RungeKutta)
      forwardPass(e, a, env, sc, rk_split && a.rk_stage > 0, first, second);
      sc.nsync = 0;
      if (a.mode == MODE_FORWARD || a.mode == MODE_STEP_BEGIN) break;
      if (rk_split && a.rk_stage > 0) {
        // second half of sub-step rk_stage done: record it, then either open the next sub-step (first half only: the
        // host's hooks run before its second half) or finish the step
        rk_record(e, a.rk_stage);
        if (a.rk_stage < 3) {
          rk_setup_stage(e, a.rk_stage + 1);
          forwardPass(e, a, env, sc, true, true, false);
        } else {
          rk_finish(e);
        }
        break;
      }
      // mj_checkAcc
      {
        int bad = 0;
        const double* qa = e.D(B2MJ_F_QACC);
        FORL(i, m.nv) bad |= isBad(qa[i]);
        if (__any_sync(e.mask, bad)) {
          resetEnv(e, warning, B2MJ_WARN_BADQACC);
          forwardPass(e, a, env, sc, false, true, true);
        }
      }
      if (rk_split) {  // sub-step 0 of a split RK4 step: open sub-step 1 and yield to the host
        rk_init(e);
        rk_setup_stage(e, 1);
        forwardPass(e, a, env, sc, true, true, false);
        break;
      }
      if (m.opt.integrator == B2MJ_INT_RK4 && a.mode == MODE_STEP) stage_rk4(e, a, env, sc);
      else if (m.opt.integrator == B2MJ_INT_IMPLICIT || m.opt.integrator == B2MJ_INT_IMPLICITFAST) stage_implicit(e);
      else stage_euler(e);
      PROF_MARK(PROF_INTEGRATE)
      if (a.traj_qpos) {
        double* dst = a.traj_qpos + ((size_t)step * a.nenv + env) * m.nq;
        const double* q = e.D(B2MJ_F_QPOS);
        FORL(i, m.nq) dst[i] = q[i];
      }
      if (a.traj_qvel) {
        double* dst = a.traj_qvel + ((size_t)step * a.nenv + env) * m.nv;
        const double* v = e.D(B2MJ_F_QVEL);
        FORL(i, m.nv) dst[i] = v[i];
      }
      if (a.traj_sensor && m.nsensordata) {
        double* dst = a.traj_sensor + ((size_t)step * a.nenv + env) * m.nsensordata;
        const double* sdat = e.D(B2MJ_F_SENSORDATA);
        FORL(i, m.nsensordata) dst[i] = sdat[i];
      }
    }

    // controls that came from a stream (fused rollout, zero-copy host exchange) are the env's ctrl from now on: the
    // bulk store below covers segments B + C only, so write segment A's ctrl back by hand
    if (a.ctrl_seq && m.nu) {
      const double* ctrl = e.D(B2MJ_F_CTRL);
      FORL(i, m.nu) rec[m.rec_ctrl + i] = ctrl[i];
    }
    // fused publish: this env's row into every rank's gathered slab, straight from the record image in shared memory
    // (peer stores over NVLink; they overlap the rest of the launch, whose length the slowest env sets)
    if (a.pub) {
      const PubArgs& p = *a.pub;
      B2K_NOUNROLL for (int r = 0; r < p.nranks; r++) {
        double* dst = p.slab[r] + ((size_t)p.rank * a.nenv + env) * p.count;
        B2K_NOUNROLL for (int f = 0; f < p.nfields; f++) {
          const double* src = sd + p.foff[f];
          FORL(i, p.fcnt[f]) dst[i] = src[i];
          dst += p.fcnt[f];
        }
      }
      // the env whose rows leave last raises this rank's sequence flag in every peer: the launch IS the collective,
      // no second kernel.  Every lane fences its own peer stores at system scope before the warp's arrival is counted.
      __threadfence_system();
      WSYNC();
      if (lane == 0) {
        const unsigned prev = atomicAdd(p.done, 1u);
        if (prev == (unsigned)a.nenv - 1u) {
          *p.done = 0;
          __threadfence_system();
          B2K_NOUNROLL for (int r = 0; r < p.nranks; r++) *reinterpret_cast<volatile int*>(p.flags[r] + p.rank) = a.pub_seq;
        }
      }
    }
    // ---- results: counters, state record SMEM -> HBM (segments B+C contiguous), optional arena dump ----
    if (lane == 0) {
      int* st = a.stats + (size_t)env * 4;
      st[0] = sc.ncon; st[1] = sc.nefc; st[2] = sc.iters;
      st[3] = (int)((clock64() - t_item0) >> 10);  // this work item's residency in 1024-cycle units
      if (a.cost) a.cost[env] = st[3];
      e.I(B2MJ_F_SOLVER_ITER)[0] = sc.iters;
      for (int k = 0; k < B2MJ_NWARNING; k++) e.I(B2MJ_F_WARNING)[k] = warning[k];
    }
    fence_async_smem();
    WSYNC();
    if (lane == 0) {
      // MODE_STEP_BEGIN keeps qpos (normalised quaternions) consistent too: store B+C in every mode
      bulk_s2g(rec + m.rec_B_begin, smem_u32(sd + m.rec_B_begin), (unsigned)(m.rec_end - m.rec_B_begin) * 8u);
      if (a.sched) {
        // the next chunk of this env may be picked up by any warp on any SM: publish after the writes land
        bulk_commit_wait_all();
        __threadfence();
        *((volatile int*)(a.sched + 1 + env)) = chunk_id + 1;
      } else {
        bulk_commit_wait();
      }
    }
    if (a.dump || a.mode == MODE_STEP_BEGIN ||
        (a.mode == MODE_STEP_END && m.opt.integrator == B2MJ_INT_RK4 && a.rk_stage < 3)) {
      for (int f = 0; f < B2MJ_NFIELD; f++) {
        const int os = m.off_s[f];
        if (os < 0) continue;
        if (m.fis_int[f]) { FORL(k, m.fsize[f]) gi[m.off_g[f] + k] = si[os + k]; }
        else { FORL(k, m.fsize[f]) gd[m.off_g[f] + k] = sd[os + k]; }
      }
      for (int f = 0; f < XF_COUNT; f++) {
        const int os = m.xoff_s[f];
        if (os < 0) continue;
        FORL(k, m.xsize[f]) gd[m.xoff_g[f] + k] = sd[os + k];
      }
    }
    WSYNC();
    PROF_MARK(PROF_STORE)
    if (tw > 1) team_release(e);  // retire the helpers of this env slot
    if (!a.sched) break;
  }
}

}  // namespace b2k

using namespace b2k;

#ifdef B2K_PER_ENV_MODEL
#define b2k_launch_step b2k_em_launch_step
#endif

// host shadow of what c_dm currently holds (per device), so the constant is re-uploaded only when a
// different handle / an edited model launches
static DevModel g_shadow[16];
static bool g_shadow_valid[16];
// one lock per device covers compare + upload + launch, so two host threads driving different handles on the same
// GPU cannot launch with each other's model (the constant is process-global per device)
static std::mutex g_launch_mutex[16];

extern "C" int b2k_launch_step(const DevModel* m, const LaunchArgs* a, int warps_per_cta, size_t smem_bytes,
                               cudaStream_t stream) {
  static size_t attr_bytes[16];
  int dev = 0;
  cudaGetDevice(&dev);
  const int dev_id = dev;
  dev &= 15;
  std::lock_guard<std::mutex> lock(g_launch_mutex[dev]);
  if (!attr_bytes[dev]) {  // opt in to the full 227 KB once per device (never lowered: occupancy queries rely on it)
    cudaError_t err = cudaFuncSetAttribute(b2k_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) return (int)err;
    attr_bytes[dev] = 227 * 1024;
  }
  if (const char* cv = getenv("B2MJ_CARVEOUT")) {  // experiment: shared-memory carveout in percent (more L1 for tables)
    static int applied = -1;
    if (applied != atoi(cv)) {
      cudaFuncSetAttribute(b2k_step_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv));
      applied = atoi(cv);
    }
  }
  if (!g_shadow_valid[dev] || memcmp(&g_shadow[dev], m, sizeof(DevModel)) != 0) {
    // kernels of another handle may still be reading the constant: drain the device first
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess) err = cudaMemcpyToSymbol(c_dm, m, sizeof(DevModel));
    if (err != cudaSuccess) return (int)err;
    memcpy(&g_shadow[dev], m, sizeof(DevModel));
    g_shadow_valid[dev] = true;
  }
  int ctas = (a->nenv + warps_per_cta - 1) / warps_per_cta;
  if (a->sched) {
    // persistent grid: exactly the CTAs that are co-resident (spinning on a ticket needs its producer running)
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, b2k_step_kernel, warps_per_cta * B2K_G * m->team_warps, smem_bytes);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    ctas = std::max(1, std::min(ctas, per_sm * sms));
  }
  // warps_per_cta counts env slots; in team mode every slot is team_warps warps wide
  b2k_step_kernel<<<ctas, warps_per_cta * B2K_G * m->team_warps, smem_bytes, stream>>>(*a);
  return (int)cudaGetLastError();
}

#ifndef B2K_PER_ENV_MODEL  /* launch order, attributes and occupancy are shared with the main build */
// Launch order for the next launch: envs sorted heaviest first by what their LAST step cost.  A launch ends when its
// slowest env does; an env with 21 rows at the 100-iteration PGS cap takes 4x the median step, and if it starts in the
// second wave its whole run is added to the launch.  Contact states persist from step to step, so last step's cost
// predicts this step's.  Weight = the env's measured residency of the last launch (stats[3], 1024-cycle units): it
// ranks PGS, CG and Newton envs, collision-heavy envs and RK4 sub-steps alike (round 2 used rows x iterations in four
// fixed classes tuned on the PGS config; every env of the Newton configs fell into the two lightest classes, so their
// launches ran 40-55 % over the balanced bound).  64 linear classes up to the launch's maximum; one CTA.
// Stable counting sort: within a class envs keep their index order.  The order never changes results
// (tests/test_gpu_paths.py::test_launch_order_does_not_change_results).
#define B2K_ORDER_CLASSES 64
__global__ void b2k_order_kernel(const int* __restrict__ stats, const int* __restrict__ wt, int wstride, int nenv,
                                 int* __restrict__ perm, int legacy) {
  __shared__ int count[B2K_ORDER_CLASSES], cursor[B2K_ORDER_CLASSES], wmax;
  __shared__ int wtot[32][B2K_ORDER_CLASSES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x < B2K_ORDER_CLASSES) count[threadIdx.x] = 0;
  if (threadIdx.x == 0) wmax = 0;
  __syncthreads();
  if (!legacy) {
    int mx = 0;
    for (int e = threadIdx.x; e < nenv; e += blockDim.x) mx = max(mx, wt[(size_t)wstride * e]);
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) atomicMax(&wmax, mx);
    __syncthreads();
  }
  const long long span = (long long)wmax + 1;
  auto cls = [&](int e) {
    if (legacy) {  // rows x iterations in the four round-2 classes (B2MJ_ORDER_LEGACY=1, kept for A/B runs)
      const int w = stats[4 * e + 1] * stats[4 * e + 2];
      return w >= 1200 ? 0 : w >= 400 ? 1 : w >= 100 ? 2 : 3;
    }
    const int k = (int)(((long long)max(wt[(size_t)wstride * e], 0) * B2K_ORDER_CLASSES) / span);
    return B2K_ORDER_CLASSES - 1 - min(k, B2K_ORDER_CLASSES - 1);
  };
  for (int e = threadIdx.x; e < nenv; e += blockDim.x) atomicAdd(&count[cls(e)], 1);  // integer totals: order-free
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < B2K_ORDER_CLASSES; c++) { cursor[c] = acc; acc += count[c]; }
  }
  __syncthreads();
  for (int base = 0; base < nenv; base += blockDim.x) {
    const int e = base + threadIdx.x;
    const int c = e < nenv ? cls(e) : -1;
    for (int k = lane; k < B2K_ORDER_CLASSES; k += 32) wtot[warp][k] = 0;
    __syncwarp();
    // lanes of the warp that share a class: rank inside the group in lane (= env index) order
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (c >= 0 && rank == 0) wtot[warp][c] = __popc(peers);
    __syncthreads();
    if (c >= 0) {
      int off = cursor[c];
      for (int w = 0; w < warp; w++) off += wtot[w][c];
      perm[off + rank] = e;
    }
    __syncthreads();
    if (threadIdx.x < B2K_ORDER_CLASSES) {
      int t = 0;
      for (int w = 0; w < nwarp; w++) t += wtot[w][threadIdx.x];
      cursor[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// cost == null: the weights are the residencies in stats[4 e + 3]; else a dense [nenv] array (asynchronous refresh)
extern "C" int b2k_launch_order(const int* stats, const int* cost, int nenv, int* perm, int legacy, cudaStream_t stream) {
  b2k_order_kernel<<<1, 1024, 0, stream>>>(stats, cost ? cost : stats + 3, cost ? 1 : 4, nenv, perm, legacy);
  return (int)cudaGetLastError();
}

#ifdef B2K_SOLVE_PROF
extern "C" int b2k_sprof_read(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  if (out16) cudaMemcpyFromSymbol(out16, b2k::g_sprof, 16 * sizeof(unsigned long long));
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(b2k::g_sprof, z, sizeof(z)); }
  return 0;
}
#endif

extern "C" int b2k_step_kernel_attrs(int* regs, int* static_smem, int* max_threads) {
  cudaFuncAttributes at;
  cudaError_t err = cudaFuncGetAttributes(&at, b2k_step_kernel);
  if (err != cudaSuccess) return (int)err;
  if (regs) *regs = at.numRegs;
  if (static_smem) *static_smem = (int)at.sharedSizeBytes;
  if (max_threads) *max_threads = at.maxThreadsPerBlock;
  return 0;
}

// CTAs of this shape the hardware really keeps resident per SM (registers, shared-memory allocation granularity and
// the per-CTA reserve included: a hand formula mis-sized the 7-warp rollout CTA by 96 bytes and silently halved its
// residency, -20 % on the fused rollout)
extern "C" int b2k_occupancy(int threads, size_t smem_bytes, int* ctas_per_sm) {
  cudaFuncSetAttribute(b2k_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, b2k_step_kernel, threads, smem_bytes);
  return (int)err;
}
#endif  // !B2K_PER_ENV_MODEL
