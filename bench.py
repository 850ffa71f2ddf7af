#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched physics step (BASELINE.json metric) on N B200s.

Workload (config C2 of BASELINE.json / SURVEY.md 8d): Panda-like 7+2-DoF arm, 4096 envs PER GPU, Euler, PGS, fresh
random ctrl ~ U(ctrlrange) for every env at every step.  A "step" is one mj_step of every env of the batch.

Every timed leg starts from the SAME contact-rich state: before anything is timed the batch is pre-rolled --preroll
(1000) steps with the workload's own random controls (untimed, one fused launch), and that state is snapshotted and
restored before each leg.  The first ~300 steps of the workload are contact free and ~30 % cheaper; without the
pre-roll a short run (--steps 20) would report that easy regime.  The reference arm pre-rolls the same way on the CPU.

  value        env-steps/s with inputs resident in HBM: one fused b2mj_rollout call advances all envs K steps reading
               the device-resident ctrl stream and writing the per-step trajectory (CUDA events on the stepping stream).
  per_step_launch  the same K steps as K closed-loop launches (b2mj_set_device(ctrl) + b2mj_step), L2 flushed between.
  e2e          the headline: same metric through the C-ABI with HOST buffers, one b2mj_step_host call per step
               (ctrl from pinned host memory -> step -> qpos / qvel / sensordata into pinned host memory -> one
               synchronisation).  With pinned buffers the step kernel itself moves those bytes over PCIe -- it reads
               an env's ctrl from host memory when it picks the env up and writes its outputs to host memory when
               the env's step is done -- so the transfers overlap the launch (B2MJ_NO_ZERO_COPY=1 = staged copies).
  roofline     algorithmic state bytes per env-step x env-steps per launch / kernel time against the
               measured HBM copy bandwidth (MEASURED_PEAKS.json); roofline.fp64 = FP64 instruction rate against the
               DFMA peak measured live on this GPU (b2mj_ubench_dfma) -- the roofline that can actually bind.
  parity       the gate reported with every number: 64-env subsample of the timed state, per-step
               state-injected comparison with the CPU oracle + free-running end state + contact-pair indices.
  publish      (N > 1) the one exchange step of the path, two ways: per-step launches WITH the NCCL all-gather of
               qpos | qvel | sensordata every step, against the same steps without it.
  configs      the other BASELINE configs at their per-GPU batch (C3 hand through the robot_hw path, C4 humanoid +
               sensor readout (+ publish), C5 bin) and, for N > 1, C2 strong-scaled (4096 / N envs per GPU), each with
               the CPU oracle timed beside it.
  cpu_baseline / --impl reference: the CPU oracle (restatement of the reference's mj_step loop; the reference itself
               cannot be built here: libmujoco + ROS are absent) on all host cores.

Launch: `python bench.py --gpus 1` or, for N>1,
`python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...`
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (batched mj_step)"
UNIT = "env-steps/s"
STATE = ("qpos", "qvel", "act", "qacc_warmstart", "time")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_profile_constants(model_name):
    """Per-env-step counters of the dominant kernel from the committed ncu capture of the fused rollout launch
    (profiles/r2c_traffic.json, else earlier rounds): DRAM bytes and FP64 thread-instructions.  PROFILE CONSTANTS, not live."""
    for name in ("r2c_traffic.json", "r2b_traffic.json", "r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            return {"dram": t.get("rollout_dram_bytes_per_env_step", {}).get(model_name),
                    "fp64_pipe_pct": t.get("rollout_fp64_pipe_active_pct_elapsed", {}).get(model_name),
                    "warp_inst": t.get("rollout_warp_inst_per_env_step", {}).get(model_name),
                    "source": f"profiles/{name}"}
        except Exception:
            continue
    return {"dram": None, "fp64_pipe_pct": None, "warp_inst": None, "source": None}


def state_bytes(model):
    """Algorithmic bytes per env-step (SURVEY 8d / DESIGN.md): state + inputs read, state + outputs written."""
    nq, nv, na, nu, ns = model.nq, model.nv, model.na, model.nu, model.nsensordata
    rd = nq + nv + na + nu + nv + nv + 1
    wr = nq + nv + na + nv + nv + ns + 1
    return 8 * (rd + wr)


def make_inputs(model, nenv, nsteps, seed, rank=0, amp=0.1):
    """qpos0 + U(-amp, amp) per scalar joint coordinate; ctrl ~ U(ctrlrange) per env per step (Philox counter RNG)."""
    rng = np.random.Generator(np.random.Philox(key=seed + 1000003 * rank))
    qpos = np.tile(model.qpos0, (nenv, 1))
    for j in range(model.njnt):
        t, qa = model.jnt_type[j], model.jnt_qposadr[j]
        if t in (2, 3):
            qpos[:, qa] += rng.uniform(-amp, amp, nenv)
        elif t == 0:  # free joint: jitter x, y, height; orientation stays the model's
            qpos[:, qa:qa + 2] += rng.uniform(-0.2 * amp, 0.2 * amp, (nenv, 2))
            qpos[:, qa + 2] += rng.uniform(0, 0.5 * amp, nenv)
    qvel = np.zeros((nenv, model.nv))
    if model.nu:
        lo, hi = model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1]
        lim = model.actuator_ctrllimited.astype(bool)
        lo, hi = np.where(lim, lo, -1.0), np.where(lim, hi, 1.0)
        ctrl = rng.uniform(lo, hi, (nsteps, nenv, model.nu))
    else:
        ctrl = np.zeros((nsteps, nenv, 0))
    return qpos, qvel, ctrl


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions through NVML (a 20-step timed region is 5 ms: an
    nvidia-smi poll at 100 ms would never land inside it).  Falls back to nvidia-smi -lms when pynvml is missing."""

    def __init__(self, torch_device_index):
        self.idx = torch_device_index
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        self._nv = None
        self._smi = None

    def _open(self):
        try:
            import pynvml as nv
            import torch

            nv.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
                h = nv.nvmlDeviceGetHandleByIndex(phys)
            self._nv, self._h = nv, h
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            return True
        except Exception:
            return False

    def _loop(self):
        nv, h = self._nv, self._h
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._open():
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
            return
        try:  # fallback: nvidia-smi polling
            fd, self._path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self._smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=open(self._path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self._smi = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.sm_max, reasons=sorted(self.reasons),
                           samples=len(self.samples), sm_min_mhz=float(min(self.samples)), source="nvml, 2 ms period")
            return out
        if self._smi is not None:
            self._smi.terminate()
            try:
                self._smi.wait(timeout=5)
            except Exception:
                self._smi.kill()
            sm, smax, reasons = [], [], set()
            try:
                for line in open(self._path):
                    p = [x.strip() for x in line.split(",")]
                    if len(p) < 6:
                        continue
                    try:
                        sm.append(float(p[0]))
                        smax.append(float(p[1]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
                os.unlink(self._path)
            except Exception:
                pass
            if sm:
                out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                           samples=len(sm), source="nvidia-smi -lms 20")
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_rollout_from(model, state, ctrl, nthreads, want_sensors=None, hw=None, hw_cmd=None):
    from oracle import binding as ob

    nsteps = ctrl.shape[0] if ctrl is not None else hw_cmd.shape[0]
    secs, st, _ = ob.rollout_ex(model, state, nsteps, ctrl=ctrl if model.nu else None, nthreads=nthreads,
                                want_sensors=(model.nsensordata > 0) if want_sensors is None else want_sensors,
                                hw=hw, hw_cmd=hw_cmd)
    return secs, st


def run_reference(args, model, workload):
    """--impl reference: the CPU implementation of the path on all host cores (oracle port; the reference's own mj_step
    lives in libmujoco 2.3.7 which is absent -> oracle/_ref unbuildable).  Same workload, same pre-roll, same steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nenv, K, W, P = args.nenv, args.steps, args.warmup, args.preroll
    qpos, qvel, ctrl = make_inputs(model, nenv, P + W + K, args.seed, 0)
    state = {"qpos": qpos, "qvel": qvel}
    if P:
        _, state = cpu_rollout_from(model, state, ctrl[:P] if model.nu else np.zeros((P, nenv, 0)), cores)
    _, state = cpu_rollout_from(model, state, ctrl[P:P + W] if model.nu else np.zeros((W, nenv, 0)), cores)
    secs, state = cpu_rollout_from(model, state, ctrl[P + W:] if model.nu else np.zeros((K, nenv, 0)), cores)
    value = nenv * K / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{nenv} envs x {K} steps after a {P}-step CPU pre-roll, oracle restatement of mj_step "
                                   f"(not libmujoco), {cores} std::threads over envs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "trajectory_finite": bool(np.all(np.isfinite(state["qpos"]))),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU helpers
class Ctx:
    pass


def snapshot(sim, model):
    return {k: sim.get(k) for k in STATE if model.field_size_by_name(k) > 0}


def restore(sim, snap):
    sim.reset()
    for k, v in snap.items():
        sim.set(k, v)


def parity_gate(model, snap, ctrl_seq, nsteps, nsub=64):
    """BASELINE.md "parity gate": 64-env subsample of the timed state.  Per step both sides start from the batch's own
    state (max rel error over qpos / qvel / qacc, per step), a second set of oracles runs free for the end-state
    divergence, and the contact-pair indices after the last step must be identical."""
    from mujoco_ros_pkgs_b200.batch import BatchSim
    from oracle import binding as ob

    nenv = snap["qpos"].shape[0]
    idx = np.linspace(0, nenv - 1, min(nsub, nenv)).astype(int)
    sub = {k: v[idx] for k, v in snap.items()}
    sim = BatchSim(model, len(idx))
    for k, v in sub.items():
        sim.set(k, v)
    inj = [ob.Oracle(model) for _ in idx]
    free = [ob.Oracle(model) for _ in idx]
    for e, o in enumerate(free):
        for k, v in sub.items():
            o.set(k, v[e])

    def rel(a, b):
        return float(np.max(np.abs(a - b) / (1.0 + np.abs(b)))) if a.size else 0.0

    worst_step, per_step = 0.0, []
    for s in range(nsteps):
        st = {k: sim.get(k) for k in sub}
        c = ctrl_seq[s][idx] if model.nu else None
        if model.nu:
            sim.set("ctrl", c)
        sim.step(1)
        out = {k: sim.get(k) for k in ("qpos", "qvel", "qacc")}
        w = 0.0
        for e in range(len(idx)):
            o = inj[e]
            for k, v in st.items():
                o.set(k, v[e])
            if model.nu:
                o.set("ctrl", c[e])
                free[e].set("ctrl", c[e])
            o.step(1)
            free[e].step(1)
            w = max(w, max(rel(out[k][e], o.get(k)) for k in out))
        per_step.append(w)
        worst_step = max(worst_step, w)
    gq = sim.get("qpos")
    end_free = max(rel(gq[e], free[e].get("qpos")) for e in range(len(idx)))
    # contact-pair indexing, bit-exact: forward pass on both sides from the batch's final state
    st = {k: sim.get(k) for k in sub}
    sim.keep_intermediates(True)
    sim.forward()
    ncon = sim.get("ncon")[:, 0]
    g1, g2 = sim.get("contact_geom1"), sim.get("contact_geom2")
    pairs = equal = 0
    for e in range(len(idx)):
        o = inj[e]
        for k, v in st.items():
            o.set(k, v[e])
        o.forward()
        n = int(o.get("ncon")[0])
        pairs += max(n, int(ncon[e]))
        if n == ncon[e]:
            equal += int(np.sum((g1[e][:n] == o.get("contact_geom1")[:n]) & (g2[e][:n] == o.get("contact_geom2")[:n])))
    sim.close()
    return {"envs": int(len(idx)), "steps": nsteps, "worst": worst_step, "per_step_worst_rel": worst_step,
            "end_state_free_running_rel": end_free, "contact_pairs": int(pairs), "contact_pairs_equal": int(equal),
            "tolerance": 1e-5, "pass": bool(worst_step < 1e-5 and pairs == equal),
            "how": "state-injected single steps vs the CPU oracle (oracle/), max over qpos/qvel/qacc of |a-b|/(1+|b|)"}


def time_fused(C, sim, snap, ctrl_dev, W, K, traj=True):
    """restore -> warm-up rollout (W steps) -> L2 flush -> ONE timed rollout launch of K steps.  Returns
    (ms, launches, trajectory-finite flag)."""
    torch, model, stream, nenv = C.torch, sim.model, C.stream, sim.nenv
    restore(sim, snap)
    nu, ns = model.nu, model.nsensordata
    n = max(K, W)
    tq = torch.empty(n, nenv, model.nq, dtype=torch.float64, device=C.dev) if traj else None
    tv = torch.empty(n, nenv, model.nv, dtype=torch.float64, device=C.dev) if traj else None
    ts = torch.empty(n, nenv, ns, dtype=torch.float64, device=C.dev) if (traj and ns) else None
    p = lambda t: t.data_ptr() if t is not None else 0  # noqa: E731
    with torch.cuda.stream(stream):
        sim.rollout(W, p(ctrl_dev[:W]) if nu else 0, p(tq), p(tv), p(ts))
        C.barrier()
        if not C.args.no_flush:
            C.flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = sim.launch_info()["launches"]
        e0.record(stream)
        sim.rollout(K, p(ctrl_dev[W:W + K]) if nu else 0, p(tq), p(tv), p(ts))
        e1.record(stream)
        launches = sim.launch_info()["launches"] - l0
        C.barrier()
    ok = bool(torch.isfinite(tq[K - 1]).all().item()) if traj else True
    return e0.elapsed_time(e1), launches, ok


def time_per_step(C, sim, snap, W, K, step_fn):
    """restore -> W closed-loop warm-up steps -> K timed steps, L2 flushed before each, CUDA events per step.
    step_fn(k) enqueues one step's launches on the stream.  Returns (total ms, per-step ms array, launches)."""
    torch, stream = C.torch, C.stream
    restore(sim, snap)
    with torch.cuda.stream(stream):
        for k in range(W):
            step_fn(k)
        C.barrier()
        l0 = sim.launch_info()["launches"]
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k in range(K):
            if not C.args.no_flush:
                C.flush_buf.fill_(k)  # evict state + model from L2 (buffer > 126 MB L2)
            evs[k][0].record(stream)
            step_fn(W + k)
            evs[k][1].record(stream)
        C.barrier()
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    return float(ms.sum()), ms, sim.launch_info()["launches"] - l0


def time_e2e(C, sim, snap, W, K, host_step):
    """restore -> W warm-up -> K timed end-to-end steps by wall clock with a device sync on both sides."""
    restore(sim, snap)
    for k in range(W):
        host_step(k)
    C.barrier()
    t0 = time.perf_counter()
    for k in range(K):
        host_step(W + k)
    C.barrier()
    return time.perf_counter() - t0


def reduce_max(C, *vals):
    t = C.torch.tensor(list(vals), dtype=C.torch.float64, device=C.dev)
    if C.world > 1:
        C.dist.all_reduce(t, op=C.dist.ReduceOp.MAX)
    return [float(x) for x in t]


def per_rank_stats(C, ms):
    """{min, median, max} of every rank's per-launch kernel times, so a scaling loss can be attributed."""
    t = C.torch.tensor([float(np.min(ms)), float(np.median(ms)), float(np.max(ms)), float(np.sum(ms))],
                       dtype=C.torch.float64, device=C.dev)
    if C.world > 1:
        out = C.torch.empty(C.world, 4, dtype=C.torch.float64, device=C.dev)
        C.dist.all_gather_into_tensor(out, t)
        out = out.cpu().numpy()
    else:
        out = t.cpu().numpy()[None]
    return [{"rank": r, "min_ms": float(x[0]), "median_ms": float(x[1]), "max_ms": float(x[2]), "sum_ms": float(x[3])}
            for r, x in enumerate(out)]


# ------------------------------------------------------------------------------------------------ other configs
HAND_JOINTS = ["WRJ1", "WRJ0", "FFJ3", "FFJ2", "FFJ1", "MFJ3", "MFJ2", "MFJ1", "RFJ3", "RFJ2", "RFJ1", "LFJ4", "LFJ3",
               "LFJ2", "LFJ1", "THJ4", "THJ3", "THJ2", "THJ1", "THJ0"]


def hand_hw_setup(model, capi):
    """C3: 20 actuated joints through the robot_hw path, EFFORT and POSITION_PID alternating."""
    jids = [model.name2id(capi.OBJ_JOINT, n) for n in HAND_JOINTS]
    jids = [j for j in jids if j >= 0]
    nj = len(jids)
    modes = [0 if k % 2 == 0 else 2 for k in range(nj)]
    lower = [float(model.jnt_range[j, 0]) for j in jids]
    upper = [float(model.jnt_range[j, 1]) for j in jids]
    effort = [1.0] * nj
    pid5 = np.tile([3.0, 0.5, 0.05, 0.2, -0.2], (nj, 1))
    return dict(joint_ids=jids, modes=modes, kinds=[0] * nj, lower=lower, upper=upper, effort=effort, pid5=pid5,
                pid6=np.concatenate([pid5, np.zeros((nj, 1))], axis=1))


def hand_commands(hw, nsteps, nenv, seed):
    rng = np.random.Generator(np.random.Philox(key=seed + 77))
    lo, hi, eff = np.array(hw["lower"]), np.array(hw["upper"]), np.array(hw["effort"])
    is_pos = np.array(hw["modes"]) == 2
    return np.where(is_pos, rng.uniform(lo, hi, (nsteps, nenv, len(lo))), rng.uniform(-1, 1, (nsteps, nenv, len(lo))) * eff)


def measure_config(C, tag, model_file, nenv, pre, W, K, with_cpu, comm=None, describe=""):
    """One BASELINE config at its per-GPU batch: closed-loop per-step launches incl. the plugin kernels of its path,
    the fused rollout where the path allows one, end to end with host buffers, and the CPU oracle beside it."""
    from mujoco_ros_pkgs_b200 import _capi
    from mujoco_ros_pkgs_b200.batch import BatchSim

    torch = C.torch
    model = _capi.Model.from_xml_file(os.path.join(ROOT, "mujoco_ros_pkgs_b200", "models", model_file))
    nu, ns = model.nu, model.nsensordata
    total = pre + W + K
    qpos, qvel, ctrl = make_inputs(model, nenv, total, C.args.seed + 31, C.rank, amp=0.05)
    sim = BatchSim(model, nenv, device=C.local_rank)
    sim.set_stream(C.stream.cuda_stream)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    ctrl_dev = torch.from_numpy(ctrl).to(C.dev) if nu else None
    hw = cmd = cmd_dev = None
    if tag == "C3":
        hw = hand_hw_setup(model, _capi)
        sim.robot_hw_configure(hw["joint_ids"], hw["modes"], effort_limit=hw["effort"], pid=hw["pid5"],
                               lower=hw["lower"], upper=hw["upper"], kind=hw["kinds"])
        cmd = hand_commands(hw, total, nenv, C.args.seed)
        cmd_dev = torch.from_numpy(cmd).to(C.dev)
    dt = model.opt.timestep
    pub_dst = None
    pub_fields = ["qpos", "qvel", "sensordata"]
    if comm is not None and tag == "C4":
        pub_dst = torch.empty(C.world, nenv, model.nq + model.nv + ns, dtype=torch.float64, device=C.dev)

    def dev_step(k, off=pre):
        if hw is not None:
            sim.robot_hw_refresh()
            sim.robot_hw_write_device(cmd_dev[off + k].data_ptr(), period=dt)
        elif nu:
            sim.set_device("ctrl", ctrl_dev[off + k].data_ptr(), nu)
        sim.step(1)
        if tag == "C4":
            sim.sensor_readout_device(False)          # mujoco_ros_sensors lastStage readout, device resident
            if pub_dst is not None:
                sim.allgather_publish_multi(pub_fields, comm.ptr, pub_dst.data_ptr())

    # pre-roll through the config's own path (untimed)
    with torch.cuda.stream(C.stream):
        if hw is None:
            sim.rollout(pre, ctrl_dev[:pre].data_ptr() if nu else 0)
        else:
            for k in range(pre):
                dev_step(k, 0)
        C.barrier()
    snap = snapshot(sim, model)
    res = {"model": model_file, "nenv_per_gpu": nenv, "preroll_steps": pre, "steps": K, "warmup": W, "describe": describe,
           "nq": model.nq, "nv": model.nv, "integrator": {0: "Euler", 1: "RK4"}.get(model.opt.integrator),
           "solver": {0: "PGS", 1: "CG", 2: "Newton"}.get(model.opt.solver)}
    ps_ms, ps_arr, ps_l = time_per_step(C, sim, snap, W, K, dev_step)
    (ps_ms,) = reduce_max(C, ps_ms)
    res["per_step_launch"] = {"value": nenv * C.world * K / (ps_ms * 1e-3), "unit": UNIT, "ms_per_step": ps_ms / K,
                              "gpu_launches": ps_l}
    if hw is None:
        f_ms, f_l, ok = time_fused(C, sim, snap, ctrl_dev[pre:] if nu else None, W, K, traj=True)
        (f_ms,) = reduce_max(C, f_ms)
        res["rollout"] = {"value": nenv * C.world * K / (f_ms * 1e-3), "unit": UNIT, "ms_per_step": f_ms / K,
                          "gpu_launches": f_l, "trajectory_finite": ok}
    # end to end with host buffers
    KE = min(K, C.args.e2e_steps or K)
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()  # noqa: E731
    h_in = pin(nenv, len(hw["joint_ids"])) if hw is not None else (pin(nenv, nu) if nu else None)
    h_q, h_v, h_s = pin(nenv, model.nq), pin(nenv, model.nv), pin(nenv, max(ns, 1))
    stats = {k: sim.get(k)[:, 0] for k in ("ncon", "nefc", "solver_iter")}

    def host_step(k):
        if hw is not None:
            np.copyto(h_in, cmd[pre + k])
            sim.robot_hw_refresh()
            sim.robot_hw_write(h_in, period=dt)
            sim.step_host(1, None, h_q, h_v, h_s if ns else None)
        else:
            if nu:
                np.copyto(h_in, ctrl[pre + k])
            sim.step_host(1, h_in if nu else None, h_q, h_v, h_s if ns else None)

    e_s = time_e2e(C, sim, snap, min(W, 5), KE, host_step)
    (e_s,) = reduce_max(C, e_s)
    res["e2e"] = {"value": nenv * C.world * KE / e_s, "unit": UNIT, "steps": KE,
                  "h2d_bytes_per_step": int(h_in.nbytes) if h_in is not None else 0,
                  "d2h_bytes_per_step": int(nenv * (model.nq + model.nv + ns) * 8)}
    res["workload_stats"] = {"ncon_mean": float(stats["ncon"].mean()), "nefc_mean": float(stats["nefc"].mean()),
                             "nefc_max": int(stats["nefc"].max()), "solver_iter_mean": float(stats["solver_iter"].mean())}
    if with_cpu and C.rank == 0:
        cores = os.cpu_count() or 1
        cs = max(4, min(K, 20))
        kw = dict(hw={k: hw[k] for k in ("joint_ids", "modes", "kinds", "lower", "upper", "effort", "pid6")},
                  hw_cmd=cmd[pre + W:pre + W + cs]) if hw is not None else {}
        secs, _ = cpu_rollout_from(model, snap, ctrl[pre + W:pre + W + cs] if nu else np.zeros((cs, nenv, 0)), cores, **kw)
        # grow the sample towards ~3 s of CPU work, bounded by the timed step count
        cs2 = int(min(max(cs, 3.0 / max(secs / cs, 1e-9)), K))
        if cs2 > cs:
            if hw is not None:
                kw["hw_cmd"] = cmd[pre + W:pre + W + cs2]
            secs, _ = cpu_rollout_from(model, snap, ctrl[pre + W:pre + W + cs2] if nu else np.zeros((cs2, nenv, 0)), cores, **kw)
            cs = cs2
        res["cpu_baseline"] = {"value": nenv * cs / secs, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{nenv} envs x {cs} steps from the same pre-rolled state, same path, {cores} threads"}
        res["e2e_vs_cpu_one_gpu_share"] = res["e2e"]["value"] / C.world / res["cpu_baseline"]["value"]
    sim.close()
    return res


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--preroll", type=int, default=1000,
                    help="untimed steps before every timed leg (independent of --warmup): reach the contact-rich regime")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nenv", type=int, default=4096, help="envs per GPU (weak scaling)")
    ap.add_argument("--model", default="panda_like.xml")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps (diagnostic only)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5 (and strong-scaled C2) block")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate (diagnostic sweeps only)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default: --steps)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    from mujoco_ros_pkgs_b200 import _capi

    model = _capi.Model.from_xml_file(os.path.join(ROOT, "mujoco_ros_pkgs_b200", "models", args.model))
    integ = {0: "Euler", 1: "RK4"}.get(model.opt.integrator, str(model.opt.integrator))
    solver = {0: "PGS", 1: "CG", 2: "Newton"}.get(model.opt.solver, str(model.opt.solver))
    workload = {
        "workload": f"C2: {args.model} (nq={model.nq} nv={model.nv} nu={model.nu} nbody={model.nbody}), "
                    f"{args.nenv} envs per GPU, {integ}, {solver}, dt={model.opt.timestep}, random ctrl every step, "
                    f"timed from the contact-rich state after a {args.preroll}-step pre-roll",
        "nenv_per_gpu": args.nenv, "global_envs": args.nenv * args.gpus, "parallelism": f"env-shard x{args.gpus}",
        "preroll_steps": args.preroll,
        "launch_mode": "fused open-loop rollout: ONE b2mj_rollout call = one b2k_step_kernel launch that advances every env "
                       "--steps steps (+ the one-CTA launch-order kernel); ctrl stream [steps][nenv][nu] resident in HBM, "
                       "per-step qpos/qvel/sensordata trajectory written to HBM",
        "l2": "L2 flushed (256 MB fill) before the timed rollout launch and before every per-step launch"
        if not args.no_flush else "NOT flushed (diagnostic)",
    }
    if args.impl == "reference":
        run_reference(args, model, workload)
        return

    import torch
    import torch.distributed as dist

    from mujoco_ros_pkgs_b200.batch import BatchSim, lib as b2lib

    C = Ctx()
    C.args, C.torch, C.dist, C.model = args, torch, dist, model
    C.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    C.rank = rank = int(os.environ.get("RANK", "0"))
    C.local_rank = local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the batched step has no CPU fallback")
    torch.cuda.set_device(local_rank)
    C.dev = dev = torch.device("cuda", local_rank)
    nenv, K, W, P = args.nenv, args.steps, args.warmup, args.preroll

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    C.barrier = barrier
    qpos, qvel, ctrl = make_inputs(model, nenv, P + W + K, args.seed, rank)
    sim = BatchSim(model, nenv, device=local_rank)
    C.stream = stream = torch.cuda.Stream(device=dev)
    sim.set_stream(stream.cuda_stream)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    nu, ns = model.nu, model.nsensordata
    ctrl_dev = torch.from_numpy(ctrl).to(dev) if nu else None   # [P+W+K][nenv][nu] resident in HBM
    C.flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    wall0 = time.perf_counter()

    # ---------------- pre-roll to the contact-rich regime (untimed), snapshot ----------------
    with torch.cuda.stream(stream):
        if P:
            sim.rollout(P, ctrl_dev[:P].data_ptr() if nu else 0)
        barrier()
    snap = snapshot(sim, model)
    tctrl = ctrl_dev[P:] if nu else None   # [W+K]: warm-up then timed controls, the same for every leg

    # ---------------- leg A: closed-loop per-step launches ----------------
    def dev_step(k):
        if nu:
            sim.set_device("ctrl", tctrl[k].data_ptr(), nu)
        sim.step(1)

    ps_ms, ps_arr, ps_launches = time_per_step(C, sim, snap, W, K, dev_step)
    stats = {k: sim.get(k)[:, 0] for k in ("ncon", "nefc", "solver_iter")}
    rank_stats = per_rank_stats(C, ps_arr)

    # ---------------- leg B (value): fused open-loop rollout ----------------
    step_ms, rollout_launches, traj_ok = time_fused(C, sim, snap, tctrl, W, K)
    warn = sim.get("warning").sum(0).tolist()

    # ---------------- leg C (e2e, the headline): host buffers through the C-ABI ----------------
    KE = args.e2e_steps or K
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()  # noqa: E731
    h_ctrl = pin(nenv, nu) if nu else None
    h_qpos, h_qvel, h_sens = pin(nenv, model.nq), pin(nenv, model.nv), pin(nenv, max(ns, 1))
    hctrl = ctrl[P:]

    def host_step(k):
        if nu:
            np.copyto(h_ctrl, hctrl[k % (W + K)])  # the producer's write into the pinned staging buffer
        sim.step_host(1, h_ctrl if nu else None, h_qpos, h_qvel, h_sens if ns else None)

    e2e_s = time_e2e(C, sim, snap, W, KE, host_step)

    # ---------------- leg D (N > 1): per-step launches + the publish all-gather ----------------
    publish = None
    comm = None
    if world > 1:
        from mujoco_ros_pkgs_b200.nccl_comm import NcclComm

        comm = NcclComm(rank, world)
        row = model.nq + model.nv + ns
        pub_dst = torch.empty(world, nenv, row, dtype=torch.float64, device=dev)
        fields = ["qpos", "qvel", "sensordata"]

        def pub_step(k):
            dev_step(k)
            sim.allgather_publish_multi(fields, comm.ptr, pub_dst.data_ptr())

        pb_ms, pb_arr, pb_launches = time_per_step(C, sim, snap, W, K, pub_step)
        # correctness of the exchange: every rank's slot holds that rank's state (checked against a plain gather)
        mine = torch.from_numpy(np.concatenate([sim.get(f) for f in fields], axis=1)).to(dev)
        allv = torch.empty(world, nenv, row, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allv, mine)
        pub_ok = bool(torch.equal(allv, pub_dst))
        pb_ms_max, ps_ms_max = reduce_max(C, pb_ms, ps_ms)
        publish = {"us_per_step": 1e3 * (pb_ms_max - ps_ms_max) / K, "bytes_per_rank_per_step": nenv * row * 8,
                   "bytes_gathered_per_step": world * nenv * row * 8, "fields": fields,
                   "overhead_pct": 100.0 * (pb_ms_max - ps_ms_max) / ps_ms_max,
                   "with_publish_value": nenv * world * K / (pb_ms_max * 1e-3), "unit": UNIT,
                   "gathered_equals_plain_all_gather": pub_ok, "gpu_launches": pb_launches,
                   "how": "leg A's K closed-loop steps, each followed by b2mj_allgather_publish_multi (one pack kernel + one "
                          "ncclAllGather over NVLink); CUDA events per step, max over ranks", "per_rank": per_rank_stats(C, pb_arr)}

        # ---- the same exchange FUSED into the step kernel: every finished env's row is stored straight into every
        #      rank's gathered slab over NVLink peer memory (b2mj_step_publish), a stream-ordered flag wait follows
        try:
            handles = [None] * world
            dist.all_gather_object(handles, sim.publish_fused_create(world, rank, fields))
            sim.publish_fused_connect(handles)
            dist.barrier()
            fused_ptr = [0]

            def fused_step(k):
                if nu:
                    sim.set_device("ctrl", tctrl[k].data_ptr(), nu)
                sim.step_publish()
                fused_ptr[0], _ = sim.publish_fused_wait()

            pf_ms, pf_arr, pf_launches = time_per_step(C, sim, snap, W, K, fused_step)

            class _Slab:
                __cuda_array_interface__ = {"shape": (world, nenv, row), "typestr": "<f8", "data": (fused_ptr[0], False), "version": 3}

            fused_slab = torch.as_tensor(_Slab(), device=dev).clone()
            mine = torch.from_numpy(np.concatenate([sim.get(f) for f in fields], axis=1)).to(dev)
            dist.all_gather_into_tensor(allv, mine)
            fused_ok = bool(torch.equal(allv, fused_slab))
            (pf_ms_max,) = reduce_max(C, pf_ms)
            publish["fused"] = {"us_per_step": 1e3 * (pf_ms_max - ps_ms_max) / K,
                                "overhead_pct": 100.0 * (pf_ms_max - ps_ms_max) / ps_ms_max,
                                "with_publish_value": nenv * world * K / (pf_ms_max * 1e-3), "unit": UNIT,
                                "gathered_equals_plain_all_gather": fused_ok, "gpu_launches": pf_launches,
                                "how": "leg A's K closed-loop steps as b2mj_step_publish (the step kernel stores each finished "
                                       "env's row into every rank's slab through CUDA-IPC peer pointers and its last env raises "
                                       "the rank's flag in every peer) + b2mj_publish_fused_wait (stream-ordered spin on the "
                                       "local flags); no NCCL call; "
                                       "CUDA events per step, max over ranks", "per_rank": per_rank_stats(C, pf_arr)}
        except Exception as ex:  # peer access unavailable on this box: report, keep the NCCL number
            publish["fused"] = {"unavailable": str(ex)[:200]}

    clocks = sampler.stop()
    wall = time.perf_counter() - wall0
    rollout_rank_stats = per_rank_stats(C, [step_ms])  # every rank's own rollout launch: attributes a weak-scaling loss
    step_ms, ps_ms, e2e_s = reduce_max(C, step_ms, ps_ms, e2e_s)
    value = nenv * world * K / (step_ms * 1e-3)
    ps_value = nenv * world * K / (ps_ms * 1e-3)
    e2e_value = nenv * world * KE / e2e_s

    # ---------------- parity gate (rank 0, outside every timed region) ----------------
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_gate(model, snap, ctrl[P + W:], min(K, 25))

    # ---------------- FP64 peak measured live ----------------
    import ctypes as Cc
    tf, per_clk = Cc.c_double(), Cc.c_double()
    b2lib.b2mj_ubench_dfma(local_rank, Cc.byref(tf), Cc.byref(per_clk))

    # ---------------- other configs ----------------
    configs = None
    if not args.no_configs and args.model == "panda_like.xml":
        configs = {}
        kc, wc = min(K, 100), min(W, 10)
        with_cpu = not args.no_cpu
        if world > 1:
            n_strong = max(1, 4096 // world)
            configs["C2_strong"] = measure_config(C, "C2s", "panda_like.xml", n_strong, P, wc, min(K, 200), with_cpu,
                                                  describe=f"4096 Panda envs over {world} GPUs ({n_strong} per GPU), north-star target config")
        configs["C3"] = measure_config(C, "C3", "hand_like.xml", 1024, 300, wc, kc, with_cpu,
                                       describe="Shadow-Hand-like, RK4, Newton, elliptic; every step: robot_hw read + write (EFFORT / POSITION_PID) + mj_step")
        configs["C4"] = measure_config(C, "C4", "humanoid_like.xml", 2048, 300, wc, kc, with_cpu, comm=comm,
                                       describe="humanoid + floor contacts; every step: ctrl + mj_step + sensor readout kernel"
                                                + (" + publish all-gather" if comm else ""))
        configs["C5"] = measure_config(C, "C5", "bin.xml", 512, 300, wc, min(kc, 50), with_cpu,
                                       describe="20 free boxes in a bin, Newton, elliptic (collision heavy)")

    if rank == 0:
        peak, peak_src = load_peaks()
        bstate = state_bytes(model)
        kernel_s = step_ms * 1e-3          # one launch = nenv * K env-steps
        achieved = bstate * nenv * K / kernel_s / 1e9
        info = sim.launch_info()
        prof = load_profile_constants(args.model)
        fp64 = {"peak_measured_tflops": tf.value, "dfma_per_clk_per_sm": per_clk.value,
                "peak_source": "b2mj_ubench_dfma, measured live on this GPU"}
        if prof["fp64_pipe_pct"]:
            fp64.update(frac=prof["fp64_pipe_pct"] / 100.0, achieved_tflops=tf.value * prof["fp64_pipe_pct"] / 100.0,
                        warp_inst_per_env_step=prof["warp_inst"],
                        frac_source=f"PROFILE CONSTANT ({prof['source']}: ncu sm__pipe_fp64_cycles_active, % of elapsed cycles, of "
                                    "the rollout launch) -- the share of the FP64 pipe's issue capacity the kernel uses; "
                                    "achieved_tflops = that share x the peak measured live")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nenv * nu * 8,
                    "d2h_bytes_per_step": nenv * (model.nq + model.nv + ns) * 8, "steps": KE,
                    "timing": "wall clock, sync both sides, pinned host buffers, one b2mj_step_host call per step; the step "
                              "kernel reads ctrl from / writes the outputs to the pinned host buffers itself (zero-copy over "
                              "PCIe inside the timed region)" if not os.environ.get("B2MJ_NO_ZERO_COPY") else
                              "wall clock, sync both sides, pinned host buffers, one b2mj_step_host call per step (staged copies)"},
            "gpu_launches": rollout_launches,
            "per_rank_rollout_ms": [{"rank": r["rank"], "ms": r["median_ms"]} for r in rollout_rank_stats],
            "per_step_launch": {"value": ps_value, "unit": UNIT, "ms_per_step": ps_ms / K, "gpu_launches": ps_launches,
                                "per_rank_kernel_ms": rank_stats,
                                "note": "K x (b2mj_set_device(ctrl) + b2mj_step(1) + launch-order refresh), CUDA events per step, "
                                        "L2 flushed between steps"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (prof["dram"] * nenv * K) if prof["dram"] else None,
                         "traffic_source": f"PROFILE CONSTANT, not measured in this run ({prof['source']}: dram__bytes of the ncu "
                                           "--set full capture of the rollout launch, per env-step x this launch's env-steps)"
                         if prof["dram"] else None,
                         "peak_source": peak_src, "kernel": "b2k_step_kernel",
                         "algorithmic_bytes_per_env_step": bstate, "env_steps_per_launch": nenv * K,
                         "kernel_ms_per_launch": step_ms, "fp64": fp64,
                         "note": "latency / FP64-issue-bound path, not HBM-bound: see DESIGN.md 'Roofline'"},
            "kernel": {k: info[k] for k in ("warps_per_cta", "ctas", "smem_bytes_per_cta", "regs_per_thread",
                                            "arena_in_smem", "state_record_bytes")},
            "trajectory_finite": traj_ok,
            "workload_stats": {"warnings": warn, "ncon_mean": float(stats["ncon"].mean()),
                               "nefc_mean": float(stats["nefc"].mean()), "nefc_max": int(stats["nefc"].max()),
                               "solver_iter_mean": float(stats["solver_iter"].mean()),
                               "solver_iter_max": int(stats["solver_iter"].max())},
            "wall_s_timed_legs": wall,
        }
        if parity is not None:
            line["parity"] = parity
        if publish is not None:
            line["publish"] = publish
        if configs is not None:
            line["configs"] = configs
        if not args.no_cpu and world == 1:  # CPU baseline: rank 0 at N=1 only (bounded sample of the same timed steps)
            cores = os.cpu_count() or 1
            cs = min(K, 20)
            tc = ctrl[P + W:]
            secs, _ = cpu_rollout_from(model, snap, tc[:cs], cores)
            cs2 = int(min(max(cs, 10.0 / max(secs / cs, 1e-9)), K))   # ~10 s of CPU work, bounded by the GPU arm's steps
            if cs2 > cs:
                secs, _ = cpu_rollout_from(model, snap, tc[:cs2], cores)
                cs = cs2
            n1 = max(64, nenv // 16)
            secs1, _ = cpu_rollout_from(model, {k: v[:n1] for k, v in snap.items()}, tc[:cs, :n1], 1)
            line["cpu_baseline"] = {
                "value": nenv * cs / secs, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{nenv} envs x {cs} steps of the same timed workload (from the pre-rolled state), oracle "
                          f"restatement of mj_step (not libmujoco), {cores} threads",
                "single_thread_value": n1 * cs / secs1,
            }
        print(json.dumps(line), flush=True)
    if comm is not None:
        barrier()
        comm.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
