#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched physics step (BASELINE.json metric) on N B200s.

Workload (config C2 of BASELINE.json / SURVEY.md 8d): Panda-like 7+2-DoF arm, 4096 envs PER GPU,
Euler integrator, PGS solver, fresh random ctrl ~ U(ctrlrange) for every env at every step.
A "step" is one mj_step of every env of the batch with fresh controls.

  value      env-steps/s with inputs resident in HBM: one fused b2mj_rollout call advances all envs K steps
             reading the device-resident ctrl stream and writing the per-step trajectory (CUDA events on the
             stepping stream).  per_step_launch reports the same K steps as K separate launches
             (b2mj_set_device + b2mj_step, L2 flushed between steps) -- the closed-loop usage.
  e2e        same metric through the C-ABI with HOST buffers: per step b2mj_set(ctrl) from pinned host
             memory -> b2mj_step -> b2mj_get(qpos, qvel, sensordata) into pinned host memory (b2mj_step_host:
             the same four transfers and the launch queued in one call, one synchronisation).
  roofline   algorithmic state bytes per env-step x envs / kernel time, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the CPU oracle (restatement of the reference's mj_step loop; the
             reference itself cannot be built here: libmujoco + ROS are absent) on the host cores.

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
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (batched mj_step)"
UNIT = "env-steps/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(model_name, env_steps_per_launch):
    """DRAM bytes of the dominant kernel per launch, from the committed ncu --set full capture of the same launch
    shape (profiles/r1_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the fused rollout launch,
    recorded per env-step so it can be scaled to this run's launch size).  None if no capture exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)
        per = t["rollout_dram_bytes_per_env_step"].get(model_name)
        return None if per is None else float(per) * env_steps_per_launch
    except Exception:
        return None


def state_bytes(model):
    """Algorithmic bytes per env-step (SURVEY 8d / DESIGN.md): state + inputs read, state + outputs written."""
    nq, nv, na, nu, ns = model.nq, model.nv, model.na, model.nu, model.nsensordata
    rd = nq + nv + na + nu + nv + nv + 1
    wr = nq + nv + na + nv + nv + ns + 1
    return 8 * (rd + wr)


def make_inputs(model, nenv, nsteps, seed, rank=0):
    """qpos0 + U(-0.1,0.1) per joint; ctrl ~ U(ctrlrange) per env per step (Philox counter RNG)."""
    rng = np.random.Generator(np.random.Philox(key=seed + 1000003 * rank))
    qpos = np.tile(model.qpos0, (nenv, 1)) + rng.uniform(-0.1, 0.1, (nenv, model.nq))
    qvel = np.zeros((nenv, model.nv))
    lo, hi = model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1]
    lim = model.actuator_ctrllimited.astype(bool)
    lo = np.where(lim, lo, -1.0)
    hi = np.where(lim, hi, 1.0)
    ctrl = rng.uniform(lo, hi, (nsteps, nenv, model.nu)) if model.nu else np.zeros((nsteps, nenv, 0))
    return qpos, qvel, ctrl


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_rollout(model, nenv, nsteps, seed, nthreads):
    from oracle import binding as ob

    qpos, qvel, ctrl = make_inputs(model, nenv, nsteps, seed)
    secs, _, _, _ = ob.rollout(model, qpos, qvel, nsteps, ctrl=ctrl if model.nu else None, nthreads=nthreads,
                               want_sensors=model.nsensordata > 0)
    return secs


def run_reference(args, model, workload):
    """--impl reference: the CPU implementation of the path on all host cores (oracle port; the
    reference's own mj_step lives in libmujoco 2.3.7 which is absent -> oracle/_ref unbuildable)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nenv = args.nenv
    # one step = one pass of the oracle over the nenv-env batch (threads over envs)
    cpu_rollout(model, nenv, max(1, args.warmup), args.seed, cores)
    secs = cpu_rollout(model, nenv, args.steps, args.seed, cores)
    value = nenv * args.steps / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{nenv} envs x {args.steps} steps, oracle restatement of mj_step (not libmujoco), "
                                   f"{cores} std::threads over envs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nenv", type=int, default=4096, help="envs per GPU (weak scaling)")
    ap.add_argument("--model", default="panda_like.xml")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps (diagnostic only)")
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
                    f"{args.nenv} envs per GPU, {integ}, {solver}, dt={model.opt.timestep}, random ctrl every step",
        "nenv_per_gpu": args.nenv, "global_envs": args.nenv * args.gpus, "parallelism": f"env-shard x{args.gpus}",
        "launch_mode": "fused open-loop rollout: ONE b2mj_rollout call = one b2k_step_kernel launch that advances every env "
                       "--steps steps (+ the one-CTA launch-order kernel); ctrl stream [steps][nenv][nu] resident in HBM, "
                       "per-step qpos/qvel/sensordata trajectory written to HBM",
        "l2": ("inputs larger than L2 (ctrl stream + trajectory, %.0f MB); L2 flushed before the timed launch"
               % ((args.steps * args.nenv * (model.nu + model.nq + model.nv + model.nsensordata) * 8) / 1e6))
        if not args.no_flush else "NOT flushed (diagnostic)",
    }
    if args.impl == "reference":
        run_reference(args, model, workload)
        return

    import torch
    import torch.distributed as dist

    from mujoco_ros_pkgs_b200.batch import BatchSim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the batched step has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nenv, K, W = args.nenv, args.steps, args.warmup

    qpos, qvel, ctrl = make_inputs(model, nenv, K + W, args.seed, rank)
    sim = BatchSim(model, nenv, device=local_rank)
    stream = torch.cuda.Stream(device=dev)
    sim.set_stream(stream.cuda_stream)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    nu = model.nu
    ctrl_dev = torch.from_numpy(ctrl).to(dev) if nu else None  # [K+W][nenv][nu] resident in HBM
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(k):
        if nu:
            sim.set_device("ctrl", ctrl_dev[k].data_ptr(), nu)
        sim.step(1)

    ns = model.nsensordata
    # ---------------- device-resident leg A: per-step launches (what a closed-loop user does) ----------------
    with torch.cuda.stream(stream):
        for k in range(W):
            one_step(k)
        barrier()
        launches0 = sim.launch_info()["launches"]
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k in range(K):
            if not args.no_flush:
                flush_buf.fill_(k)  # evict the state / model from L2 (buffer > 126 MB L2)
            evs[k][0].record(stream)
            if nu:
                sim.set_device("ctrl", ctrl_dev[W + k].data_ptr(), nu)
            kev[k][0].record(stream)
            sim.step(1)
            kev[k][1].record(stream)
            evs[k][1].record(stream)
        barrier()
    ps_launches = sim.launch_info()["launches"] - launches0  # the handle's own count of kernels it launched
    ps_step_ms = sum(a.elapsed_time(b) for a, b in evs)
    ps_kern_ms = sum(a.elapsed_time(b) for a, b in kev)
    stats = {k: sim.get(k)[:, 0] for k in ("ncon", "nefc", "solver_iter")}

    # ---------------- device-resident leg B (headline): fused open-loop rollout ----------------
    # One b2mj_rollout launch advances every env K steps; the ctrl stream for all steps is resident in HBM
    # (K*nenv*nu*8 bytes, larger than L2 at the default K) and the per-step qpos/qvel/sensordata trajectory
    # is written back to HBM, so every step's result stays observable.
    sim.reset()
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    tq = torch.empty(K, nenv, model.nq, dtype=torch.float64, device=dev)
    tv = torch.empty(K, nenv, model.nv, dtype=torch.float64, device=dev)
    ts = torch.empty(K, nenv, max(ns, 1), dtype=torch.float64, device=dev) if ns else None
    cptr = lambda t: t.data_ptr() if t is not None else 0  # noqa: E731
    with torch.cuda.stream(stream):
        sim.rollout(W, cptr(ctrl_dev[:W]) if nu else 0, tq.data_ptr(), tv.data_ptr(), cptr(ts))
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        if not args.no_flush:
            flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        launches1 = sim.launch_info()["launches"]
        e0.record(stream)
        sim.rollout(K, cptr(ctrl_dev[W:]) if nu else 0, tq.data_ptr(), tv.data_ptr(), cptr(ts))
        e1.record(stream)
        rollout_launches = sim.launch_info()["launches"] - launches1
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop()
    step_ms = e0.elapsed_time(e1)
    kern_ms = step_ms
    t = torch.tensor([step_ms, kern_ms, ps_step_ms, ps_kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms, ps_step_ms, ps_kern_ms = (float(x) for x in t)
    value = nenv * world * K / (step_ms * 1e-3)
    ps_value = nenv * world * K / (ps_step_ms * 1e-3)
    warn = sim.get("warning").sum(0).tolist()
    traj_ok = bool(torch.isfinite(tq[-1]).all().item())

    # ---------------- end-to-end leg: host buffers through the C-ABI ----------------
    KE = args.e2e_steps or K
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()  # noqa: E731
    h_ctrl = pin(nenv, nu) if nu else None
    h_qpos, h_qvel, h_sens = pin(nenv, model.nq), pin(nenv, model.nv), pin(nenv, max(ns, 1))
    sim.reset()
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    ctrl_e2e = ctrl[W:W + KE] if KE <= K else np.resize(ctrl, (KE, nenv, nu))

    def e2e_step(k):
        if nu:
            np.copyto(h_ctrl, ctrl_e2e[k])  # the producer's write into the pinned staging buffer
        # one C-ABI call: ctrl H2D -> step -> qpos / qvel / sensordata D2H -> one synchronisation
        sim.step_host(1, h_ctrl if nu else None, h_qpos, h_qvel, h_sens if ns else None)

    for k in range(min(W, KE)):
        e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(KE):
        e2e_step(k)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nenv * world * KE / float(te[0])
    h2d = nenv * nu * 8
    d2h = nenv * (model.nq + model.nv + ns) * 8

    if rank == 0:
        peak, peak_src = load_peaks()
        bstate = state_bytes(model)
        kernel_s = kern_ms * 1e-3          # one launch = nenv * K env-steps
        achieved = bstate * nenv * K / kernel_s / 1e9
        info = sim.launch_info()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": KE, "timing": "wall clock, sync both sides, pinned host buffers, one b2mj_step_host call per step"},
            "gpu_launches": rollout_launches,
            "per_step_launch": {"value": ps_value, "unit": UNIT, "ms_per_step": ps_step_ms / K,
                                "kernel_ms_per_launch": ps_kern_ms / K, "gpu_launches": ps_launches,
                                "note": "K x (b2mj_set_device(ctrl) + b2mj_step(1) + launch-order refresh), CUDA events per step, "
                                        "L2 flushed between steps"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(args.model, nenv * K), "peak_source": peak_src, "kernel": "b2k_step_kernel",
                         "traffic_source": "profiles/r1_traffic.json (ncu --set full of the rollout launch, scaled per env-step)",
                         "algorithmic_bytes_per_env_step": bstate, "env_steps_per_launch": nenv * K,
                         "kernel_ms_per_launch": kern_ms,
                         "note": "latency/FP64-bound path: see DESIGN.md 'Roofline'"},
            "kernel": {k: info[k] for k in ("warps_per_cta", "ctas", "smem_bytes_per_cta", "regs_per_thread",
                                            "arena_in_smem", "state_record_bytes")},
            "trajectory_finite": traj_ok,
            "workload_stats": {"warnings": warn, "ncon_mean": float(stats["ncon"].mean()),
                               "nefc_mean": float(stats["nefc"].mean()), "nefc_max": int(stats["nefc"].max()),
                               "solver_iter_mean": float(stats["solver_iter"].mean())},
            "wall_s_timed_region": wall,
        }
        if not args.no_cpu and world == 1:  # CPU baseline: rank 0 at N=1 only (bounded sample)
            cores = os.cpu_count() or 1
            csteps = 50
            secs = cpu_rollout(model, nenv, csteps, args.seed, cores)
            # scale the sample to ~10 s of CPU work, bounded by the GPU arm's step count
            csteps = int(min(max(csteps, 10.0 / max(secs / csteps, 1e-9)), max(K, 50), 2000))
            secs = cpu_rollout(model, nenv, csteps, args.seed, cores)
            secs1 = cpu_rollout(model, max(64, nenv // 16), csteps, args.seed, 1)
            line["cpu_baseline"] = {
                "value": nenv * csteps / secs, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{nenv} envs x {csteps} steps of the same workload, oracle restatement of mj_step "
                          f"(not libmujoco), {cores} threads",
                "single_thread_value": max(64, nenv // 16) * csteps / secs1,
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
