"""Developer diagnostic (run on a GPU box): field-by-field diff of the CUDA batched step against the
CPU oracle.  Usage: python tools/gpu_check.py [model.xml ...] [--nenv N] [--steps K] [--pert P]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_ros_pkgs_b200 import _capi  # noqa: E402
from mujoco_ros_pkgs_b200.batch import BatchSim  # noqa: E402
from oracle import binding as ob  # noqa: E402

MODELS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mujoco_ros_pkgs_b200", "models")

SKIP = {"efc_AR", "xfrc_applied", "warning"}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / (1e-9 + np.maximum(np.abs(a), np.abs(b))))) if False else float(
        np.max(np.abs(a - b)) / (1e-12 + np.max(np.abs(b))))


def perturb(model, nenv, pert, seed=1234):
    rng = np.random.default_rng(seed)
    qpos = np.tile(model.qpos0, (nenv, 1))
    qvel = np.zeros((nenv, model.nv))
    for j in range(model.njnt):
        t, qa, da = model.jnt_type[j], model.jnt_qposadr[j], model.jnt_dofadr[j]
        if t == 0:  # free
            qpos[:, qa:qa + 3] += rng.uniform(-pert, pert, (nenv, 3)) * np.array([1, 1, 0.2])
            q = qpos[:, qa + 3:qa + 7] + rng.uniform(-pert, pert, (nenv, 4))
            qpos[:, qa + 3:qa + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
            qvel[:, da:da + 6] = rng.uniform(-pert, pert, (nenv, 6))
        elif t == 1:  # ball
            q = qpos[:, qa:qa + 4] + rng.uniform(-pert, pert, (nenv, 4))
            qpos[:, qa:qa + 4] = q / np.linalg.norm(q, axis=1, keepdims=True)
            qvel[:, da:da + 3] = rng.uniform(-pert, pert, (nenv, 3))
        else:
            qpos[:, qa] += rng.uniform(-pert, pert, nenv)
            qvel[:, da] = rng.uniform(-pert, pert, nenv)
    return qpos, qvel


def check_model(path, nenv, steps, pert, verbose):
    model = _capi.Model.from_xml_file(path)
    print(f"== {os.path.basename(path)} nq={model.nq} nv={model.nv} nu={model.nu} nbody={model.nbody} "
          f"ngeom={model.ngeom} ncollpair={model.ncollpair} nconmax={model.nconmax} njmax={model.njmax} "
          f"solver={model.opt.solver} integ={model.opt.integrator} cone={model.opt.cone}")
    sim = BatchSim(model, nenv)
    print("   launch:", sim.launch_info())
    qpos, qvel = perturb(model, nenv, pert)
    rng = np.random.default_rng(7)
    ctrl = np.zeros((nenv, model.nu))
    if model.nu:
        lo, hi = model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1]
        lim = model.actuator_ctrllimited.astype(bool)
        lo = np.where(lim, lo, -1.0)
        hi = np.where(lim, hi, 1.0)
        ctrl = rng.uniform(lo, hi, (nenv, model.nu))
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    if model.nu:
        sim.set("ctrl", ctrl)
    sim.keep_intermediates(True)
    sim.forward()
    sim.sync()
    oracles = []
    for e in range(nenv):
        o = ob.Oracle(model)
        o.set("qpos", qpos[e])
        o.set("qvel", qvel[e])
        if model.nu:
            o.set("ctrl", ctrl[e])
        o.forward()
        oracles.append(o)
    worst = {}
    ncon_g = sim.get("ncon")[:, 0]
    nefc_g = sim.get("nefc")[:, 0]
    ncon_o = np.array([o.get("ncon")[0] for o in oracles])
    nefc_o = np.array([o.get("nefc")[0] for o in oracles])
    print(f"   forward: ncon gpu/orc max {ncon_g.max()}/{ncon_o.max()} mismatch {np.sum(ncon_g != ncon_o)}; "
          f"nefc max {nefc_g.max()}/{nefc_o.max()} mismatch {np.sum(nefc_g != nefc_o)}")
    for name in _capi.FIELD_NAMES:
        if name in SKIP:
            continue
        n, is_int = model.field_size(_capi.field_id(name))
        if n <= 0:
            continue
        try:
            g = sim.get(name)
        except _capi.B2mjError as ex:
            print(f"   {name}: get failed: {ex}")
            continue
        errs = []
        for e in range(nenv):
            ov = oracles[e].get(name)
            gv = g[e]
            # only the live rows of contact / efc arrays are defined
            if name.startswith("contact_"):
                per = n // model.nconmax
                k = per * min(ncon_o[e], ncon_g[e])
                ov, gv = ov[:k], gv[:k]
            elif name.startswith("efc_"):
                per = n // model.njmax
                k = per * min(nefc_o[e], nefc_g[e])
                ov, gv = ov[:k], gv[:k]
            if is_int:
                errs.append(float(np.sum(ov != gv)))
            else:
                errs.append(relerr(gv, ov))
        worst[name] = max(errs) if errs else 0.0
    bad = {k: v for k, v in worst.items() if v > 1e-9}
    if verbose:
        for k, v in worst.items():
            print(f"      {k:24s} {v:.3e}")
    print(f"   forward fields: {len(worst)} compared, {len(bad)} above 1e-9: "
          + ", ".join(f"{k}={v:.2e}" for k, v in sorted(bad.items(), key=lambda kv: -kv[1])[:12]))
    # stepping divergence
    sim.keep_intermediates(False)
    t0 = time.time()
    maxrel = 0.0
    report_at = {1, 10, 100, steps}
    for s in range(1, steps + 1):
        if model.nu:
            ctrl = rng.uniform(lo, hi, (nenv, model.nu))
            sim.set("ctrl", ctrl)
        sim.step(1)
        for e, o in enumerate(oracles):
            if model.nu:
                o.set("ctrl", ctrl[e])
            o.step(1)
        if s in report_at:
            gq, gv = sim.get("qpos"), sim.get("qvel")
            oq = np.stack([o.get("qpos") for o in oracles])
            ov = np.stack([o.get("qvel") for o in oracles])
            rq = float(np.max(np.abs(gq - oq) / (1.0 + np.abs(oq))))
            rv = float(np.max(np.abs(gv - ov) / (1.0 + np.abs(ov))))
            gt = sim.get("time")[:, 0]
            ot = np.array([o.time for o in oracles])
            ng = sim.get("ncon")[:, 0]
            no = np.array([o.get("ncon")[0] for o in oracles])
            it = sim.get("solver_iter")[:, 0]
            ito = np.array([o.get("solver_iter")[0] for o in oracles])
            print(f"   step {s:5d}: qpos err {rq:.3e} qvel err {rv:.3e} time bitwise {np.array_equal(gt, ot)} "
                  f"ncon mismatch {np.sum(ng != no)} (max {ng.max()}) iters gpu {it.max()} orc {ito.max()} "
                  f"warn {sim.get('warning').sum(0).tolist()}")
            maxrel = max(maxrel, rq, rv)
    print(f"   {steps} steps checked in {time.time() - t0:.1f}s; worst state err {maxrel:.3e}")
    return len(bad), maxrel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("models", nargs="*")
    ap.add_argument("--nenv", type=int, default=16)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--pert", type=float, default=0.1)
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    paths = a.models or sorted(os.path.join(MODELS, f) for f in os.listdir(MODELS) if f.endswith(".xml"))
    for p in paths:
        if not os.path.exists(p):
            p = os.path.join(MODELS, p)
        try:
            check_model(p, a.nenv, a.steps, a.pert, a.v)
        except Exception as ex:  # keep going: this is a diagnostic
            print(f"   FAILED: {type(ex).__name__}: {ex}")


if __name__ == "__main__":
    main()
