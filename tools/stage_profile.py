"""Developer diagnostic (GPU box): per-stage SM-cycle profile of the fused step kernel.
Usage: python tools/stage_profile.py [--model panda_like.xml] [--nenv 4096] [--steps 200] [--warm 100]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mujoco_ros_pkgs_b200 import _capi  # noqa: E402
from mujoco_ros_pkgs_b200.batch import BatchSim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="panda_like.xml")
ap.add_argument("--nenv", type=int, default=4096)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--warm", type=int, default=100)
a = ap.parse_args()
model = _capi.Model.from_xml_file(os.path.join(bench.ROOT, "mujoco_ros_pkgs_b200", "models", a.model))
qpos, qvel, ctrl = bench.make_inputs(model, a.nenv, a.steps + a.warm, 1)
sim = BatchSim(model, a.nenv)
sim.set("qpos", qpos)
sim.set("qvel", qvel)
for k in range(a.warm):
    if model.nu:
        sim.set("ctrl", ctrl[k])
    sim.step(1)
sim.stage_profile(True)
for k in range(a.steps):
    if model.nu:
        sim.set("ctrl", ctrl[a.warm + k])
    sim.step(1)
prof = sim.stage_profile(False)
tot = sum(prof.values())
n = a.nenv * a.steps
print(f"{a.model}: {a.nenv} envs x {a.steps} steps; launch {sim.launch_info()}")
print(f"cycles per env-step (warp-resident time): {tot / n:.0f}")
for k, v in prof.items():
    print(f"  {k:22s} {v / n:10.0f} cyc  {100.0 * v / max(tot, 1):5.1f}%")
print("nefc mean", sim.get("nefc").mean(), "ncon mean", sim.get("ncon").mean(), "iters mean", sim.get("solver_iter").mean())
