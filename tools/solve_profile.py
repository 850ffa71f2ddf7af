"""Developer diagnostic (GPU box): cycle split inside the primal (Newton / CG) solver.
Needs the profiling build:  make lib BUILD=build_prof EXTRA=-DB2K_SOLVE_PROF LIB=$PWD/mujoco_ros_pkgs_b200/libb2mj_prof.so
Run:  B2MJ_LIB=$PWD/mujoco_ros_pkgs_b200/libb2mj_prof.so python tools/solve_profile.py --model bin.xml --nenv 512"""
import argparse, ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="bin.xml"); ap.add_argument("--nenv", type=int, default=512)
ap.add_argument("--steps", type=int, default=20); ap.add_argument("--warm", type=int, default=300)
a = ap.parse_args()
model = _capi.Model.from_xml_file(os.path.join(bench.ROOT, "mujoco_ros_pkgs_b200", "models", a.model))
qpos, qvel, ctrl = bench.make_inputs(model, a.nenv, a.steps + a.warm, 1, amp=0.05)
sim = BatchSim(model, a.nenv); sim.set("qpos", qpos); sim.set("qvel", qvel)
for k in range(a.warm):
    if model.nu: sim.set("ctrl", ctrl[k])
    sim.step(1)
buf = (C.c_ulonglong * 16)()
_capi.lib.b2k_sprof_read(buf, 1)
for k in range(a.steps):
    if model.nu: sim.set("ctrl", ctrl[a.warm + k])
    sim.step(1)
_capi.lib.b2k_sprof_read(buf, 0)
names = ["init Ma/Jaref", "primalUpdate", "primalHessian", "primalGradient", "primalSearch", "team: zero + M", "team: J'DJ accumulate", "team: Cholesky"]
n = a.nenv * a.steps
print(f"{a.model} {a.nenv} envs: solver cycles per env-step; iters mean {sim.get('solver_iter').mean():.2f} nefc mean {sim.get('nefc').mean():.1f}")
for i, nm in enumerate(names):
    print(f"  {nm:16s} {buf[i] / n:12.0f}")
