"""Run a few batched steps so ncu can capture the step kernel (used under `ncu -k regex:b2k_step`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
nenv = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
name = sys.argv[3] if len(sys.argv) > 3 else "panda_like.xml"
model = _capi.Model.from_xml_file(os.path.join(bench.ROOT, "mujoco_ros_pkgs_b200", "models", name))
qpos, qvel, ctrl = bench.make_inputs(model, nenv, nsteps, 1)
sim = BatchSim(model, nenv)
sim.set("qpos", qpos); sim.set("qvel", qvel)
for k in range(nsteps):
    if model.nu: sim.set("ctrl", ctrl[k])
    sim.step(1)
sim.sync()
print("done", sim.get("nefc").mean())
