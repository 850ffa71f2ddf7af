"""Per-env residency of single step launches for any bench model: who sets the length of a launch.
Usage: python tools/imbalance_model.py <model.xml> <nenv> [preroll steps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
name, nenv = sys.argv[1], int(sys.argv[2])
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 300
model = _capi.Model.from_xml_file(os.path.join(bench.ROOT, "mujoco_ros_pkgs_b200", "models", name))
qpos, qvel, ctrl = bench.make_inputs(model, nenv, pre + 8, 1)
sim = BatchSim(model, nenv)
sim.set("qpos", qpos); sim.set("qvel", qvel)
cdev = torch.from_numpy(ctrl).cuda() if model.nu else None
for k in range(pre):
    if model.nu: sim.set_device("ctrl", cdev[k].data_ptr(), model.nu)
    sim.step(1)
sim.sync()
print(name, nenv, "envs", sim.launch_info())
for k in range(3):
    if model.nu: sim.set_device("ctrl", cdev[pre + k].data_ptr(), model.nu)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); sim.step(1); ev1.record(); sim.sync()
    c = sim.env_cycles().astype(float)
    it = sim.get("solver_iter")[:, 0]; ne = sim.get("nefc")[:, 0]
    print(f"step {k}: kernel {ev0.elapsed_time(ev1) * 1e3:.0f} us = {ev0.elapsed_time(ev1) * 1.965e6:.0f} cycles; env cycles mean {c.mean():.0f} "
          f"p50 {np.percentile(c, 50):.0f} p90 {np.percentile(c, 90):.0f} p99 {np.percentile(c, 99):.0f} max {c.max():.0f}; max/mean {c.max() / c.mean():.2f}")
    idx = np.argsort(c)[-6:]
    print("   slowest envs: cycles", c[idx].astype(int).tolist(), "nefc", ne[idx].tolist(), "iters", it[idx].tolist())
    print("   nefc mean", ne.mean(), "max", ne.max(), "iters mean", it.mean(), "max", it.max(),
          " corr(cycles, nefc*iters)", np.corrcoef(c, ne * it)[0, 1])
