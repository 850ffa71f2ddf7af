"""Developer diagnostic: per-step launch time and per-env residency as the batch grows (C2 workload, contact-rich state).
Usage: python tools/batch_probe.py nenv [nenv ...]   (env B2MJ_NO_REORDER=1 etc. apply)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
model = _capi.Model.from_xml_file(os.path.join(bench.ROOT, "mujoco_ros_pkgs_b200", "models", "panda_like.xml"))
for nenv in [int(a) for a in sys.argv[1:]]:
    qpos, qvel, ctrl = bench.make_inputs(model, nenv, 1040, 1)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos); sim.set("qvel", qvel)
    cdev = torch.from_numpy(ctrl).cuda()
    sim.rollout(1000, cdev[:1000].data_ptr()); sim.sync()
    ts, cyc = [], []
    for k in range(30):
        sim.set_device("ctrl", cdev[1000 + k].data_ptr(), model.nu)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); sim.step(1); ev1.record(); sim.sync()
        if k >= 10:
            ts.append(ev0.elapsed_time(ev1)); cyc.append(sim.env_cycles().astype(float) * 1024)
    c = np.concatenate(cyc)
    info = sim.launch_info()
    t = float(np.median(ts))
    print(f"nenv {nenv:7d}: step {t * 1e3:8.1f} us = {nenv / t / 1e3:6.2f} M env-steps/s; env cycles mean {c.mean():.0f} p50 {np.percentile(c, 50):.0f} "
          f"p99 {np.percentile(c, 99):.0f} max {c.max():.0f}; sum(env cycles)/kernel cycles = {c.sum() / len(cyc) / (t * 1.965e6):.0f} envs in flight; {info}")
    del sim
