"""Developer diagnostic: fused-rollout and per-step throughput of the C2 workload from its pre-rolled state."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
nenv, P, K = 4096, 1000, 300
model = _capi.Model.from_xml_file(os.path.join(bench.ROOT, "mujoco_ros_pkgs_b200", "models", "panda_like.xml"))
qpos, qvel, ctrl = bench.make_inputs(model, nenv, P + 2 * K, 1)
sim = BatchSim(model, nenv)
sim.set("qpos", qpos); sim.set("qvel", qvel)
cdev = torch.from_numpy(ctrl).cuda()
sim.rollout(P, cdev[:P].data_ptr()); sim.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sim.rollout(K, cdev[P:P + K].data_ptr()); e1.record(); sim.sync()
ms = e0.elapsed_time(e1)
e0.record()
for k in range(K):
    sim.set_device("ctrl", cdev[P + K + k].data_ptr(), model.nu); sim.step(1)
e1.record(); sim.sync()
ms2 = e0.elapsed_time(e1)
li = sim.launch_info()
print(f"{os.environ.get('TAG','')}: rollout {nenv*K/ms/1e3:.2f} M/s  per-step {nenv*K/ms2/1e3:.2f} M/s  W={li['warps_per_cta']} smem/cta={li['smem_bytes_per_cta']}")
