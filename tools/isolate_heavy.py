"""Developer diagnostic: do the heaviest envs of a per-step launch run faster when nothing shares their SM?
Runs the C2 workload to its contact-rich state, probes one step, then replays that same step for the 64 heaviest envs
alone (one env per SM) and prints in-situ vs isolated residency."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
nenv = 4096
model = _capi.Model.from_xml_file(os.path.join(bench.ROOT, "mujoco_ros_pkgs_b200", "models", "panda_like.xml"))
qpos, qvel, ctrl = bench.make_inputs(model, nenv, 1100, 1)
sim = BatchSim(model, nenv)
sim.set("qpos", qpos); sim.set("qvel", qvel)
cdev = torch.from_numpy(ctrl).cuda()
os.environ["B2MJ_ROLLOUT_CHUNK"] = "0"
sim.rollout(1000, cdev[:1000].data_ptr()); sim.sync()
for k in range(3):   # let the launch order settle
    sim.set("ctrl", ctrl[1000 + k]); sim.step(1)
st = {k: sim.get(k) for k in ("qpos", "qvel", "qacc_warmstart", "time")}
c = ctrl[1003]
sim.set("ctrl", c); sim.step(1); sim.sync()
cyc = sim.env_cycles(); ne = sim.get("nefc")[:, 0]; it = sim.get("solver_iter")[:, 0]
idx = np.argsort(cyc)[-64:]
small = BatchSim(model, 64)
for k, v in st.items():
    small.set(k, v[idx])
small.set("ctrl", c[idx]); small.step(1); small.sync()
cyc2 = small.env_cycles(); ne2 = small.get("nefc")[:, 0]; it2 = small.get("solver_iter")[:, 0]
assert np.array_equal(ne[idx], ne2) and np.array_equal(it[idx], it2)
rows = (ne[idx] * it[idx]).astype(float)
print("heaviest 64 envs: rows (nefc x iters) min/median/max", rows.min(), np.median(rows), rows.max())
print("in situ   cycles: median %.0f max %.0f" % (np.median(cyc[idx]), cyc[idx].max()))
print("isolated  cycles: median %.0f max %.0f   (launch %s)" % (np.median(cyc2), cyc2.max(), small.launch_info()))
b1, a1 = np.polyfit(rows, cyc[idx].astype(float), 1)
b2, a2 = np.polyfit(rows, cyc2.astype(float), 1)
print("fit in situ : %.0f + %.1f x rows" % (a1, b1))
print("fit isolated: %.0f + %.1f x rows" % (a2, b2))
light = np.argsort(cyc)[:64]
small2 = BatchSim(model, 64)
for k, v in st.items():
    small2.set(k, v[light])
small2.set("ctrl", c[light]); small2.step(1); small2.sync()
print("lightest 64 envs: in situ median %.0f  isolated median %.0f" % (np.median(cyc[light]), np.median(small2.env_cycles())))
