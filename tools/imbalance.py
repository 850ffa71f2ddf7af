"""Developer diagnostic: distribution of per-env work-item residency (cycles) for per-step launches and
for a fused rollout, after the workload has reached steady state."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
nenv = 4096
model = _capi.Model.from_xml_file(os.path.join(bench.ROOT, "mujoco_ros_pkgs_b200", "models", "panda_like.xml"))
qpos, qvel, ctrl = bench.make_inputs(model, nenv, 1400, 1)
sim = BatchSim(model, nenv)
sim.set("qpos", qpos); sim.set("qvel", qvel)
cdev = torch.from_numpy(ctrl).cuda()
os.environ["B2MJ_ROLLOUT_CHUNK"] = "0"
sim.rollout(1000, cdev[:1000].data_ptr()); sim.sync()
def show(tag, c):
    print(f"{tag}: mean {c.mean():.0f} p50 {np.percentile(c,50):.0f} p90 {np.percentile(c,90):.0f} p99 {np.percentile(c,99):.0f} max {c.max():.0f} cycles; max/mean {c.max()/c.mean():.2f}")
for k in range(3):
    sim.set_device("ctrl", cdev[1000 + k].data_ptr(), model.nu); sim.step(1); sim.sync()
    show(f"single step {k}", sim.env_cycles())
    it = sim.get("solver_iter")[:, 0]; ne = sim.get("nefc")[:, 0]; c = sim.env_cycles()
    idx = np.argsort(c)[-5:]
    print("   slowest envs: cycles", c[idx].tolist(), "nefc", ne[idx].tolist(), "iters", it[idx].tolist())
    b, a = np.polyfit((ne * it).astype(float), c.astype(float), 1)
    print(f"   fit cycles = {a:.0f} + {b:.1f} x (nefc x iters)   [cycles per PGS row update in situ = {b:.1f}]")
    big = ne * it >= 1000
    if big.any():
        print(f"   envs with nefc x iters >= 1000: {int(big.sum())}, (cycles - {a:.0f}) / rows: median {np.median((c[big] - a) / (ne * it)[big]):.1f}")
    print("   corr(cycles, nefc*iters) =", np.corrcoef(c, ne * it)[0, 1], " base (nefc==0):", c[ne == 0].mean() if (ne == 0).any() else None,
          " nefc<=1:", c[ne <= 1].mean())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); sim.rollout(300, cdev[1003:1303].data_ptr()); ev1.record(); sim.sync()
c = sim.env_cycles()
show("rollout 300 steps (per env total)", c)
print("   kernel ms", ev0.elapsed_time(ev1), " sum(env cycles)/(148*13 slots) ms:", c.sum() / (148 * 13) / 1.965e6, " max env ms:", c.max() / 1.965e6)
