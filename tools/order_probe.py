"""Developer diagnostic: a few multi-wave steps so that the launch-order kernel runs (used under compute-sanitizer)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
m = _capi.Model.from_xml_file(os.path.join(here, "mujoco_ros_pkgs_b200", "models", "panda_like.xml"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4500
rng = np.random.default_rng(1)
sim = BatchSim(m, n)
sim.set("qpos", np.tile(m.qpos0, (n, 1)) + rng.uniform(-0.3, 0.3, (n, m.nq)))
for k in range(6):
    sim.step(1)
sim.sync()
print("ok", sim.launch_info())
