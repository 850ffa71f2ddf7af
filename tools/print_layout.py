"""Developer diagnostic (GPU box): print the shared-memory / L2 placement of a model's per-env arrays.
Usage: B2MJ_PRINT_LAYOUT=1 [B2MJ_ENVS_PER_SM=n] python tools/print_layout.py model.xml nenv"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B2MJ_PRINT_LAYOUT", "1")
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
m = _capi.Model.from_xml_file(os.path.join(here, "mujoco_ros_pkgs_b200", "models", sys.argv[1]))
s = BatchSim(m, int(sys.argv[2]))
print(sys.argv[1], "njmax", m.njmax, "nconmax", m.nconmax, s.launch_info())
