"""Freeze the reference's own MJCF test assets as a fixture (run in the build container only).

/root/reference does not exist on the GPU box, so the GPU parity tests cannot open the reference's XML files.
This script reads them where they lie and stores their bytes, unmodified, in tests/golden/ref_models.npz together
with oracle outputs (forward fields at qpos0 and a 200-step trajectory) for each; tests/test_gpu_paths.py compiles the
stored text with b2mj_model_from_xml_string.  tests/test_model_compiler.py re-checks fixture == file whenever
/root/reference is present.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
FILES = {
    "pendulum_world": "mujoco_ros/assets/pendulum_world.xml",
    "empty_world": "mujoco_ros/test/empty_world.xml",
    "equality_world": "mujoco_ros/test/equality_world.xml",
    "mocap_world": "mujoco_ros_mocap_plugin/assets/mocap_world.xml",
    "sensors_world": "mujoco_ros_sensors/test/sensors_world.xml",
}


def main():
    from mujoco_ros_pkgs_b200 import _capi
    from oracle import binding as ob

    out = {}
    for name, rel in FILES.items():
        raw = open(os.path.join(REF, rel), "rb").read()
        out[f"{name}__xml"] = np.frombuffer(raw, dtype=np.uint8)
        out[f"{name}__path"] = np.array(rel)
        m = _capi.Model.from_xml_string(raw.decode())
        o = ob.Oracle(m)
        traj_q, traj_v = [], []
        for _ in range(200):
            o.step(1)
            traj_q.append(o.get("qpos"))
            traj_v.append(o.get("qvel"))
        out[f"{name}__qpos"] = np.array(traj_q)
        out[f"{name}__qvel"] = np.array(traj_v)
        out[f"{name}__time"] = np.array([o.time])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_models.npz"), **out)
    print("wrote tests/golden/ref_models.npz:", sorted(FILES))


if __name__ == "__main__":
    main()
