"""Freeze oracle trajectories as small fixtures under tests/golden/ (run here; committed with its output).
The oracle is a restatement of MuJoCo 2.3.7's step (no libmujoco in this image: parity unpinned), so
these vectors pin the oracle against accidental change and give the GPU tests a box-independent target."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_ros_pkgs_b200 import _capi  # noqa: E402
from oracle import binding as ob  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
for name, nrec, stride in (("panda_like", 50, 10), ("pendulum_scene", 50, 10), ("equality_scene", 50, 10), ("box_stack", 60, 10), ("hand_like", 30, 10),
                           ("humanoid_like", 12, 5), ("bin", 15, 10)):
    m = _capi.Model.from_xml_file(os.path.join(ROOT, "mujoco_ros_pkgs_b200", "models", name + ".xml"))
    rng = np.random.default_rng(2024)
    qpos = m.qpos0.copy()
    qvel = np.zeros(m.nv)
    for j in range(m.njnt):
        t, qa, da = m.jnt_type[j], m.jnt_qposadr[j], m.jnt_dofadr[j]
        amp = 0.02 if name in ("hand_like", "humanoid_like", "bin") else 0.1
        if t >= 2:
            qpos[qa] += rng.uniform(-amp, amp)
            qvel[da] = rng.uniform(-5 * amp, 5 * amp)
        elif t == 1:
            q = qpos[qa:qa + 4] + rng.uniform(-0.1, 0.1, 4)
            qpos[qa:qa + 4] = q / np.linalg.norm(q)
    o = ob.Oracle(m)
    o.set("qpos", qpos)
    o.set("qvel", qvel)
    lo, hi = (m.actuator_ctrlrange[:, 0], m.actuator_ctrlrange[:, 1]) if m.nu else (None, None)
    Q, V, U = [], [], []
    for k in range(nrec):
        if m.nu:
            u = rng.uniform(lo, hi)
            o.set("ctrl", u)
            U.append(u)
        o.step(stride)
        Q.append(o.get("qpos").copy())
        V.append(o.get("qvel").copy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), qpos_init=qpos, qvel_init=qvel, stride=stride,
                        qpos=np.array(Q), qvel=np.array(V), ctrl=np.array(U) if U else np.zeros((nrec, 0)))
    print(name, "final qpos", Q[-1][:6])
