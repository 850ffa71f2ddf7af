"""One-command parity pin against the real MuJoCo, for the day it is available.

The reference's physics is MuJoCo 2.3.7 (mujoco_ros/CMakeLists.txt:61), which is absent from /root/reference and from
this image, so the CPU oracle (oracle/) is "parity unpinned" for post-step dynamics.  This tool
closes that gap whenever a MuJoCo build can be imported -- the `mujoco` Python module, found on sys.path, under
$MUJOCO_DIR/python, or under baseline/_ref:

  1. loads every model XML of the repo (and the reference's own five MJCF files from tests/golden/ref_models.npz) in
     BOTH MuJoCo and this repo's compiler;
  2. diffs every mjModel array the two have in common (compiler + mj_setConst pin);
  3. diffs every mjData array of b2mj_field after mj_forward at a perturbed state (per-stage pin);
  4. rolls 1000 mj_step with the seeded control stream of the parity tests and diffs qpos / qvel per step;
  5. writes tests/golden/<model>_mujoco.npz (inputs + MuJoCo's outputs), which tests/test_libmujoco_pin.py then
     enforces on the oracle (CPU) and on the CUDA path (GPU) from that moment on.

Without MuJoCo it prints what it would do and exits 0 (no-op)."""
import argparse
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def find_mujoco():
    for extra in (None, os.path.join(os.environ.get("MUJOCO_DIR", ""), "python"), os.path.join(ROOT, "baseline", "_ref")):
        if extra and os.path.isdir(extra) and extra not in sys.path:
            sys.path.insert(0, extra)
        try:
            import mujoco  # noqa: F401

            return mujoco
        except Exception:
            continue
    return None


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    if a.size != b.size:
        return float("inf")
    return float(np.max(np.abs(a - b) / (1.0 + np.abs(b)))) if a.size else 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--tol", type=float, default=1e-5)
    args = ap.parse_args()
    mj = find_mujoco()
    if mj is None:
        print("compare_libmujoco: no MuJoCo build importable (tried sys.path, $MUJOCO_DIR/python, baseline/_ref): nothing to do.\n"
              "The oracle stays parity-unpinned; install mujoco==2.3.7 and re-run this script to write tests/golden/*_mujoco.npz.")
        return 0
    from mujoco_ros_pkgs_b200 import _capi
    from oracle import binding as ob

    print(f"MuJoCo {mj.__version__} found" + ("" if mj.__version__.startswith("2.3.7") else " (reference pins 2.3.7: expect small differences)"))
    sources = {os.path.basename(p)[:-4]: open(p).read() for p in sorted(glob.glob(os.path.join(ROOT, "mujoco_ros_pkgs_b200", "models", "*.xml")))}
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_models.npz"))
    for k in g.files:
        if k.endswith("__xml"):
            sources["ref_" + k[:-5]] = bytes(g[k]).decode()
    # scenes of the CPU tests that lean on recalled MuJoCo behaviour: every constraint row type, contact dimensions
    # 1 / 4 / 6 with margin / gap / solmix / priority (incl. the normal-force coupling of saturated elliptic friction),
    # every disable flag, the sensor chain, springdamper + actuator shorthands
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_round2d_gpu
    import test_sensor_derivatives_cpu as tsd
    import test_solver_optimality_cpu as tso
    import test_zz_disable_flags as tdf
    sources.update({"t_rows": tso.ROWS, "t_condim_pyramidal": tso.CONDIM,
                    "t_condim_elliptic": tso.CONDIM.replace('impratio="3"', 'impratio="3" cone="elliptic"'),
                    "t_sensors": tsd.XML % "".join(tsd.PER_SITE.format(s) for s in tsd.SITES), "t_rest": tsd.REST,
                    "t_arm_shorthands": test_round2d_gpu.ARM % "Euler"})
    for flag in tdf.FLAGS:
        sources["t_flag_" + flag] = tdf.SCENE.format(flag=flag, solver="Newton")
    worst_all = 0.0
    for name, xml in sources.items():
        mm = mj.MjModel.from_xml_string(xml)
        md = mj.MjData(mm)
        ours = _capi.Model.from_xml_string(xml)
        report = []
        # ---- 2. model arrays
        for arr_name, arr in ours._arrays.items():
            if hasattr(mm, arr_name) and arr.dtype != np.uint8:
                theirs = np.asarray(getattr(mm, arr_name))
                if theirs.size == arr.size:
                    report.append((rel(arr, theirs), "model." + arr_name))
        # ---- 3. forward fields at a perturbed state
        rng = np.random.default_rng(11)
        qpos = mm.qpos0 + 0.0
        qvel = rng.uniform(-0.1, 0.1, mm.nv)
        for j in range(mm.njnt):
            if mm.jnt_type[j] in (2, 3):
                qpos[mm.jnt_qposadr[j]] += rng.uniform(-0.1, 0.1)
        ctrl = rng.uniform(-1, 1, mm.nu)
        md.qpos[:], md.qvel[:] = qpos, qvel
        if mm.nu:
            md.ctrl[:] = ctrl
        mj.mj_forward(mm, md)
        o = ob.Oracle(ours)
        o.set("qpos", qpos)
        o.set("qvel", qvel)
        if mm.nu:
            o.set("ctrl", ctrl)
        o.forward()
        for fname in _capi.FIELD_NAMES:
            if fname.startswith(("contact_", "efc_")) or fname in ("warning", "solver_iter", "ncon", "nefc"):
                continue
            if hasattr(md, fname) and ours.field_size_by_name(fname) > 0:
                report.append((rel(o.get(fname), np.asarray(getattr(md, fname))), "data." + fname))
        report.append((0.0 if int(o.get("ncon")[0]) == md.ncon else float("inf"), "data.ncon"))
        # ---- 4. rollout
        mj.mj_resetData(mm, md)
        md.qpos[:] = qpos
        o = ob.Oracle(ours)
        o.set("qpos", qpos)
        traj_q, traj_v, ctrls, div = [], [], [], 0.0
        for s in range(args.steps):
            c = rng.uniform(-1, 1, mm.nu)
            if mm.nu:
                md.ctrl[:] = c
                o.set("ctrl", c)
            mj.mj_step(mm, md)
            o.step(1)
            div = max(div, rel(o.get("qpos"), md.qpos), rel(o.get("qvel"), md.qvel))
            if s % 10 == 9:
                traj_q.append(md.qpos.copy()); traj_v.append(md.qvel.copy())
            ctrls.append(c)
        report.append((div, f"rollout {args.steps} steps (qpos, qvel, free running)"))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"{name}_mujoco.npz"), xml=np.frombuffer(xml.encode(), dtype=np.uint8),
                            qpos_init=qpos, ctrl=np.array(ctrls), qpos=np.array(traj_q), qvel=np.array(traj_v), stride=10,
                            mujoco_version=np.array(mj.__version__))
        bad = [(e, n) for e, n in report if not e < args.tol]
        worst = max(e for e, _ in report)
        worst_all = max(worst_all, worst)
        print(f"{name:24s} worst {worst:.2e}  " + ("OK" if not bad else "MISMATCH: " + ", ".join(f"{n} {e:.1e}" for e, n in sorted(bad, reverse=True)[:6])))
    print(f"wrote tests/golden/*_mujoco.npz; overall worst {worst_all:.2e} (tolerance {args.tol})")
    return 0 if worst_all < args.tol else 1


if __name__ == "__main__":
    sys.exit(main())
