"""bench.py's reference arm runs on the CPU: check the JSON contract keys the driver reads and that the pre-roll really
moves the workload into its contact-rich regime (so a short driver run does not time the contact-free prefix)."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT, model_path


def test_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3",
                        "--preroll", "10", "--nenv", "32"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["steps"] == 4 and line["unit"] == "env-steps/s"
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["preroll_steps"] == 10


def test_preroll_reaches_contact_regime(capi, orc):
    """C2 workload: the batch starts contact free (no contacts at all over the first steps) and only after a few
    hundred steps of random control do arm / gripper / floor contacts appear (a tail of envs with 9..25 constraint rows
    and tens of PGS iterations).  The timed legs must see that regime, hence the pre-roll."""
    sys.path.insert(0, ROOT)
    import bench

    model = capi.Model.from_xml_file(model_path("panda_like.xml"))
    nenv, P = 64, 1000
    qpos, qvel, ctrl = bench.make_inputs(model, nenv, P, 1)
    ncon0, ncon1, nefc1, it1 = [], [], [], []
    for e in range(nenv):
        o = orc.Oracle(model)
        o.set("qpos", qpos[e])
        for s in range(P):
            o.set("ctrl", ctrl[s, e])
            o.step(1)
            if s == 20:
                ncon0.append(int(o.get("ncon")[0]))
        ncon1.append(int(o.get("ncon")[0]))
        nefc1.append(int(o.get("nefc")[0]))
        it1.append(int(o.get("solver_iter")[0]))
    print("ncon at step 20:", np.mean(ncon0), " after pre-roll: ncon", np.mean(ncon1), "nefc max", max(nefc1), "iter max", max(it1))
    assert max(ncon0) == 0
    assert np.mean(ncon1) > 0.1 and max(nefc1) >= 8 and max(it1) >= 30, (ncon1, nefc1, it1)
