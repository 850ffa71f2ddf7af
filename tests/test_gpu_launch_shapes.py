"""Launch shapes that only appear at the BASELINE batch sizes and beyond (handle.cu::handle_launch, make_layout): the
lock-stepped wide CTA of per-step launches longer than 2.5 waves, the launch order refreshed from measured cost (on a side
stream for small models), stage barriers between the CTA mates of Newton models, the residency trade that brings the
Newton working set on chip.  None of them may change a result: a subsample of a large batch (first, middle, last and the
heaviest envs) is stepped against the CPU oracle from injected states, the north star's per-step 1e-5 bar."""
import os

import numpy as np
import pytest

from conftest import MODELS
from parity_util import STATE_FIELDS, TOL, ctrl_sample, perturbed, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,nenv,pre", [("panda_like.xml", 6144, 300), ("hand_like.xml", 1536, 60), ("humanoid_like.xml", 2048, 60)])
def test_large_batch_subsample_matches_oracle(capi, orc, name, nenv, pre):
    from mujoco_ros_pkgs_b200.batch import BatchSim

    model = capi.Model.from_xml_file(os.path.join(MODELS, name))
    qpos, qvel = perturbed(model, nenv, seed=3, amp=0.05)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    rng = np.random.default_rng(9)
    for _ in range(pre):  # into contact, and through enough launches for the cost-ordered launch slots to be in use
        sim.set("ctrl", ctrl_sample(model, rng, nenv))
        sim.step(1)
    nefc = sim.get("nefc")[:, 0]
    heavy = np.argsort(nefc)[-4:].tolist()
    sub = sorted(set([0, 1, 2, nenv // 2, nenv // 2 + 1, nenv - 2, nenv - 1] + heavy))
    oracles = {e: orc.Oracle(model) for e in sub}
    worst = 0.0
    for s in range(12):
        st = {k: sim.get(k) for k in STATE_FIELDS if model.field_size_by_name(k) > 0}
        ctrl = ctrl_sample(model, rng, nenv)
        sim.set("ctrl", ctrl)
        sim.step(1)
        out = {k: sim.get(k) for k in ("qpos", "qvel", "qacc")}
        nefc = sim.get("nefc")[:, 0]
        for e, o in oracles.items():
            for k, v in st.items():
                o.set(k, v[e])
            o.set("ctrl", ctrl[e])
            o.step(1)
            assert int(o.get("nefc")[0]) == int(nefc[e]), (name, s, e)
            for k, v in out.items():
                err = rel(v[e], o.get(k))
                worst = max(worst, err)
                assert err < TOL, f"{name} step {s} env {e} field {k}: {err:.3e} (nefc {nefc[e]})"
    info = sim.launch_info()
    print(f"{name} x {nenv}: worst {worst:.1e}, launch {info}")
    assert np.all(np.isfinite(sim.get("qpos")))
