"""CPU-side checks of the inputs the GPU parity tests rely on: the frozen reference MJCF fixture is the reference's
files byte for byte (when /root/reference is present), and the solver-matrix models really reach the three PGS size
classes of the CUDA solver (<= 32 rows, 33..64, > 64)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, model_path

REF = "/root/reference"


def test_ref_model_fixture_is_verbatim():
    g = np.load(os.path.join(GOLDEN, "ref_models.npz"))
    names = sorted(k[:-5] for k in g.files if k.endswith("__xml"))
    assert names == ["empty_world", "equality_world", "mocap_world", "pendulum_world", "sensors_world"]
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present on this box")
    for n in names:
        raw = open(os.path.join(REF, str(g[f"{n}__path"])), "rb").read()
        assert bytes(g[f"{n}__xml"]) == raw, n


def test_ref_models_compile_and_oracle_reproduces_fixture(capi, orc):
    g = np.load(os.path.join(GOLDEN, "ref_models.npz"))
    for n in ("pendulum_world", "equality_world", "mocap_world", "sensors_world", "empty_world"):
        m = capi.Model.from_xml_string(bytes(g[f"{n}__xml"]).decode())
        o = orc.Oracle(m)
        o.step(200)
        np.testing.assert_array_equal(o.get("qpos"), g[f"{n}__qpos"][-1])
        np.testing.assert_array_equal(o.get("qvel"), g[f"{n}__qvel"][-1])


def test_pgs_size_classes_are_reached(capi, orc):
    seen = {}
    for name, cone in (("box_stack.xml", 0), ("humanoid_like.xml", 0), ("hand_like.xml", 0), ("bin.xml", 1)):
        m = capi.Model.from_xml_file(model_path(name))
        m.opt.solver, m.opt.cone = 0, cone
        o = orc.Oracle(m)
        mx = 0
        for _ in range(150):
            o.step(1)
            mx = max(mx, int(o.get("nefc")[0]))
        seen[name] = mx
    assert any(32 < seen[k] <= 64 for k in ("box_stack.xml", "humanoid_like.xml", "hand_like.xml")), seen
    assert seen["bin.xml"] > 64, seen
