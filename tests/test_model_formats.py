"""Model-side formats beyond plain MJCF files: keyframes, binary model files, load-from-string,
extension dispatch.  The reference loads .xml / .mjb by extension and from a VFS string (mujoco_env.cpp:771-911) and
reaches keyframes through its viewer (viewer.cpp:1735-1751)."""
import os

import numpy as np
import pytest

from conftest import model_path

KEYED = """
<mujoco>
  <option timestep="0.002"/>
  <worldbody>
    <body name="mover" mocap="true" pos="0.5 0 1"><geom type="sphere" size="0.05" contype="0" conaffinity="0"/></body>
    <geom type="plane" size="2 2 0.1"/>
    <body pos="0 0 0.5">
      <freejoint name="root"/>
      <geom type="box" size="0.1 0.1 0.1"/>
      <body pos="0.2 0 0">
        <joint name="hinge" type="hinge" axis="0 1 0"/>
        <geom type="capsule" fromto="0 0 0 0.3 0 0" size="0.03"/>
      </body>
    </body>
  </worldbody>
  <actuator>
    <motor joint="hinge" name="m0"/>
    <general joint="hinge" name="filt" dyntype="filter" dynprm="0.05"/>
  </actuator>
  <keyframe>
    <key name="home"/>
    <key name="tilted" time="1.5" qpos="0.1 0.2 0.8 0.7071067811865476 0 0.7071067811865476 0 0.3"
         qvel="0 0 0 0 0 0 2.0" act="0.25" ctrl="0.5 -0.5" mpos="1 1 1" mquat="0 1 0 0"/>
  </keyframe>
</mujoco>
"""


def test_keyframes_compile(capi):
    m = capi.Model.from_xml_string(KEYED)
    assert m.nkey == 2 and m.nq == 8 and m.nv == 7 and m.na == 1 and m.nu == 2 and m.nmocap == 1
    assert m.name2id(capi.OBJ_KEY, "home") == 0 and m.name2id(capi.OBJ_KEY, "tilted") == 1
    assert m.id2name(capi.OBJ_KEY, 1) == "tilted"
    kq = m.key_qpos.reshape(2, m.nq)
    np.testing.assert_array_equal(kq[0], m.qpos0)                     # unspecified qpos -> reference pose
    np.testing.assert_array_equal(m.key_mpos.reshape(2, 3)[0], [0.5, 0, 1])   # unspecified mocap pose -> body pose
    np.testing.assert_array_equal(m.key_mquat.reshape(2, 4)[0], [1, 0, 0, 0])
    assert kq[1][7] == 0.3 and m.key_time[1] == 1.5
    np.testing.assert_array_equal(m.key_qvel.reshape(2, m.nv)[1], [0, 0, 0, 0, 0, 0, 2.0])
    np.testing.assert_array_equal(m.key_act, [0, 0.25])
    np.testing.assert_array_equal(m.key_ctrl.reshape(2, 2), [[0, 0], [0.5, -0.5]])
    np.testing.assert_array_equal(m.key_mpos.reshape(2, 3)[1], [1, 1, 1])


def test_keyframe_with_wrong_length_is_rejected(capi):
    bad = KEYED.replace('act="0.25"', 'act="0.25 1"')
    with pytest.raises(capi.B2mjError, match="needs 1 numbers"):
        capi.Model.from_xml_string(bad)


def test_oracle_reset_keyframe(capi, orc):
    m = capi.Model.from_xml_string(KEYED)
    o = orc.Oracle(m)
    o.set("qvel", np.ones(m.nv))
    o.step(5)
    o.reset_keyframe(1)
    assert o.time == 1.5
    np.testing.assert_array_equal(o.get("qpos"), m.key_qpos.reshape(2, -1)[1])
    np.testing.assert_array_equal(o.get("qvel"), m.key_qvel.reshape(2, -1)[1])
    np.testing.assert_array_equal(o.get("act"), [0.25])
    np.testing.assert_array_equal(o.get("ctrl"), [0.5, -0.5])
    np.testing.assert_array_equal(o.get("mocap_pos"), [1, 1, 1])
    np.testing.assert_array_equal(o.get("mocap_quat"), [0, 1, 0, 0])
    assert np.all(o.get("qacc_warmstart") == 0) and np.all(o.get("warning") == 0)
    o.reset_keyframe(0)
    assert o.time == 0 and np.array_equal(o.get("qpos"), m.qpos0)


@pytest.mark.parametrize("name", ["panda_like.xml", "humanoid_like.xml", "bin.xml", "equality_scene.xml"])
def test_binary_round_trip_is_bit_exact(capi, tmp_path, name):
    a = capi.Model.from_xml_file(model_path(name))
    path = str(tmp_path / (name.replace(".xml", "") + ".b2mjb"))
    a.save_binary(path)
    b = capi.Model.from_file(path)            # extension dispatch
    assert a._sizes == b._sizes
    for k, v in a._arrays.items():
        np.testing.assert_array_equal(v, b._arrays[k], err_msg=k)
    assert bytes(a.opt) == bytes(b.opt) and bytes(a.stat) == bytes(b.stat)
    assert b.name2id(capi.OBJ_BODY, a.id2name(capi.OBJ_BODY, a.nbody - 1)) == a.nbody - 1


def test_binary_round_trip_keeps_keyframes_and_steps_identically(capi, orc, tmp_path):
    a = capi.Model.from_xml_string(KEYED)
    path = str(tmp_path / "keyed.b2mjb")
    a.save_binary(path)
    b = capi.Model.load_binary(path)
    assert b.nkey == 2 and np.array_equal(a.key_qpos, b.key_qpos)
    oa, ob = orc.Oracle(a), orc.Oracle(b)
    for o in (oa, ob):
        o.reset_keyframe(1)
        o.step(50)
    np.testing.assert_array_equal(oa.get("qpos"), ob.get("qpos"))


def test_binary_loader_rejects_garbage(capi, tmp_path):
    p = tmp_path / "junk.b2mjb"
    p.write_bytes(b"not a model at all")
    with pytest.raises(capi.B2mjError, match="not a b2mj binary model"):
        capi.Model.from_file(str(p))
    a = capi.Model.from_xml_file(model_path("panda_like.xml"))
    good = tmp_path / "ok.b2mjb"
    a.save_binary(str(good))
    cut = tmp_path / "cut.b2mjb"
    cut.write_bytes(good.read_bytes()[:2000])
    with pytest.raises(capi.B2mjError):
        capi.Model.from_file(str(cut))


def test_from_file_dispatches_xml(capi):
    m = capi.Model.from_file(model_path("pendulum_scene.xml"))
    assert m.nq > 0


@pytest.mark.gpu
def test_gpu_reset_keyframe_matches_oracle(capi, orc):
    from mujoco_ros_pkgs_b200.batch import BatchSim

    m = capi.Model.from_xml_string(KEYED)
    nenv = 6
    sim = BatchSim(m, nenv)
    sim.set("qvel", np.ones((nenv, m.nv)))
    sim.step(7)
    mask = np.array([1, 0, 1, 1, 0, 1], dtype=np.uint8)
    before = {k: sim.get(k) for k in ("qpos", "qvel", "time")}
    sim.reset_keyframe(1, mask)
    o = orc.Oracle(m)
    o.reset_keyframe(1)
    for k in ("qpos", "qvel", "act", "ctrl", "time", "mocap_pos", "mocap_quat", "qacc_warmstart"):
        g = sim.get(k)
        for e in range(nenv):
            if mask[e]:
                np.testing.assert_array_equal(g[e], o.get(k), err_msg=f"{k} env {e}")
            elif k in before:
                np.testing.assert_array_equal(g[e], before[k][e], err_msg=f"{k} env {e} (unmasked)")
    # stepping from the keyframe follows the oracle
    sim.reset_keyframe(1)
    sim.step(100)
    o.step(100)
    gq = sim.get("qpos")
    assert np.max(np.abs(gq[0] - o.get("qpos")) / (1 + np.abs(o.get("qpos")))) < 1e-8
    with pytest.raises(capi.B2mjError):
        sim.reset_keyframe(2)


def test_include_files_are_expanded_in_place(capi, tmp_path):
    """<include file=.../> anywhere in the tree (MJCF's way of splitting robot / scene files); the reference loads such
    models through mj_loadXML (mujoco_env.cpp:840-843)."""
    (tmp_path / "arm.xml").write_text("""<mujocoinclude>
      <body name="upper" pos="0 0 1"><joint name="sh" type="hinge" axis="0 1 0"/>
        <geom type="capsule" fromto="0 0 0 0.3 0 0" size="0.03"/>
        <include file="fore.xml"/>
      </body></mujocoinclude>""")
    (tmp_path / "fore.xml").write_text("""<mujoco><body name="fore" pos="0.3 0 0"><joint name="el" type="hinge" axis="0 1 0"/>
      <geom type="capsule" fromto="0 0 0 0.25 0 0" size="0.025"/></body></mujoco>""")
    (tmp_path / "act.xml").write_text('<mujoco><actuator><motor joint="sh"/><motor joint="el"/></actuator></mujoco>')
    main = tmp_path / "scene.xml"
    main.write_text("""<mujoco><option timestep="0.002"/><worldbody><geom type="plane" size="2 2 0.1"/>
      <include file="arm.xml"/></worldbody><include file="act.xml"/></mujoco>""")
    m = capi.Model.from_xml_file(str(main))
    assert (m.nbody, m.njnt, m.nu, m.ngeom) == (3, 2, 2, 3)
    assert m.name2id(capi.OBJ_BODY, "fore") == 2 and m.body_parentid[2] == 1
    with pytest.raises(capi.B2mjError, match="cannot open included file"):
        capi.Model.from_xml_string('<mujoco><include file="nope.xml"/></mujoco>')
    (tmp_path / "loop.xml").write_text('<mujoco><include file="loop.xml"/></mujoco>')
    with pytest.raises(capi.B2mjError, match="nested too deeply"):
        capi.Model.from_xml_file(str(tmp_path / "loop.xml"))


def test_composite_is_rejected_not_ignored(capi):
    with pytest.raises(capi.B2mjError, match="composite"):
        capi.Model.from_xml_string('<mujoco><worldbody><body><composite type="grid" count="2 2 1"/></body></worldbody></mujoco>')


PAIRS = """
<mujoco>
  <option timestep="0.002" collision="{mode}"/>
  <worldbody>
    <geom name="floor" type="plane" size="2 2 0.1" friction="0.3 0.005 0.0001"/>
    <body name="a" pos="0 0 0.1"><freejoint/><geom name="ga" type="sphere" size="0.1" friction="0.5 0.005 0.0001" contype="0" conaffinity="0"/></body>
    <body name="b" pos="0.5 0 0.1"><freejoint/><geom name="gb" type="sphere" size="0.1" friction="0.5 0.005 0.0001"/></body>
    <body name="c" pos="0.5 0 0.29"><freejoint/><geom name="gc" type="sphere" size="0.1"/></body>
  </worldbody>
  <contact>
    <pair name="grip" geom1="ga" geom2="floor" condim="4" friction="2 2 0.1 0.01 0.01" solref="0.01 1" margin="0.02" gap="0.005"/>
    <pair geom1="gc" geom2="gb"/>
    <exclude body1="b" body2="c"/>
  </contact>
</mujoco>
"""


def test_explicit_contact_pairs(capi, orc):
    """<contact><pair>: own parameters, no contype / conaffinity / exclude filtering, merged with the dynamic pairs in
    body-pair order; option collision = all / predefined / dynamic selects the sources."""
    m = capi.Model.from_xml_string(PAIRS.format(mode="all"))
    gid = lambda n: m.name2id(capi.OBJ_GEOM, n)
    assert m.npair == 2 and m.name2id(capi.OBJ_PAIR, "grip") == 0
    cand = list(zip(m.collpair_geom1.tolist(), m.collpair_geom2.tolist(), m.collpair_pairid.tolist()))
    # floor-ga only through its pair (ga has contype 0), floor-gb and floor-gc dynamic, gb-gc through its pair although
    # the bodies are excluded; ga-gb / ga-gc filtered out by contype
    assert (gid("floor"), gid("ga"), 0) in cand and (gid("floor"), gid("gb"), -1) in cand and (gid("floor"), gid("gc"), -1) in cand
    assert (gid("gb"), gid("gc"), 1) in cand or (gid("gc"), gid("gb"), 1) in cand
    assert len(cand) == 4
    np.testing.assert_allclose(m.pair_friction[1], [1, 1, 0.005, 0.0001, 0.0001])   # filled from the geoms (max rule)
    o = orc.Oracle(m)
    o.forward()
    n = int(o.get("ncon")[0])
    g1, g2 = o.get("contact_geom1")[:n].tolist(), o.get("contact_geom2")[:n].tolist()
    assert n == 3 and (gid("floor"), gid("ga")) in zip(g1, g2) and (gid("gc"), gid("gb")) in zip(g1, g2)  # pair order kept
    k = list(zip(g1, g2)).index((gid("floor"), gid("ga")))
    assert o.get("contact_dim")[k] == 4
    np.testing.assert_allclose(o.get("contact_friction").reshape(-1, 5)[k], [2, 2, 0.1, 0.01, 0.01])
    np.testing.assert_allclose(o.get("contact_solref").reshape(-1, 2)[k], [0.01, 1])
    assert abs(o.get("contact_includemargin")[k] - 0.015) < 1e-15
    k2 = list(zip(g1, g2)).index((gid("gc"), gid("gb")))
    assert abs(o.get("contact_dist")[k2] + 0.01) < 1e-12 and o.get("contact_dim")[k2] == 3
    for mode, want in (("predefined", 2), ("dynamic", 2)):
        mm = capi.Model.from_xml_string(PAIRS.format(mode=mode))
        assert mm.ncollpair == want, (mode, mm.ncollpair)
        assert all((p >= 0) == (mode == "predefined") for p in mm.collpair_pairid)


@pytest.mark.gpu
def test_gpu_explicit_contact_pairs(capi, orc):
    from mujoco_ros_pkgs_b200.batch import BatchSim

    m = capi.Model.from_xml_string(PAIRS.format(mode="all"))
    nenv = 4
    sim = BatchSim(m, nenv)
    sim.keep_intermediates(True)
    rng = np.random.default_rng(0)
    qpos = np.tile(m.qpos0, (nenv, 1))
    qpos[:, [0, 7, 14]] += rng.uniform(-0.02, 0.02, (nenv, 3))
    qvel = rng.uniform(-0.5, 0.5, (nenv, m.nv))
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = []
    for e in range(nenv):
        o = orc.Oracle(m)
        o.set("qpos", qpos[e]); o.set("qvel", qvel[e])
        oracles.append(o)
    for s in range(120):
        sim.step(1)
        for o in oracles:
            o.step(1)
        if s % 20 == 19:
            sim.forward()
            for e, o in enumerate(oracles):
                o.forward()
                n = int(o.get("ncon")[0])
                assert int(sim.get("ncon")[e, 0]) == n
                for k in ("contact_geom1", "contact_geom2", "contact_dim"):
                    np.testing.assert_array_equal(sim.get(k)[e][:n], o.get(k)[:n])
                np.testing.assert_allclose(sim.get("contact_friction")[e][:5 * n], o.get("contact_friction")[:5 * n], rtol=0, atol=0)
                np.testing.assert_allclose(sim.get("qpos")[e], o.get("qpos"), rtol=0, atol=1e-9)
