"""The option panel's disable flags (mjtDisableBit; the reference exposes every one of them, viewer.cpp "Disable Flags"
section of the physics panel, and mj_loadXML reads <option><flag .../>).  Each flag is checked on the oracle for what it
is defined to switch off, on one scene that has something for every flag to remove.  The GPU parity half is at the end
of the file, and the file sorts last in the suite on purpose: that half has not run on hardware yet (the round's GPU
budget was spent), so nothing is queued behind it."""
import numpy as np
import pytest

SCENE = """<mujoco><compiler angle="radian"/><option timestep="0.002" solver="{solver}"><flag {flag}="disable"/></option>
<worldbody>
  <geom type="plane" size="2 2 .1"/>
  <body name="ball" pos="0 0 0.095"><freejoint/><geom size="0.1" mass="1.5"/></body>
  <body name="arm" pos="1 0 1"><joint name="h" axis="0 1 0" range="-0.5 0.02" limited="true" damping="0.5" stiffness="3" springref="0.4"
      frictionloss="0.2" solreflimit="0.001 1"/>
    <geom name="upper" type="capsule" fromto="0 0 0 0.4 0 0" size="0.05" mass="0.8"/>
    <body pos="0.4 0 0"><joint name="k" axis="0 1 0" damping="0.1"/><geom name="lower" type="capsule" fromto="0 0 0 0.3 0 0" size="0.05" mass="0.5"/>
      <site name="tip" pos="0.3 0 0"/></body></body>
  <body name="anchor" pos="1.7 0 1.1"><joint type="slide" axis="0 0 1" damping="2"/><geom size="0.03" mass="0.3"/></body>
</worldbody>
<equality><connect body1="anchor" body2="arm" anchor="0 0 -0.1"/></equality>
<actuator><motor name="m" joint="k" gear="2" ctrlrange="-1 1" ctrllimited="true"/></actuator>
<sensor><jointpos joint="h"/><framepos objtype="site" objname="tip"/><accelerometer site="tip"/></sensor>
</mujoco>"""
FLAGS = ("constraint", "equality", "frictionloss", "limit", "contact", "passive", "gravity", "clampctrl", "warmstart",
         "filterparent", "actuation", "refsafe", "sensor", "midphase", "eulerdamp")
EQUALITY, FRICTION_DOF, LIMIT_JOINT, CONTACT = 0, 1, 3, (5, 6, 7)


def build(capi, flag, solver="Newton"):
    xml = SCENE.format(flag=flag, solver=solver)
    if flag is None:
        xml = xml.replace('<flag None="disable"/>', "")
    return capi.Model.from_xml_string(xml)


def forward(orc, m, ctrl=0.6, hinge=0.05):
    o = orc.Oracle(m)
    q = m.qpos0.copy()
    q[7] = hinge  # past the upper limit of 0.02
    o.set("qpos", q)
    o.set("qvel", np.full(m.nv, 0.3))
    o.set("ctrl", [ctrl])
    o.forward()
    return o


def row_types(o):
    return set(int(t) for t in o.get("efc_type")[:int(o.get("nefc")[0])])


def test_every_flag_switches_off_what_it_names(capi, orc):
    base = forward(orc, build(capi, None))
    assert row_types(base) >= {EQUALITY, FRICTION_DOF, LIMIT_JOINT} and row_types(base) & set(CONTACT)
    assert int(base.get("ncon")[0]) == 1 and np.abs(base.get("qfrc_passive")).max() > 0.1

    def flagged(f, **kw):
        m = build(capi, f)
        assert m.opt.disableflags == 1 << FLAGS.index(f), f
        return forward(orc, m, **kw), m
    o, _ = flagged("constraint")
    assert int(o.get("nefc")[0]) == 0 and int(o.get("ncon")[0]) == 0
    np.testing.assert_array_equal(o.get("qacc"), o.get("qacc_smooth"))
    for f, gone in (("equality", {EQUALITY}), ("frictionloss", {FRICTION_DOF}), ("limit", {LIMIT_JOINT}), ("contact", set(CONTACT))):
        o, _ = flagged(f)
        assert not (row_types(o) & gone), f
        assert row_types(o) == row_types(base) - gone, f
    o, _ = flagged("contact")
    assert int(o.get("ncon")[0]) == 0
    o, _ = flagged("passive")
    assert not o.get("qfrc_passive").any()
    o, m = flagged("gravity")
    o.set("qvel", np.zeros(m.nv))
    o.forward()
    assert np.abs(o.get("qfrc_bias")).max() < 1e-14 and np.abs(base.get("qfrc_bias")).max() > 1
    # ctrl beyond its range: clamped to 1 by default, taken as given with the flag
    assert base.get("actuator_force")[0] == 0.6 and forward(orc, build(capi, None), ctrl=3).get("actuator_force")[0] == 1.0
    o, _ = flagged("clampctrl", ctrl=3)
    assert o.get("actuator_force")[0] == 3.0
    o, _ = flagged("actuation")
    assert not o.get("qfrc_actuator").any() and base.get("qfrc_actuator")[7] == 1.2
    # filterparent: the two arm capsules (parent and child, overlapping at the elbow) collide only without the filter
    o, m = flagged("filterparent")
    assert m.ncollpair == build(capi, None).ncollpair + 1 and int(o.get("ncon")[0]) > int(base.get("ncon")[0])
    # refsafe: the limit's 1 ms time constant is raised to two time steps unless the flag is set
    def limit_k(o):
        i = list(o.get("efc_type")[:int(o.get("nefc")[0])]).index(LIMIT_JOINT)
        return o.get("efc_KBIP")[4 * i]
    dmax = 0.95
    np.testing.assert_allclose(limit_k(base), 1 / (dmax * 0.004) ** 2, rtol=1e-12)
    np.testing.assert_allclose(limit_k(flagged("refsafe")[0]), 1 / (dmax * 0.001) ** 2, rtol=1e-12)
    o, _ = flagged("sensor")
    assert not o.get("sensordata").any() and np.abs(base.get("sensordata")).max() > 0.1
    o, _ = flagged("midphase")  # broad-phase choice only: same answer
    np.testing.assert_array_equal(o.get("qacc"), base.get("qacc"))


def test_warmstart_and_eulerdamp_flags(capi, orc):
    # warmstart: with the flag the solver starts from qacc_smooth whatever qacc_warmstart holds; by default it starts
    # from qacc_warmstart when that has the lower cost (the converged answer does, zeros do not)
    conv = build(capi, None, solver="PGS")
    conv.opt.iterations, conv.opt.tolerance = 500, 1e-14
    best = forward(orc, conv).get("qacc").copy()
    for flag, same in (("warmstart", True), (None, False)):
        m = build(capi, flag, solver="PGS")
        m.opt.iterations = 1  # an unconverged answer shows where the iteration started
        outs = []
        for ws in (np.zeros(m.nv), best):
            o = forward(orc, m)
            o.set("qacc_warmstart", ws)
            o.forward()
            outs.append(o.get("efc_force")[:int(o.get("nefc")[0])].copy())
        assert np.array_equal(outs[0], outs[1]) == same, flag
    # eulerdamp: with the flag Euler is fully explicit, qvel' = qvel + h qacc; by default joint damping is implicit
    for flag, explicit in (("eulerdamp", True), (None, False)):
        m = build(capi, flag)
        o = forward(orc, m)
        v0 = o.get("qvel").copy()
        o.step(1)
        # qacc after the step is the one the step integrated (forward ran inside it on the same state)
        exact = np.allclose(o.get("qvel"), v0 + m.opt.timestep * o.get("qacc"), rtol=0, atol=1e-15)
        assert exact == explicit, flag


@pytest.mark.gpu
@pytest.mark.timeout(60, method="thread")
@pytest.mark.xfail(strict=False, reason="written after the round's GPU budget was spent: not yet run on hardware; an XPASS in "
                   "the driver's log is the first confirmation, an xfail a device-side difference to chase")
@pytest.mark.parametrize("flag", [f for f in FLAGS if f != "contact"])
def test_flag_gpu_parity(flag, capi, orc):
    from mujoco_ros_pkgs_b200.batch import BatchSim
    from parity_util import compare_forward_fields, injected_steps, make_oracles, perturbed

    model = build(capi, flag, solver="PGS" if flag == "warmstart" else "Newton")
    nenv = 8
    rng = np.random.default_rng(3)
    qpos, qvel = perturbed(model, nenv, seed=2, amp=0.05)
    qpos[:, 7] = rng.uniform(-0.1, 0.06, nenv)  # some envs past the hinge limit
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)
    injected_steps(model, sim, oracles, 30, rng, tag=f"flag {flag}")
    sim.keep_intermediates(True)
    sim.forward()
    st = {k: sim.get(k) for k in ("qpos", "qvel", "ctrl", "qacc_warmstart")}
    for e, o in enumerate(oracles):
        for k, v in st.items():
            o.set(k, v[e])
        o.forward()
    compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tag=f"flag {flag}")


@pytest.mark.gpu
@pytest.mark.timeout(60, method="thread")
@pytest.mark.xfail(strict=False, reason="written after the round's GPU budget was spent: not yet run on hardware")
@pytest.mark.parametrize("cone,solver", [(0, 2), (1, 2), (1, 0), (0, 1)])
def test_contact_dimensions_margin_gap_mixing_gpu_parity(cone, solver, capi, orc):
    """condim 1 / 4 / 6 rows (torsional, rolling), margin + gap, solmix / priority: no bench model has them."""
    from mujoco_ros_pkgs_b200.batch import BatchSim
    from parity_util import compare_forward_fields, injected_steps, make_oracles, perturbed
    from test_solver_optimality_cpu import load_case

    model, v0 = load_case("CONDIM", capi)
    model.opt.cone, model.opt.solver = cone, solver
    nenv = 8
    rng = np.random.default_rng(5)
    qpos, _ = perturbed(model, nenv, seed=4, amp=0.0)
    qvel = np.tile(v0, (nenv, 1)) * rng.uniform(0.5, 1.5, (nenv, 1))
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    sim.step(150)
    oracles = make_oracles(orc, model, qpos, qvel)
    _, max_nefc = injected_steps(model, sim, oracles, 30, rng, tag=f"condim cone {cone} solver {solver}")
    assert max_nefc >= 18
    sim.keep_intermediates(True)
    sim.forward()
    st = {k: sim.get(k) for k in ("qpos", "qvel", "qacc_warmstart")}
    for e, o in enumerate(oracles):
        for k, v in st.items():
            o.set(k, v[e])
        o.forward()
    compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tag="condim")
