"""CPU oracle: implicit / implicitfast integrators (oracle/orc_implicit.cpp) against answers that involve no MuJoCo.

The reference exposes the integrator choice (mujoco_ros/src/viewer.cpp:579-582) and reaches it through mj_step
(mujoco_env.cpp:498).  PARITY UNPINNED against libmujoco (absent); what is checked here is self-consistency:
the analytic RNE velocity derivative against central differences, the linear solves against numpy, and the
special cases where implicit, implicitfast and Euler must coincide."""
import numpy as np
import pytest

from conftest import model_path

DAMPED_ARM = """
<mujoco>
  <compiler angle="radian"/>
  <option timestep="0.002" integrator="{integ}" gravity="0 0 -9.81"><flag contact="disable"/></option>
  <worldbody>
    <body pos="0 0 1">
      <joint name="j1" type="hinge" axis="0 1 0" damping="0.7"/>
      <geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.04" density="800"/>
      <body pos="0.4 0 0">
        <joint name="j2" type="hinge" axis="0 0 1" damping="0.3"/>
        <geom type="capsule" fromto="0 0 0 0.3 0.1 0" size="0.03" density="800"/>
        <body pos="0.3 0.1 0">
          <joint name="j3" type="ball" damping="0.05"/>
          <geom type="box" size="0.05 0.08 0.03" pos="0.05 0 0.02" density="900"/>
        </body>
      </body>
    </body>
    <body pos="0 1 1">
      <freejoint/>
      <geom type="box" size="0.1 0.2 0.05" density="500"/>
      <body pos="0.2 0 0">
        <joint name="k1" type="slide" axis="1 0 0" damping="0.2"/>
        <geom type="sphere" size="0.05" pos="0.1 0.05 0" density="700"/>
      </body>
    </body>
  </worldbody>
  {actuators}
</mujoco>
"""
ACT_NONE = ""
ACT_VEL = """<actuator>
    <velocity joint="j1" kv="3.5"/>
    <position joint="j2" kp="20" kv="1.5"/>
    <general joint="k1" gaintype="affine" gainprm="2 0 -0.8" biastype="affine" biasprm="0 -1 -0.4"/>
  </actuator>"""


def _model(capi, integ, actuators=ACT_NONE):
    return capi.Model.from_xml_string(DAMPED_ARM.format(integ=integ, actuators=actuators))


def _random_state(o, seed):
    m = o.model
    rng = np.random.default_rng(seed)
    qpos = m.qpos0.copy() + rng.uniform(-0.3, 0.3, m.nq)
    o.set("qpos", qpos)
    o.set("qvel", rng.uniform(-2, 2, m.nv))
    if m.nu:
        o.set("ctrl", rng.uniform(-1, 1, m.nu))


def test_rne_velocity_derivative_matches_central_differences(capi, orc):
    m = _model(capi, "implicit")
    o = orc.Oracle(m)
    _random_state(o, 1)
    o.forward()
    qvel = o.get("qvel").copy()
    D = orc.rne_vel_derivative(o)
    eps = 1e-6
    fd = np.zeros_like(D)
    for c in range(m.nv):
        for sgn in (+1, -1):
            v = qvel.copy()
            v[c] += sgn * eps
            o.set("qvel", v)
            o.forward()
            fd[:, c] += sgn * o.get("qfrc_bias")
    fd /= 2 * eps
    assert np.max(np.abs(D - fd)) < 1e-7 * (1 + np.max(np.abs(fd)))
    # the exact derivative vanishes outside MuJoCo's ancestor/descendant pattern: nothing is lost by storing only D
    o.set("qvel", qvel)
    o.forward()
    qd = orc.smooth_vel_derivative(o, True)
    off = (qd == 0)
    assert np.max(np.abs(D[off])) < 1e-12


@pytest.mark.parametrize("name", ["humanoid_like.xml", "hand_like.xml", "panda_like.xml"])
def test_rne_velocity_derivative_on_bench_models(load_model, orc, name):
    m = load_model(name)
    o = orc.Oracle(m)
    _random_state(o, 2)
    o.forward()
    qvel = o.get("qvel").copy()
    D = orc.rne_vel_derivative(o)
    rng = np.random.default_rng(3)
    eps = 1e-6
    for _ in range(3):  # directional derivatives
        dv = rng.normal(size=m.nv)
        o.set("qvel", qvel + eps * dv); o.forward(); fp = o.get("qfrc_bias")
        o.set("qvel", qvel - eps * dv); o.forward(); fm = o.get("qfrc_bias")
        fd = (fp - fm) / (2 * eps)
        assert np.max(np.abs(D @ dv - fd)) < 1e-6 * (1 + np.max(np.abs(fd)))


def _dense_M(o):
    m = o.model
    qM = o.get("qM")
    M = np.zeros((m.nv, m.nv))
    for i in range(m.nv):
        adr = m.dof_Madr[i]
        j = i
        while j >= 0:
            M[i, j] = M[j, i] = qM[adr]
            adr += 1
            j = m.dof_parentid[j]
    return M


@pytest.mark.parametrize("integ", ["implicit", "implicitfast"])
def test_implicit_step_solves_the_modified_system(capi, orc, integ):
    """qvel' = qvel + h * inv(M - h qDeriv) (qfrc_smooth + qfrc_constraint), checked with numpy's solver."""
    m = _model(capi, integ, ACT_VEL)
    o = orc.Oracle(m)
    _random_state(o, 4)
    o.forward()
    h = m.opt.timestep
    M = _dense_M(o)
    qD = orc.smooth_vel_derivative(o, integ == "implicit")
    rhs = o.get("qfrc_smooth") + o.get("qfrc_constraint")
    A = M - h * qD
    if integ == "implicitfast":
        A = np.tril(A) + np.tril(A, -1).T  # the L'DL path reads the lower triangle
        assert np.allclose(qD, qD.T, atol=1e-14)
    want = o.get("qvel") + h * np.linalg.solve(A, rhs)
    o.step(1)
    assert np.max(np.abs(o.get("qvel") - want)) < 1e-11 * (1 + np.max(np.abs(want)))


def test_actuator_and_damping_derivatives(capi, orc):
    m = _model(capi, "implicitfast", ACT_VEL)
    o = orc.Oracle(m)
    _random_state(o, 5)
    o.forward()
    qD = orc.smooth_vel_derivative(o, False)
    nv = m.nv
    want = -np.diag(m.dof_damping.copy())
    mom = o.get("actuator_moment").reshape(m.nu, nv)
    ctrl = o.get("ctrl")
    bias_vel = [-3.5, -1.5, -0.4 + (-0.8) * ctrl[2]]
    for a in range(m.nu):
        want = want + bias_vel[a] * np.outer(mom[a], mom[a])
    assert np.max(np.abs(qD - want)) < 1e-13


def test_implicitfast_equals_euler_when_only_joint_damping(capi, orc):
    a, b = orc.Oracle(_model(capi, "implicitfast")), orc.Oracle(_model(capi, "Euler"))
    for o in (a, b):
        _random_state(o, 6)
        o.step(50)
    assert np.max(np.abs(a.get("qpos") - b.get("qpos"))) < 1e-13
    assert np.max(np.abs(a.get("qvel") - b.get("qvel"))) < 1e-13


def test_implicit_equals_implicitfast_without_velocity_dependent_bias(capi, orc):
    xml = """<mujoco><option timestep="0.002" integrator="{integ}"><flag contact="disable"/></option>
      <worldbody><body pos="0 0 1"><joint name="h" type="hinge" axis="0 1 0" damping="0.4"/>
      <geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.05"/></body></worldbody>
      <actuator><velocity joint="h" kv="2"/></actuator></mujoco>"""
    outs = []
    for integ in ("implicit", "implicitfast"):
        o = orc.Oracle(capi.Model.from_xml_string(xml.format(integ=integ)))
        o.set("qvel", [1.3]); o.set("ctrl", [0.5])
        o.step(100)
        outs.append(np.concatenate([o.get("qpos"), o.get("qvel")]))
    assert np.max(np.abs(outs[0] - outs[1])) < 1e-13


def test_velocity_servo_single_step_closed_form(capi, orc):
    """One hinge about the vertical axis (no gravity torque), velocity actuator kv, joint damping b:
    (I + h (kv + b)) qacc = kv (ctrl - v) - b v."""
    xml = """<mujoco><option timestep="0.01" integrator="implicitfast" gravity="0 0 -9.81"><flag contact="disable"/></option>
      <worldbody><body><joint name="h" type="hinge" axis="0 0 1" damping="0.25"/>
      <inertial pos="0 0 0" mass="1" diaginertia="0.1 0.1 0.3"/></body></worldbody>
      <actuator><velocity joint="h" kv="4"/></actuator></mujoco>"""
    m = capi.Model.from_xml_string(xml)
    o = orc.Oracle(m)
    v0, u, h, kv, b, I = 0.8, 2.0, 0.01, 4.0, 0.25, 0.3
    o.set("qvel", [v0]); o.set("ctrl", [u])
    o.step(1)
    qacc = (kv * (u - v0) - b * v0) / (I + h * (kv + b))
    assert abs(o.get("qvel")[0] - (v0 + h * qacc)) < 1e-14
