"""CPU oracle vs facts pinned by the reference's own tests and vs closed-form answers.

PARITY UNPINNED for post-step dynamics: the reference delegates the arithmetic to libmujoco 2.3.7,
which is absent; what its tests DO pin (time arithmetic, qpos0, reset, callback order, static
pendulum) and analytic answers that involve no MuJoCo are checked here.
"""
import numpy as np
import pytest

from conftest import model_path

SINGLE_HINGE = """
<mujoco>
  <compiler angle="radian"/>
  <option timestep="{dt}" integrator="{integ}" gravity="0 0 -9.81">
    <flag contact="disable"/>
  </option>
  <worldbody>
    <body pos="0 0 0">
      <joint name="h" type="hinge" axis="0 1 0"/>
      <inertial pos="0 0 -1" mass="1" diaginertia="1e-9 1e-9 1e-9"/>
    </body>
  </worldbody>
</mujoco>
"""

FREE_BODY = """
<mujoco>
  <option timestep="0.001" gravity="0 0 0"><flag contact="disable"/></option>
  <worldbody>
    <body pos="0 0 1">
      <freejoint/>
      <inertial pos="0 0 0" mass="2" diaginertia="0.1 0.2 0.3"/>
    </body>
  </worldbody>
</mujoco>
"""


def test_time_after_steps(orc, load_model):
    # mujoco_env_test.cpp:198-200 (DOUBLE_EQ after 1 step), :219-221 (100 steps within 1e-6)
    m = load_model("pendulum_scene.xml")
    o = orc.Oracle(m)
    o.step(1)
    assert o.time == m.opt.timestep
    o.step(99)
    assert abs(o.time - 100 * m.opt.timestep) < 1e-6


def test_reset_restores_state(orc, load_model):
    # mujoco_env_test.cpp:507,519-526
    m = load_model("pendulum_scene.xml")
    o = orc.Oracle(m)
    o.set("qpos", m.qpos0 + 0.01)
    o.set("qvel", np.full(m.nv, 0.3))
    o.step(10)
    o.reset()
    assert o.time == 0
    np.testing.assert_array_equal(o.get("qpos"), m.qpos0)
    np.testing.assert_array_equal(o.get("qvel"), np.zeros(m.nv))


def test_free_fall_closed_form(orc, load_model):
    # v_k = -g h k ; z_k = z0 - g h^2 k(k+1)/2 ; first contact in the pass from t=0.045
    m = load_model("pendulum_scene.xml")
    o = orc.Oracle(m)
    g, h = 9.81, 1e-3
    for k in range(1, 45):
        o.step(1)
        assert o.get("ncon")[0] == 0
        assert o.get("qvel")[7] == pytest.approx(-g * h * k, rel=1e-12)
        assert o.get("qpos")[8] == pytest.approx(0.06 - g * h * h * k * (k + 1) / 2, rel=1e-12)
    o.step(1)  # forward pass from t=0.044: still no contact
    o.step(1)
    assert o.get("ncon")[0] == 1
    assert o.get("contact_geom1")[0] == 0  # plane first: bit-exact pair indexing
    assert m.id2name(5, int(o.get("contact_geom2")[0])) == "ball"


def test_ball_rest_depth_closed_form(orc, load_model):
    # steady-state penetration 3.6718e-4 m, normal force m g
    m = load_model("pendulum_scene.xml")
    o = orc.Oracle(m)
    o.step(3000)
    z = o.get("qpos")[8]
    assert 0.05 - z == pytest.approx(3.6718e-4, rel=2e-3)
    assert abs(o.get("qvel")[7]) < 1e-6
    assert o.get("efc_force")[0] == pytest.approx(0.1 * 9.81, rel=1e-4)


def test_hanging_pendulum_is_a_bitwise_fixed_point(orc, load_model):
    # mujoco_sensors_test.cpp:389-391,584: GT variance exactly 0 over 1001 steps while the ball lands
    m = load_model("pendulum_scene.xml")
    o = orc.Oracle(m)
    sens0 = None
    for _ in range(1001):
        o.step(1)
        s = o.get("sensordata").copy()
        sens0 = s if sens0 is None else sens0
        np.testing.assert_array_equal(s, sens0)
        np.testing.assert_array_equal(o.get("qvel")[:5], 0.0)
        np.testing.assert_array_equal(o.get("qpos")[:6], [1, 0, 0, 0, 0, 0])
    assert o.get("ncon")[0] == 1  # the free ball did land in the meantime


def test_sensor_values_static_scene(orc, load_model, capi):
    # framepos/framequat of the static box, zero velocities (mujoco_sensors_test.cpp:326-328,440-442)
    m = load_model("pendulum_scene.xml")
    o = orc.Oracle(m)
    o.forward()
    s = o.get("sensordata")
    adr = {m.id2name(capi.OBJ_SENSOR, i): (m.sensor_adr[i], m.sensor_dim[i]) for i in range(m.nsensor)}
    a, d = adr["immovable_pos"]
    np.testing.assert_allclose(s[a:a + d], [0.56428, 0.221972, 0.6], atol=1e-15)
    a, d = adr["immovable_quat"]
    np.testing.assert_allclose(s[a:a + d], [1, 0, 0, 0], atol=1e-15)
    a, d = adr["vel_EE"]
    np.testing.assert_array_equal(s[a:a + d], 0)


def test_callbacks_fire_once_per_step(orc, load_model):
    # mujoco_ros_plugin_test.cpp:97-121; order: passive inside the velocity stage, control before actuation
    m = load_model("pendulum_scene.xml")
    o = orc.Oracle(m)
    order = []
    o.set_callbacks(control=lambda oo: order.append("control"), passive=lambda oo: order.append("passive"))
    o.step(1)
    assert order == ["passive", "control"]
    assert o.callback_counts() == (1, 1)


def test_rk4_fires_control_four_times(orc, capi):
    # plugin_utils.h:118-124: control/passive run for every RK4 sub-step
    m = capi.Model.from_xml_string(SINGLE_HINGE.format(dt=0.01, integ="RK4"))
    o = orc.Oracle(m)
    n = []
    o.set_callbacks(control=lambda oo: n.append(1))
    o.step(1)
    assert len(n) == 4


def test_control_callback_can_write_ctrl(orc, load_model):
    # plugin_utils.h:89-95: ctrl written inside controlCallback takes effect in the same step
    m = load_model("panda_like.xml")
    a, b = orc.Oracle(m), orc.Oracle(m)
    target = np.array([0.3, -0.2, 0.1, -1.5, 0.2, 1.0, 0.1, 100.0])
    a.set_callbacks(control=lambda oo: oo.set("ctrl", target))
    b.set("ctrl", target)
    a.step(5)
    b.step(5)
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))


def test_split_step_equals_step(orc, load_model):
    m = load_model("panda_like.xml")
    a, b = orc.Oracle(m), orc.Oracle(m)
    for o in (a, b):
        o.set("qpos", m.qpos0 + 0.05)
        o.set("ctrl", np.array([0.3, -0.2, 0.1, -1.5, 0.2, 1.0, 0.1, 50.0]))
    for _ in range(20):
        a.step(1)
        b.step1()
        b.step2()
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))
    np.testing.assert_array_equal(a.get("qvel"), b.get("qvel"))


def test_mass_matrix_consistency(orc, load_model):
    # M qacc_smooth == qfrc_smooth ; L'DL == M
    for name in ("panda_like.xml", "pendulum_scene.xml", "equality_scene.xml"):
        m = load_model(name)
        o = orc.Oracle(m)
        rng = np.random.default_rng(3)
        q = m.qpos0 + rng.uniform(-0.2, 0.2, m.nq)
        o.set("qpos", q)
        o.set("qvel", rng.uniform(-1, 1, m.nv))
        o.forward()
        nv = m.nv
        M = np.zeros((nv, nv))
        L = np.eye(nv)
        qM, qLD = o.get("qM"), o.get("qLD")
        D = np.zeros(nv)
        for i in range(nv):
            adr = m.dof_Madr[i]
            D[i] = qLD[adr]
            j, k = i, 0
            while j >= 0:
                M[i, j] = M[j, i] = qM[adr + k]
                if k:
                    L[i, j] = qLD[adr + k]
                j = m.dof_parentid[j]
                k += 1
        np.testing.assert_allclose(L.T @ np.diag(D) @ L, M, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(M @ o.get("qacc_smooth"), o.get("qfrc_smooth"), rtol=1e-9, atol=1e-10)
        assert np.all(np.linalg.eigvalsh(M) > 0)


def hinge_error(orc, capi, integ, dt, T=0.5):
    m = capi.Model.from_xml_string(SINGLE_HINGE.format(dt=dt, integ=integ))
    o = orc.Oracle(m)
    o.set("qpos", [0.3])
    o.step(int(round(T / dt)))
    return o.get("qpos")[0], o.get("qvel")[0]


def test_integrator_order(orc, capi):
    # halving h reduces the RK4 error ~16x, the Euler error ~2x
    ref = hinge_error(orc, capi, "RK4", 1e-4)[0]
    e1 = abs(hinge_error(orc, capi, "RK4", 0.02)[0] - ref)
    e2 = abs(hinge_error(orc, capi, "RK4", 0.01)[0] - ref)
    assert 12 < e1 / e2 < 20
    e1 = abs(hinge_error(orc, capi, "Euler", 0.002)[0] - ref)
    e2 = abs(hinge_error(orc, capi, "Euler", 0.001)[0] - ref)
    assert 1.7 < e1 / e2 < 2.3


def test_pendulum_energy_rk4(orc, capi):
    m = capi.Model.from_xml_string(SINGLE_HINGE.format(dt=0.002, integ="RK4"))
    o = orc.Oracle(m)
    o.set("qpos", [1.0])
    I = 1.0 + 1e-9

    def energy():
        th, w = o.get("qpos")[0], o.get("qvel")[0]
        return 0.5 * I * w * w - 9.81 * np.cos(th)

    e0 = energy()
    o.step(2000)
    assert abs(energy() - e0) < 1e-7


def test_free_body_angular_momentum(orc, capi):
    # torque-free rotation conserves world angular momentum to O(h)
    m = capi.Model.from_xml_string(FREE_BODY)
    o = orc.Oracle(m)
    qv = np.zeros(6)
    qv[3:] = [1.0, 2.0, 0.5]  # body-frame angular velocity
    o.set("qvel", qv)

    def L_world():
        q = o.get("qpos")[3:7]
        w = o.get("qvel")[3:6]
        Lb = np.array([0.1, 0.2, 0.3]) * w
        qw, qx, qy, qz = q
        R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                      [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                      [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
        return R @ Lb

    L0 = L_world()
    o.step(1000)
    assert np.linalg.norm(L_world() - L0) / np.linalg.norm(L0) < 5e-3
    assert abs(np.linalg.norm(o.get("qpos")[3:7]) - 1) < 1e-12


def test_solvers_agree(orc, capi):
    # PGS and Newton reach the same optimum of the convex problem
    xml = open(model_path("panda_like.xml")).read()
    res = {}
    for solver in ("PGS", "Newton"):
        m = capi.Model.from_xml_string(xml.replace('solver="PGS"', f'solver="{solver}"').replace(
            'iterations="100"', 'iterations="1000"').replace('tolerance="1e-8"', 'tolerance="1e-14"'))
        o = orc.Oracle(m)
        q = m.qpos0.copy()
        q[1] = 1.77  # beyond the joint2 upper limit (1.7628) -> active limit rows
        q[3] = -0.05
        o.set("qpos", q)
        o.set("ctrl", np.array([0.0, 1.76, 0, -0.07, 0, 0, 0, 0]))
        o.forward()
        assert o.get("nefc")[0] >= 2
        res[solver] = o.get("qacc").copy()
    np.testing.assert_allclose(res["PGS"], res["Newton"], rtol=1e-5, atol=1e-6)


def test_golden_trajectories(orc, load_model):
    """The oracle reproduces its own frozen trajectories (tests/golden/, tools/make_golden.py)."""
    import os
    from conftest import GOLDEN
    for name in ("panda_like", "pendulum_scene", "equality_scene", "box_stack", "hand_like", "humanoid_like", "bin"):
        path = os.path.join(GOLDEN, f"{name}.npz")
        g = np.load(path)
        m = load_model(f"{name}.xml")
        o = orc.Oracle(m)
        o.set("qpos", g["qpos_init"])
        o.set("qvel", g["qvel_init"])
        for k in range(g["qpos"].shape[0]):
            if m.nu:
                o.set("ctrl", g["ctrl"][k])
            o.step(int(g["stride"]))
            np.testing.assert_allclose(o.get("qpos"), g["qpos"][k], rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(o.get("qvel"), g["qvel"][k], rtol=1e-8, atol=1e-10)


def test_box_stack_rests(orc, load_model):
    """box-box / plane-box face contacts: the slab rests on the table, the cube on the slab."""
    m = load_model("box_stack.xml")
    o = orc.Oracle(m)
    o.step(2500)
    q = o.get("qpos")
    assert q[2] == pytest.approx(0.2 + 0.06, abs=2e-3)
    assert q[9] == pytest.approx(0.2 + 0.12 + 0.05, abs=3e-3)
    assert np.abs(o.get("qvel")).max() < 1e-3
    assert o.get("ncon")[0] == 8


def test_energy_and_momentum_of_a_mixed_joint_tree(orc, capi):
    """Free + hinge + ball + slide chain with off-centre inertias, no dissipation, RK4: total energy 1/2 v'Mv - sum m g.x
    is conserved under gravity (pins the bias forces against the mass matrix), and without gravity so are the linear
    momentum and the angular momentum about the origin of the floating tree (pins the free / ball joint conventions)."""
    from test_sensor_derivatives_cpu import XML
    from test_solver_optimality_cpu import dense_mass

    xml = XML % ""
    xml = (xml[:xml.index("<sensor>")] + '<sensor><subtreecom body="root"/><subtreelinvel body="root"/><subtreeangmom body="root"/>'
           "</sensor></mujoco>")
    for gravity in ("0 0 -9.81", "0 0 0"):
        m = capi.Model.from_xml_string(xml.replace('timestep="0.002" gravity="0 0 -9.81"',
                                                   f'timestep="0.0005" integrator="RK4" gravity="{gravity}"'))
        o = orc.Oracle(m)
        rng = np.random.default_rng(1)
        q = m.qpos0.copy()
        q[3:7] = [0.8, 0.2, -0.4, 0.4]
        q[3:7] /= np.linalg.norm(q[3:7])
        o.set("qpos", q)
        o.set("qvel", rng.uniform(-2, 2, m.nv))
        masses = m.body_mass

        def invariants():
            o.forward()
            v = o.get("qvel")
            x = o.get("xipos").reshape(-1, 3)
            energy = 0.5 * v @ dense_mass(m, o.get("qM")) @ v + 9.81 * (masses @ x[:, 2]) * (gravity != "0 0 0")
            com, sv, sm = o.get("sensordata")[:9].reshape(3, 3)  # of the whole floating tree
            total = masses.sum()
            return energy, total * sv, sm + total * np.cross(com, sv)
        e0, p0, l0 = invariants()
        o.step(2000)  # one second
        e1, p1, l1 = invariants()
        assert abs(e1 - e0) < 2e-8 * max(1.0, abs(e0)), (e0, e1)
        if gravity == "0 0 0":
            np.testing.assert_allclose(p1, p0, rtol=0, atol=1e-7)  # RK4 truncation over 2000 steps, not round-off
            np.testing.assert_allclose(l1, l0, rtol=0, atol=1e-7)
            assert np.linalg.norm(l0) > 1e-3 and np.linalg.norm(p0) > 1e-3
