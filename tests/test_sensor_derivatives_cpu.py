"""Row M10 held to definitions instead of to a restated formula: velocity sensors are time derivatives of position
sensors, acceleration sensors are time derivatives of velocity sensors (plus the gravity offset MuJoCo's accelerations
carry), body-frame sensors are the world-frame ones rotated into the site frame.  Central differences over the
oracle's own forward pass; no sensor code is involved in computing the expected values."""
import numpy as np
import pytest

from test_solver_optimality_cpu import integrate_pos

XML = """<mujoco><option timestep="0.002" gravity="0 0 -9.81"/><worldbody>
  <body name="root" pos="0 0 1"><freejoint/><geom type="box" size="0.1 0.06 0.04"/>
    <site name="s0" pos="0.05 0.02 0.03" euler="0.3 -0.2 0.5"/>
    <body name="arm" pos="0.1 0 0"><joint name="h" axis="0 1 0" pos="0 0 0.01"/><geom type="capsule" fromto="0 0 0 0.25 0 0" size="0.03"/>
      <body name="tip" pos="0.25 0 0"><joint name="b" type="ball"/><geom type="ellipsoid" size="0.06 0.04 0.03" pos="0.05 0 0"/>
        <site name="s1" pos="0.08 0.01 -0.02" euler="-0.4 0.1 0.7"/>
        <body name="rod" pos="0.1 0 0"><joint name="sl" type="slide" axis="0.6 0 0.8"/><geom size="0.03"/><site name="s2" pos="0 0.02 0"/></body>
      </body></body></body>
</worldbody><sensor>
  %s
  <subtreecom name="com" body="arm"/><subtreelinvel name="comvel" body="arm"/>
  <subtreecom name="rodcom" body="rod"/><force name="frc" site="s2"/>
  <tendonpos name="tp" tendon="t"/><tendonvel name="tv" tendon="t"/><actuatorpos name="ap" actuator="act"/>
  <actuatorvel name="av" actuator="act"/><actuatorpos name="ap2" actuator="act2"/><actuatorvel name="av2" actuator="act2"/>
  <ballquat name="bq" joint="b"/><ballangvel name="bw" joint="b"/><jointpos name="jp" joint="sl"/><jointvel name="jv" joint="sl"/>
</sensor>
<tendon><spatial name="t"><site site="s0"/><site site="s1"/><pulley divisor="2"/><site site="s1"/><site site="s2"/></spatial></tendon>
<actuator><general name="act" tendon="t" gear="3"/><general name="act2" joint="h" gear="-2"/></actuator>
</mujoco>"""
PER_SITE = ('<framepos name="p{0}" objtype="site" objname="{0}"/><framequat name="q{0}" objtype="site" objname="{0}"/>'
            '<framelinvel name="v{0}" objtype="site" objname="{0}"/><frameangvel name="w{0}" objtype="site" objname="{0}"/>'
            '<framelinacc name="a{0}" objtype="site" objname="{0}"/><frameangacc name="al{0}" objtype="site" objname="{0}"/>'
            '<velocimeter name="vm{0}" site="{0}"/><gyro name="gy{0}" site="{0}"/><accelerometer name="ac{0}" site="{0}"/>'
            '<framexaxis name="x{0}" objtype="site" objname="{0}"/><frameyaxis name="y{0}" objtype="site" objname="{0}"/>'
            '<framezaxis name="z{0}" objtype="site" objname="{0}"/>')
SITES = ("s0", "s1", "s2")


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_velocity_and_acceleration_sensors_are_time_derivatives(seed, capi, orc):
    m = capi.Model.from_xml_string(XML % "".join(PER_SITE.format(s) for s in SITES))
    o = orc.Oracle(m)
    rng = np.random.default_rng(seed)
    q0 = m.qpos0.copy()
    q0[3:7] = rng.normal(size=4)
    q0[3:7] /= np.linalg.norm(q0[3:7])
    q0[8:12] = rng.normal(size=4)
    q0[8:12] /= np.linalg.norm(q0[8:12])
    q0[7], q0[12] = rng.uniform(-1, 1), rng.uniform(-0.05, 0.05)
    v0 = rng.uniform(-2, 2, m.nv)

    def read(q, v):
        o.set("qpos", q)
        o.set("qvel", v)
        o.forward()
        sd = o.get("sensordata").copy()
        out = {}
        for i in range(m.nsensor):
            out[m.id2name(capi.OBJ_SENSOR, i)] = sd[m.sensor_adr[i]:m.sensor_adr[i] + m.sensor_dim[i]]
        return out, o.get("qacc").copy()
    now, qacc = read(q0, v0)
    h = 1e-5
    # the state a moment later / earlier along the true motion: q advanced with v (+ h^2/2 qacc), v with qacc
    fwd, _ = read(integrate_pos(m, integrate_pos(m, q0, v0, h), qacc, 0.5 * h * h), v0 + h * qacc)
    bwd, _ = read(integrate_pos(m, integrate_pos(m, q0, v0, -h), qacc, 0.5 * h * h), v0 - h * qacc)
    ddt = lambda k: (fwd[k] - bwd[k]) / (2 * h)  # noqa: E731
    g = np.array([0, 0, 9.81])
    for s in SITES:
        R = np.stack([now["x" + s], now["y" + s], now["z" + s]], axis=1)  # site axes as columns
        np.testing.assert_allclose(now["v" + s], ddt("p" + s), rtol=1e-6, atol=1e-7, err_msg=f"framelinvel {s}")
        w_fd = 2 * qmul(ddt("q" + s), now["q" + s] * [1, -1, -1, -1])[1:]  # world-frame angular velocity from dq/dt
        np.testing.assert_allclose(now["w" + s], w_fd, rtol=1e-6, atol=1e-7, err_msg=f"frameangvel {s}")
        np.testing.assert_allclose(now["vm" + s], R.T @ now["v" + s], rtol=1e-10, atol=1e-12, err_msg=f"velocimeter {s}")
        np.testing.assert_allclose(now["gy" + s], R.T @ now["w" + s], rtol=1e-10, atol=1e-12, err_msg=f"gyro {s}")
        # MuJoCo's accelerations are relative to a world that accelerates at -gravity
        np.testing.assert_allclose(now["a" + s], ddt("v" + s) + g, rtol=2e-5, atol=2e-5, err_msg=f"framelinacc {s}")
        np.testing.assert_allclose(now["al" + s], ddt("w" + s), rtol=2e-5, atol=2e-5, err_msg=f"frameangacc {s}")
        np.testing.assert_allclose(now["ac" + s], R.T @ now["a" + s], rtol=1e-10, atol=1e-11, err_msg=f"accelerometer {s}")
    np.testing.assert_allclose(now["comvel"], ddt("com"), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(now["jv"], ddt("jp"), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(now["tv"], ddt("tp"), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(now["av"], ddt("ap"), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(now["av2"], ddt("ap2"), rtol=1e-6, atol=1e-8)
    assert abs(now["av"][0]) > 1e-3 and abs(now["av"][0] - 3 * now["tv"][0]) < 1e-12
    # force sensor on a leaf body with nothing else acting on it: Newton's second law for the body, in the site frame
    # (MuJoCo's sign: the force the parent exerts on the child; accelerations carry the gravity offset)
    a_com = (fwd["rodcom"] - 2 * now["rodcom"] + bwd["rodcom"]) / (h * h)
    mass = m.body_mass[m.name2id(capi.OBJ_BODY, "rod")]
    R2 = np.stack([now["xs2"], now["ys2"], now["zs2"]], axis=1)
    np.testing.assert_allclose(now["frc"], R2.T @ (mass * (a_com + g)), rtol=1e-3, atol=2e-3 * mass * 9.81)
    bw_fd = 2 * qmul(now["bq"] * [1, -1, -1, -1], ddt("bq"))[1:]  # ball joint: angular velocity in the child frame
    np.testing.assert_allclose(now["bw"], bw_fd, rtol=1e-6, atol=1e-7)


REST = """<mujoco><compiler angle="radian"/><option timestep="0.002"/><worldbody>
  <geom type="plane" size="2 2 .1"/>
  <body name="ball" pos="0 0 0.1"><freejoint/><geom size="0.1" mass="1.5"/><site name="pad" type="sphere" size="0.12"/>
    <site name="mag" euler="0.4 0.3 -0.6"/></body>
  <body name="arm" pos="1 0 1"><joint name="h" axis="0 1 0" range="-0.5 0.5" limited="true" damping="0.5"/>
    <geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.02" mass="0.8"/></body>
  <body name="lift" pos="2 0 1"><joint name="s" type="slide" axis="0 0 1" damping="5"/><geom size="0.05" mass="2"/></body>
</worldbody>
<actuator><motor name="m" joint="s" gear="4" ctrlrange="-10 10"/></actuator>
<sensor><touch name="touch" site="pad"/><magnetometer name="mag" site="mag"/><clock name="clock"/>
  <jointlimitpos name="lp" joint="h"/><jointlimitvel name="lv" joint="h"/><jointlimitfrc name="lf" joint="h"/>
  <actuatorfrc name="af" actuator="m"/><jointactuatorfrc name="jaf" joint="s"/><jointpos name="hp" joint="h"/>
  <framezaxis name="mz" objtype="site" objname="mag"/><framexaxis name="mx" objtype="site" objname="mag"/>
  <frameyaxis name="my" objtype="site" objname="mag"/></sensor></mujoco>"""


@pytest.mark.parametrize("cone", ["pyramidal", "elliptic"])
def test_force_like_sensors_at_rest_equal_the_static_loads(cone, capi, orc):
    """Scene at rest: the touch sensor reads the ball's weight, the joint-limit force the gravity torque of the arm
    resting on its limit, the actuator force sensor ctrl and the joint one gear x ctrl, the magnetometer the model's magnetic vector in the
    site frame, the clock the time."""
    m = capi.Model.from_xml_string(REST.replace('timestep="0.002"', f'timestep="0.002" cone="{cone}"'))
    o = orc.Oracle(m)
    o.set("ctrl", [2 * 9.81 / 4])  # holds the lift against gravity
    o.step(3000)
    sd = o.get("sensordata")
    val = {m.id2name(capi.OBJ_SENSOR, i): sd[m.sensor_adr[i]:m.sensor_adr[i] + m.sensor_dim[i]] for i in range(m.nsensor)}
    assert abs(val["touch"][0] - 1.5 * 9.81) < 1e-6 * 1.5 * 9.81
    theta = val["hp"][0]
    assert 0.5 < theta < 0.51  # resting just past the upper limit (positive rotation about y lowers the arm)
    np.testing.assert_allclose(val["lf"][0], 0.8 * 9.81 * 0.2 * np.cos(theta), rtol=1e-6)
    np.testing.assert_allclose(val["lp"][0], 0.5 - theta, rtol=1e-9)  # distance to the limit, negative when violated
    assert abs(val["lv"][0]) < 1e-8
    np.testing.assert_allclose(val["af"], [2 * 9.81 / 4], rtol=1e-12)  # the scalar actuator force; the gear sits in the moment arm
    np.testing.assert_allclose(val["jaf"], [2 * 9.81], rtol=1e-12)
    R = np.stack([val["mx"], val["my"], val["mz"]], axis=1)
    np.testing.assert_allclose(val["mag"], R.T @ np.asarray(m.opt.magnetic), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(val["clock"], [o.get("time")[0] - 0.002], rtol=1e-12)  # sensors are computed before the time advances
