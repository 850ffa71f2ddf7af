"""CPU oracle: the physics options the reference's option panel exposes beyond solver / integrator
(mujoco_ros/src/viewer.cpp:586-600: density, viscosity, wind, noslip) against closed-form answers."""
import numpy as np
import pytest

CUBE = """
<mujoco>
  <option timestep="0.001" gravity="0 0 -9.81" density="{rho}" viscosity="{mu}" wind="{wind}"><flag contact="disable"/></option>
  <worldbody>
    <body pos="0 0 1">
      <freejoint/>
      <geom type="box" size="0.1 0.1 0.1" density="1000"/>
    </body>
  </worldbody>
</mujoco>
"""
A = 0.2          # cube side = the equivalent inertia box of a solid cube
MASS = 1000 * A ** 3


def _cube(capi, orc, rho=0.0, mu=0.0, wind="0 0 0"):
    m = capi.Model.from_xml_string(CUBE.format(rho=rho, mu=mu, wind=wind))
    return m, orc.Oracle(m)


def test_viscous_drag_of_the_equivalent_sphere(capi, orc):
    mu = 0.8
    m, o = _cube(capi, orc, mu=mu)
    v = np.array([0.3, -0.2, 0.5])
    w = np.array([0.4, 0.1, -0.3])
    o.set("qvel", np.concatenate([v, w]))
    o.forward()
    qp = o.get("qfrc_passive")
    np.testing.assert_allclose(qp[:3], -3 * np.pi * A * mu * v, rtol=1e-12)
    np.testing.assert_allclose(qp[3:], -np.pi * A ** 3 * mu * w, rtol=1e-12)
    # terminal velocity of the falling cube: m g = 3 pi d mu v
    o.set("qvel", np.zeros(6))
    o.step(20000)
    vt = MASS * 9.81 / (3 * np.pi * A * mu)
    assert abs(o.get("qvel")[2] + vt) < 1e-3 * vt or abs(o.get("qvel")[2]) < vt  # still approaching from below
    assert o.get("qvel")[2] < 0


def test_quadratic_drag_and_wind(capi, orc):
    rho = 1.2
    m, o = _cube(capi, orc, rho=rho, wind="2 0 0")
    v = np.array([0.5, -1.0, 0.25])
    w = np.array([3.0, -2.0, 1.0])
    o.set("qvel", np.concatenate([v, w]))
    o.forward()
    qp = o.get("qfrc_passive")
    rel = v - np.array([2.0, 0, 0])  # velocity relative to the wind (identity orientation: local = world)
    np.testing.assert_allclose(qp[:3], -0.5 * rho * A * A * np.abs(rel) * rel, rtol=1e-12)
    np.testing.assert_allclose(qp[3:], -rho * A * (2 * A ** 4) * np.abs(w) * w / 64.0, rtol=1e-12)
    # a body at rest in the wind is pushed downwind
    o.set("qvel", np.zeros(6))
    o.forward()
    assert o.get("qfrc_passive")[0] > 0 and abs(o.get("qfrc_passive")[1]) < 1e-15


def test_fluid_off_by_default(capi, orc):
    m, o = _cube(capi, orc)
    o.set("qvel", np.ones(6))
    o.forward()
    assert np.all(o.get("qfrc_passive") == 0)


RANGE = """
<mujoco>
  <option gravity="0 0 0"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1"/>
    <geom name="ball" type="sphere" size="0.25" pos="2 0 1"/>
    <geom name="cap" type="capsule" size="0.1 0.3" pos="0 2 1" euler="0 90 0"/>
    <geom name="cyl" type="cylinder" size="0.2 0.4" pos="-2 0 1"/>
    <geom name="egg" type="ellipsoid" size="0.1 0.2 0.3" pos="0 -2 1"/>
    <geom name="brick" type="box" size="0.2 0.3 0.1" pos="0 0 3"/>
    <geom name="ghost" type="box" size="0.5 0.5 0.01" pos="0 0 0.5" rgba="1 0 0 0"/>
    <body name="probe" pos="0 0 1">
      <geom name="shell" type="sphere" size="0.05"/>
      <site name="down" pos="0 0 0" euler="180 0 0"/>
      <site name="up" pos="0 0 0"/>
      <site name="xp" pos="0 0 0" euler="0 90 0"/>
      <site name="xm" pos="0 0 0" euler="0 -90 0"/>
      <site name="yp" pos="0 0 0" euler="-90 0 0"/>
      <site name="ym" pos="0 0 0" euler="90 0 0"/>
      <site name="far" pos="20 0 0" euler="0 0 0"/>
    </body>
  </worldbody>
  <sensor>
    <rangefinder site="down"/><rangefinder site="up"/><rangefinder site="xp"/><rangefinder site="xm"/>
    <rangefinder site="yp"/><rangefinder site="ym"/><rangefinder site="far"/><rangefinder site="down" cutoff="0.4"/>
  </sensor>
</mujoco>
"""


def test_rangefinder_closed_forms(capi, orc):
    m = capi.Model.from_xml_string(RANGE)
    o = orc.Oracle(m)
    o.forward()
    s = o.get("sensordata")
    # down: the transparent slab at z = 0.5 and the probe's own shell are skipped -> the floor, 1 m below
    assert abs(s[0] - 1.0) < 1e-14
    assert abs(s[1] - (3 - 0.1 - 1)) < 1e-14            # up: bottom face of the brick
    assert abs(s[2] - (2 - 0.25)) < 1e-14               # +x: the ball
    assert abs(s[3] - (2 - 0.2)) < 1e-14                # -x: round side of the cylinder
    assert abs(s[4] - (2 - 0.1)) < 1e-14                # +y: round side of the capsule (axis along x)
    assert abs(s[5] - (2 - 0.2)) < 1e-14                # -y: the ellipsoid's y semi-axis
    assert s[6] == -1.0                                 # nothing above the far site
    assert abs(s[7] - 0.4) < 1e-15                      # cutoff clamps the 1 m reading


def test_ray_hits_caps_and_edges(capi, orc):
    import ctypes as C
    m = capi.Model.from_xml_string(RANGE)
    o = orc.Oracle(m)
    o.forward()
    lib = orc.lib
    lib.orc_ray.restype = C.c_double
    lib.orc_ray.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p]

    def ray(p, v, exclude=-1):
        p, v = np.asarray(p, float), np.asarray(v, float)
        gid = C.c_int(-7)
        x = lib.orc_ray(m.ptr, o._d, p.ctypes.data, v.ctypes.data, exclude, C.byref(gid))
        return x, gid.value

    gid = lambda n: m.name2id(capi.OBJ_GEOM, n)
    # capsule end cap, from +x along its axis: centre (0,2,1), half-length 0.3 along x, radius 0.1
    x, g = ray([1, 2, 1], [-1, 0, 0])
    assert g == gid("cap") and abs(x - (1 - 0.4)) < 1e-14
    # cylinder flat top from above
    x, g = ray([-2, 0, 5], [0, 0, -1])
    assert g == gid("cyl") and abs(x - (5 - 1.4)) < 1e-14
    # ellipsoid along z
    x, g = ray([0, -2, 4], [0, 0, -1])
    assert g == gid("egg") and abs(x - (4 - 1.3)) < 1e-14
    # non-unit direction: the distance is in units of |vec|
    x, g = ray([0, -2, 4], [0, 0, -2])
    assert abs(x - (4 - 1.3) / 2) < 1e-14
    # box from the side, then a miss just past its edge
    x, g = ray([3, 0.29, 3], [-1, 0, 0])
    assert g == gid("brick") and abs(x - 2.8) < 1e-14
    x, g = ray([3, 0.31, 3.2], [-1, 0, 0])
    assert g == -1 and x == -1
    # from inside a sphere: the far wall
    x, g = ray([2, 0, 1], [1, 0, 0])
    assert g == gid("ball") and abs(x - 0.25) < 1e-14
    # a plane seen from behind is not hit
    x, g = ray([0.9, 0.9, -1], [0, 0, 1], exclude=0)
    assert g != gid("floor")


SLOPE = """
<mujoco>
  <option timestep="0.002" gravity="{gx} 0 -9.81" cone="{cone}" solver="{solver}" noslip_iterations="{ns}" noslip_tolerance="1e-10"/>
  <worldbody>
    <geom type="plane" size="5 5 0.1" friction="1 0.005 0.0001"/>
    <body pos="0 0 0.1">
      <freejoint/>
      <geom type="box" size="0.1 0.1 0.1" friction="1 0.005 0.0001"/>
    </body>
  </worldbody>
</mujoco>
"""


@pytest.mark.parametrize("cone", ["pyramidal", "elliptic"])
@pytest.mark.parametrize("solver", ["PGS", "Newton"])
def test_noslip_removes_the_creep_of_soft_friction(capi, orc, cone, solver):
    """A box on a floor pushed sideways well inside the friction cone: the regularised contact model lets it creep,
    the noslip pass (viewer.cpp:590-591 exposes it) holds it.  Closed form: none -- the check is the ratio."""
    drift = {}
    for ns in (0, 20):
        m = capi.Model.from_xml_string(SLOPE.format(gx=3.0, cone=cone, solver=solver, ns=ns))
        o = orc.Oracle(m)
        o.step(500)
        drift[ns] = abs(o.get("qvel")[0])
        assert abs(o.get("qpos")[2] - 0.1) < 2e-3        # still resting on the floor
    assert drift[0] > 1e-4                               # soft friction creeps
    assert drift[20] < 0.02 * drift[0], drift            # noslip holds


def test_noslip_keeps_sliding_contacts_sliding(capi, orc):
    """Outside the cone (tangential pull 1.5 g with mu = 1) the box must still accelerate at about (1.5 - 1) g."""
    flat = SLOPE.replace('pos="0 0 0.1"', 'pos="0 0 0.01"').replace('size="0.1 0.1 0.1"', 'size="0.4 0.4 0.01"')  # cannot tip
    m = capi.Model.from_xml_string(flat.format(gx=1.5 * 9.81, cone="elliptic", solver="Newton", ns=20))
    o = orc.Oracle(m)
    o.step(250)
    a = o.get("qvel")[0] / (250 * 0.002)
    assert abs(a - 0.5 * 9.81) < 0.05 * 9.81, a


def test_gravity_compensation_closed_forms(capi, orc):
    """body gravcomp (mjModel.body_gravcomp, applied in mj_passive): a force -gravcomp * m * g at the body's centre of
    mass.  gravcomp = 1 floats a free body (no torque even with an off-centre geom), 0.5 halves its fall, 2 makes it rise;
    on a hinge the passive torque is the compensated share of the gravity torque; the gravity disable flag removes it."""
    xml = ('<mujoco><option %s/><worldbody>'
           '<body pos="0 0 1" gravcomp="%g"><freejoint/><geom type="box" size=".1 .05 .02" pos=".3 .1 0" mass="2"/></body>'
           '<body pos="1 0 1" gravcomp="0.75"><joint name="h" axis="0 1 0"/><geom type="capsule" fromto="0 0 0 .4 0 0" size=".02" mass="0.8"/>'
           '</body></worldbody></mujoco>')
    for gc, az in ((1.0, 0.0), (0.5, -9.81 / 2), (2.0, 9.81), (0.0, -9.81)):
        o = orc.Oracle(capi.Model.from_xml_string(xml % ("", gc)))
        o.forward()
        np.testing.assert_allclose(o.get("qacc")[:6], [0, 0, az, 0, 0, 0], atol=1e-12)
        # hinge: gravity torque m g l/2 about +y lowers the arm; 75 % of it is compensated
        np.testing.assert_allclose(o.get("qfrc_passive")[6], -0.75 * 0.8 * 9.81 * 0.2, rtol=1e-12)
        inertia = o.get("qM")[-1]
        np.testing.assert_allclose(o.get("qacc")[6], 0.25 * 0.8 * 9.81 * 0.2 / inertia, rtol=1e-12)
    o = orc.Oracle(capi.Model.from_xml_string(xml % ('><flag gravity="disable"/></option><option', 1.0)))
    o.forward()
    assert not o.get("qfrc_passive").any() and not o.get("qacc").any()
