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
