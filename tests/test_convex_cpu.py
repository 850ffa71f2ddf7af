"""CPU oracle: cylinder / ellipsoid narrowphase (plane-cylinder, plane-convex, general convex pairs by Minkowski
Portal Refinement) against closed-form contact answers and against the dedicated primitive functions where a convex
pair degenerates to one (an ellipsoid with equal semi-axes is a sphere).  These pairs live in MuJoCo's collision
table, which the reference reaches through mj_step and lets plugins override (mujoco_env.cpp:163-176, 498)."""
import numpy as np
import pytest

SCENE = """
<mujoco>
  <option gravity="0 0 0"/>
  <worldbody>
    {floor}
    <body name="a" pos="{pa}" {ea}><freejoint/><geom name="ga" {ga}/></body>
    <body name="b" pos="{pb}" {eb}><freejoint/><geom name="gb" {gb}/></body>
  </worldbody>
</mujoco>
"""
FLOOR = '<geom name="floor" type="plane" size="5 5 0.1"/>'


def contacts(capi, orc, **kw):
    kw.setdefault("floor", "")
    kw.setdefault("ea", "")
    kw.setdefault("eb", "")
    kw.setdefault("pb", "50 50 50")
    kw.setdefault("gb", 'type="sphere" size="0.01"')
    m = capi.Model.from_xml_string(SCENE.format(**kw))
    o = orc.Oracle(m)
    o.forward()
    n = int(o.get("ncon")[0])
    return (m, n, o.get("contact_dist")[:n], o.get("contact_pos").reshape(-1, 3)[:n],
            o.get("contact_frame").reshape(-1, 9)[:n], o.get("contact_geom1")[:n], o.get("contact_geom2")[:n])


def test_cylinder_standing_on_plane(capi, orc):
    m, n, dist, pos, frame, g1, g2 = contacts(capi, orc, floor=FLOOR, pa="0 0 0.39", ga='type="cylinder" size="0.2 0.4"')
    # flat cap 1 cm into the floor: the rim point plus two points at +-120 degrees (far cap out of range)
    assert n == 3
    np.testing.assert_allclose(dist, -0.01, atol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(pos[:, :2], axis=1), 0.2, atol=1e-14)
    np.testing.assert_allclose(pos[:, 2], -0.005, atol=1e-14)
    np.testing.assert_allclose(frame[:, :3], [[0, 0, 1]] * 3, atol=1e-15)
    # the three points form an equilateral triangle
    d = [np.linalg.norm(pos[i] - pos[j]) for i, j in ((0, 1), (0, 2), (1, 2))]
    np.testing.assert_allclose(d, 0.2 * np.sqrt(3), atol=1e-12)


def test_cylinder_lying_on_plane(capi, orc):
    m, n, dist, pos, frame, *_ = contacts(capi, orc, floor=FLOOR, pa="0 0 0.19", ea='euler="0 90 0"',
                                          ga='type="cylinder" size="0.2 0.4"')
    # axis along x: line contact -> the lowest point of each cap rim
    assert n == 2
    np.testing.assert_allclose(dist, -0.01, atol=1e-12)
    np.testing.assert_allclose(sorted(pos[:, 0]), [-0.4, 0.4], atol=1e-12)
    np.testing.assert_allclose(pos[:, 1], 0, atol=1e-12)


def test_tilted_cylinder_touches_with_its_rim(capi, orc):
    th = np.deg2rad(30)
    h = 0.4 * np.cos(th) + 0.2 * np.sin(th)    # height of the centre when the rim touches
    m, n, dist, pos, frame, *_ = contacts(capi, orc, floor=FLOOR, pa=f"0 0 {h - 0.002}", ea='euler="0 30 0"',
                                          ga='type="cylinder" size="0.2 0.4"')
    assert n == 1
    assert abs(dist[0] + 0.002) < 1e-12


def test_ellipsoid_on_plane(capi, orc):
    m, n, dist, pos, frame, *_ = contacts(capi, orc, floor=FLOOR, pa="0.3 -0.2 0.29", ga='type="ellipsoid" size="0.1 0.2 0.3"')
    assert n == 1 and abs(dist[0] + 0.01) < 1e-14
    np.testing.assert_allclose(pos[0], [0.3, -0.2, -0.005], atol=1e-14)
    # rotated by 90 degrees about y the x semi-axis (0.1) points down
    m, n, dist, *_ = contacts(capi, orc, floor=FLOOR, pa="0 0 0.095", ea='euler="0 90 0"', ga='type="ellipsoid" size="0.1 0.2 0.3"')
    assert n == 1 and abs(dist[0] + 0.005) < 1e-14
    # tilted: support height of an ellipsoid along n is sqrt(sum (s_i n_i)^2)
    th = np.deg2rad(25)
    hgt = np.sqrt((0.1 * np.sin(th)) ** 2 + (0.3 * np.cos(th)) ** 2)
    m, n, dist, *_ = contacts(capi, orc, floor=FLOOR, pa=f"0 0 {hgt - 0.003}", ea='euler="0 25 0"', ga='type="ellipsoid" size="0.1 0.2 0.3"')
    assert n == 1 and abs(dist[0] + 0.003) < 1e-13


@pytest.mark.parametrize("other,reach", [('type="sphere" size="0.15"', 0.15), ('type="capsule" size="0.1 0.2"', 0.3),
                                         ('type="box" size="0.2 0.2 0.25"', 0.25), ('type="cylinder" size="0.2 0.3"', 0.3)])
def test_round_ellipsoid_behaves_like_a_sphere(capi, orc, other, reach):
    """ellipsoid with equal semi-axes r above the top of another shape: depth, normal and position of sphere contact."""
    r, pen = 0.12, 0.004
    z = reach + r - pen
    m, n, dist, pos, frame, g1, g2 = contacts(capi, orc, pa="0 0 0", ga=other, pb=f"0.01 0.02 {z}",
                                              gb=f'type="ellipsoid" size="{r} {r} {r}"')
    assert n == 1
    flat = "sphere" not in other and "capsule" not in other
    if flat:   # flat top: exact answer
        assert abs(dist[0] + pen) < 1e-6
        np.testing.assert_allclose(np.abs(frame[0, :3]), [0, 0, 1], atol=1e-6)
        assert abs(pos[0][2] - (reach - pen / 2)) < 2e-3   # findPos interpolates the portal witnesses: approximate
    else:      # curved top, slightly off-axis: compare with the sphere-sphere formula
        c1 = np.array([0, 0, reach - (0.15 if "sphere" in other else 0.1)])
        r1 = 0.15 if "sphere" in other else 0.1
        dvec = np.array([0.01, 0.02, z]) - c1
        assert abs(dist[0] - (np.linalg.norm(dvec) - r1 - r)) < 1e-6
        sgn = 1.0 if m.geom_type[g1[0]] != 4 else -1.0   # normal points from geom1 to geom2
        np.testing.assert_allclose(frame[0, :3], sgn * dvec / np.linalg.norm(dvec), atol=5e-3)  # MPR stops at mpr_tolerance


def test_mpr_normal_points_from_geom1_to_geom2_and_margin_reports_positive_distance(capi, orc):
    xml = SCENE.format(floor="", pa="0 0 0", ea="", ga='type="ellipsoid" size="0.1 0.1 0.2" margin="0.05"',
                       pb="0 0 0.42", eb="", gb='type="cylinder" size="0.15 0.2" margin="0.05"')
    m = capi.Model.from_xml_string(xml)
    o = orc.Oracle(m)
    o.forward()
    assert o.get("ncon")[0] == 1
    # gap between the ellipsoid's top (z = 0.2) and the cylinder's bottom (z = 0.22): +0.02, inside the 0.05 margin
    assert abs(o.get("contact_dist")[0] - 0.02) < 1e-6
    fr = o.get("contact_frame")[:3]
    assert m.geom_type[o.get("contact_geom1")[0]] == 4 and fr[2] > 0.999999
    assert abs(o.get("contact_pos")[2] - 0.21) < 1e-6


def test_separated_convex_pairs_make_no_contact(capi, orc):
    m, n, *_ = contacts(capi, orc, pa="0 0 0", ga='type="ellipsoid" size="0.1 0.2 0.3"', pb="0.5 0 0",
                        gb='type="cylinder" size="0.1 0.2"')
    assert n == 0


def test_cylinder_box_edge_and_face(capi, orc):
    # cylinder lying across a box top: axis along y, 3 mm deep
    m, n, dist, pos, frame, *_ = contacts(capi, orc, pa="0 0 0", ga='type="box" size="0.3 0.3 0.1"', pb="0 0 0.247",
                                          eb='euler="90 0 0"', gb='type="cylinder" size="0.15 0.2"')
    assert n == 1 and abs(dist[0] + 0.003) < 1e-6
    np.testing.assert_allclose(np.abs(frame[0, :3]), [0, 0, 1], atol=1e-5)


def test_resting_heights_over_time(capi, orc):
    """A cylinder (standing) and an ellipsoid dropped on the floor come to rest at the analytic soft-contact depth."""
    xml = """<mujoco><option timestep="0.002"/><worldbody><geom type="plane" size="5 5 0.1"/>
      <body pos="0 0 0.45"><freejoint/><geom type="cylinder" size="0.2 0.4" density="500"/></body>
      <body pos="1 0 0.35"><freejoint/><geom type="ellipsoid" size="0.2 0.25 0.3" density="500"/></body>
      </worldbody></mujoco>"""
    m = capi.Model.from_xml_string(xml)
    o = orc.Oracle(m)
    o.step(1500)
    q = o.get("qpos")
    assert abs(o.get("qvel")).max() < 1e-3
    assert 0.39 < q[2] < 0.4001 and 0.29 < q[9] < 0.3001
