"""Closed-form contact geometry of the primitive pairs the bench models consist of (row M5): distance, normal and contact
point of sphere / capsule / plane / box pairs follow from elementary geometry (MuJoCo conventions: dist < 0 when
penetrating, the normal points from geom1 to geom2, the point sits midway between the two surfaces)."""
import numpy as np
import pytest


def contacts(capi, orc, geoms, margin=0.0):
    bodies = "".join(f'<body pos="{p}" {e}><freejoint/><geom {g} margin="{margin}"/></body>' if p else f"<geom {g} margin=\"{margin}\"/>"
                     for p, e, g in geoms)
    m = capi.Model.from_xml_string(f'<mujoco><compiler angle="radian"/><worldbody>{bodies}</worldbody></mujoco>')
    o = orc.Oracle(m)
    o.forward()
    n = int(o.get("ncon")[0])
    return [dict(dist=o.get("contact_dist")[c], pos=o.get("contact_pos")[3 * c:3 * c + 3].copy(),
                 normal=o.get("contact_frame")[9 * c:9 * c + 3].copy()) for c in range(n)]


def test_sphere_sphere(capi, orc):
    c1, c2, r1, r2 = np.array([0.1, -0.2, 1.0]), np.array([0.25, -0.05, 1.12]), 0.15, 0.12
    (c,) = contacts(capi, orc, [(" ".join(map(str, c1)), "", f'size="{r1}"'), (" ".join(map(str, c2)), "", f'size="{r2}"')])
    d = np.linalg.norm(c2 - c1)
    n = (c2 - c1) / d
    np.testing.assert_allclose(c["dist"], d - r1 - r2, atol=1e-14)
    np.testing.assert_allclose(c["normal"], n, atol=1e-14)
    np.testing.assert_allclose(c["pos"], c1 + n * (r1 + 0.5 * (d - r1 - r2)), atol=1e-14)
    # separated by more than the margin: nothing; inside the margin: a contact with positive distance
    far = [("0 0 1", "", 'size="0.1"'), ("0.25 0 1", "", 'size="0.1"')]
    assert contacts(capi, orc, far) == []
    assert contacts(capi, orc, far, margin=0.03) == []  # the pair's margin is the larger of the two, not their sum
    (c,) = contacts(capi, orc, far, margin=0.06)
    np.testing.assert_allclose(c["dist"], 0.05, atol=1e-14)


def test_plane_sphere_and_plane_capsule(capi, orc):
    plane = (None, "", 'type="plane" size="1 1 .1"')
    (c,) = contacts(capi, orc, [plane, ("0.3 0.2 0.08", "", 'size="0.1"')])
    np.testing.assert_allclose([c["dist"], *c["normal"], *c["pos"]], [-0.02, 0, 0, 1, 0.3, 0.2, -0.01], atol=1e-14)
    # capsule tilted by 30 degrees about y: its lower cap touches, the upper one is out of reach
    th, r, h = np.pi / 6, 0.05, 0.2
    axis = np.array([np.sin(th), 0, np.cos(th)])
    centre = np.array([0, 0, 0.2])
    (c,) = contacts(capi, orc, [plane, ("0 0 0.2", f'euler="0 {th} 0"', f'type="capsule" size="{r} {h}"')])
    low = centre - h * axis
    np.testing.assert_allclose(c["dist"], low[2] - r, atol=1e-14)
    np.testing.assert_allclose(c["pos"], [low[0], low[1], 0.5 * (low[2] - r)], atol=1e-14)
    # lying flat: both caps touch at the same depth
    cs = contacts(capi, orc, [plane, ("0 0 0.04", f'euler="0 {np.pi / 2} 0"', f'type="capsule" size="{r} {h}"')])
    assert len(cs) == 2
    np.testing.assert_allclose([c["dist"] for c in cs], [-0.01, -0.01], atol=1e-14)
    np.testing.assert_allclose(sorted(c["pos"][0] for c in cs), [-h, h], atol=1e-14)


def test_sphere_capsule_and_capsule_capsule(capi, orc):
    # sphere beside the cylindrical part, and beyond the cap (nearest point = cap centre)
    for centre, nearest in (([0.12, 0.05, 0.1], [0, 0, 0.1]), ([0.05, 0, 0.38], [0, 0, 0.3])):
        (c,) = contacts(capi, orc, [("0 0 0", "", 'type="capsule" size="0.05 0.3"'), (" ".join(map(str, centre)), "", 'size="0.1"')])
        # geoms are ordered by type: sphere (2) comes before capsule (3), so the normal points from the sphere
        v = np.array(nearest) - np.array(centre)
        d = np.linalg.norm(v)
        np.testing.assert_allclose(c["dist"], d - 0.15, atol=1e-14)
        np.testing.assert_allclose(c["normal"], v / d, atol=1e-14)
    # two skew capsules: axes z through the origin and x through (0, 0.08, 0.05): common normal along y
    (c,) = contacts(capi, orc, [("0 0 0", "", 'type="capsule" size="0.05 0.3"'),
                                ("0 0.08 0.05", f'euler="0 {np.pi / 2} 0"', 'type="capsule" size="0.04 0.3"')])
    np.testing.assert_allclose(c["dist"], 0.08 - 0.09, atol=1e-14)
    np.testing.assert_allclose(c["normal"], [0, 1, 0], atol=1e-14)
    np.testing.assert_allclose(c["pos"], [0, 0.045, 0.05], atol=1e-14)


def test_plane_box(capi, orc):
    # a box tilted about x and y: the corners below the plane are the contacts, each at its own depth
    from scipy.spatial.transform import Rotation

    rot = Rotation.from_euler("xyz", [0.2, -0.15, 0.4])
    size, centre = np.array([0.1, 0.07, 0.05]), np.array([0.3, -0.1, 0.06])
    q = rot.as_quat()  # x y z w
    cs = contacts(capi, orc, [(None, "", 'type="plane" size="1 1 .1"'),
                              (" ".join(map(str, centre)), f'quat="{q[3]} {q[0]} {q[1]} {q[2]}"', 'type="box" size="0.1 0.07 0.05"')])
    corners = [centre + rot.apply(size * s) for s in np.array(np.meshgrid([-1, 1], [-1, 1], [-1, 1])).T.reshape(-1, 3)]
    below = sorted(c[2] for c in corners if c[2] < 0)
    assert 1 <= len(below) <= 4 and len(cs) == len(below)
    np.testing.assert_allclose(sorted(c["dist"] for c in cs), below, atol=1e-13)
    for c in cs:
        np.testing.assert_allclose(c["normal"], [0, 0, 1], atol=1e-14)
        corner = min(corners, key=lambda p: abs(p[2] - c["dist"]))
        np.testing.assert_allclose(c["pos"], [corner[0], corner[1], 0.5 * corner[2]], atol=1e-13)
