"""Mesh assets: file readers, convex hull, mass properties, frame folding, and the convex narrowphase on
hull vertices -- against closed forms and against the box primitive a cube mesh must reproduce.  The reference loads
meshes through mj_loadXML (mujoco_env.cpp:771-911)."""
import itertools
import os
import struct

import numpy as np
import pytest

CUBE = " ".join(f"{x} {y} {z}" for x, y, z in itertools.product((-0.1, 0.1), (-0.2, 0.2), (-0.3, 0.3)))
INNER = " 0 0 0 0.05 0.05 0.05 -0.02 0.1 0.2"          # interior points must not reach the hull


def scene(asset, geom, pos="0 0 1", extra=""):
    return f"""<mujoco><option gravity="0 0 -9.81"/><asset>{asset}</asset><worldbody>
      <geom name="floor" type="plane" size="3 3 0.1"/>{extra}
      <body name="b" pos="{pos}"><freejoint/><geom name="g" {geom}/></body></worldbody></mujoco>"""


def test_cube_mesh_equals_box(capi):
    mm = capi.Model.from_xml_string(scene(f'<mesh name="c" vertex="{CUBE}{INNER}"/>', 'type="mesh" mesh="c" density="700"'))
    mb = capi.Model.from_xml_string(scene("", 'type="box" size="0.1 0.2 0.3" density="700"'))
    assert mm.nmesh == 1 and mm.mesh_vertnum[0] == 8 and mm.geom_type[1] == 7 and mm.geom_dataid[1] == 0
    assert abs(mm.body_mass[1] - mb.body_mass[1]) < 1e-12
    np.testing.assert_allclose(sorted(mm.body_inertia[1]), sorted(mb.body_inertia[1]), rtol=1e-12)
    np.testing.assert_allclose(mm.body_ipos[1], 0, atol=1e-15)
    # principal axes sorted by decreasing moment: the longest side (z, 0.3) becomes the last axis
    np.testing.assert_allclose(mm.geom_size[1], [0.1, 0.2, 0.3], atol=1e-12)
    np.testing.assert_allclose(mm.geom_rbound[1], np.linalg.norm([0.1, 0.2, 0.3]), rtol=1e-12)
    v = mm.mesh_vert.reshape(-1, 3)
    assert sorted(map(tuple, np.round(np.abs(v), 12))) == [(0.1, 0.2, 0.3)] * 8


def test_offset_mesh_folds_its_frame_into_the_geom(capi):
    shifted = " ".join(f"{x + 1.0} {y - 2.0} {z + 0.5}" for x, y, z in itertools.product((-0.1, 0.1), (-0.2, 0.2), (-0.3, 0.3)))
    m = capi.Model.from_xml_string(scene(f'<mesh name="c" vertex="{shifted}"/>', 'mesh="c" pos="0.1 0 0"'))
    np.testing.assert_allclose(m.geom_pos[1], [1.1, -2.0, 0.5], atol=1e-12)   # geom pos + centre of mass of the asset
    np.testing.assert_allclose(m.body_ipos[1], [1.1, -2.0, 0.5], atol=1e-12)
    np.testing.assert_allclose(m.mesh_vert.reshape(-1, 3).mean(0), 0, atol=1e-12)


def test_tetrahedron_mass_properties(capi):
    m = capi.Model.from_xml_string(scene('<mesh name="t" vertex="0 0 0  1 0 0  0 1 0  0 0 1"/>', 'mesh="t" density="600"'))
    assert abs(m.body_mass[1] - 600 / 6) < 1e-10                  # V = 1/6
    np.testing.assert_allclose(m.body_ipos[1], [0.25, 0.25, 0.25], atol=1e-12)
    # inertia tensor of the unit right tetrahedron about its centre of mass: eigenvalues rho*V*{3/40+..}; check trace and
    # the moment about the (1,1,1) symmetry axis via the covariance: C_ii = V*(1/10 - 1/16), C_ij = V*(1/20 - 1/16)
    V = 1 / 6
    C = V * (np.full((3, 3), 1 / 20 - 1 / 16) + np.eye(3) * (1 / 10 - 1 / 20))
    I = 600 * (np.trace(C) * np.eye(3) - C)
    np.testing.assert_allclose(sorted(m.body_inertia[1]), sorted(np.linalg.eigvalsh(I)), rtol=1e-10)


def test_scale_and_hull_of_random_points(capi):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(200, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)               # on the unit sphere: every point is a hull vertex
    pts = np.vstack([pts, rng.uniform(-0.3, 0.3, (50, 3))])         # plus interior points
    vs = " ".join(f"{x:.17g}" for x in pts.ravel())
    m = capi.Model.from_xml_string(scene(f'<mesh name="s" vertex="{vs}" scale="0.2 0.2 0.2"/>', 'mesh="s"'))
    assert m.mesh_vertnum[0] == 200
    assert 0.9 * 4 / 3 * np.pi * 0.2 ** 3 < m.body_mass[1] / 1000 < 4 / 3 * np.pi * 0.2 ** 3
    np.testing.assert_allclose(np.linalg.norm(m.mesh_vert.reshape(-1, 3) + 0, axis=1), 0.2, atol=0.01)


def _write_stl(path, tris, binary=True):
    if binary:
        with open(path, "wb") as f:
            f.write(b"\0" * 80 + struct.pack("<I", len(tris)))
            for t in tris:
                f.write(struct.pack("<12fH", 0, 0, 0, *np.asarray(t, np.float32).ravel(), 0))
    else:
        with open(path, "w") as f:
            f.write("solid x\n")
            for t in tris:
                f.write("facet normal 0 0 0\nouter loop\n" + "".join(f"vertex {v[0]} {v[1]} {v[2]}\n" for v in t) + "endloop\nendfacet\n")
            f.write("endsolid x\n")


@pytest.mark.parametrize("fmt", ["stl-binary", "stl-ascii", "obj"])
def test_mesh_files(capi, tmp_path, fmt):
    c = np.array(list(itertools.product((-0.5, 0.5), (-0.25, 0.25), (-0.125, 0.125))))
    faces = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4), (1, 5, 7), (1, 7, 3)]
    d = tmp_path / "assets"
    d.mkdir()
    if fmt == "obj":
        (d / "brick.obj").write_text("# brick\n" + "".join(f"v {v[0]} {v[1]} {v[2]}\n" for v in c) +
                                     "".join(f"f {a + 1} {b + 1} {c_ + 1}\n" for a, b, c_ in faces))
        fname = "brick.obj"
    else:
        _write_stl(str(d / "brick.stl"), [[c[i] for i in f] for f in faces], binary=fmt == "stl-binary")
        fname = "brick.stl"
    xml = scene(f'<mesh file="{fname}"/>', 'mesh="brick"').replace("<option", '<compiler meshdir="assets"/><option')
    path = tmp_path / "model.xml"
    path.write_text(xml)
    m = capi.Model.from_xml_file(str(path))
    assert m.mesh_vertnum[0] == 8 and abs(m.body_mass[1] - 1000 * 1.0 * 0.5 * 0.25) < 1e-4
    with pytest.raises(capi.B2mjError, match="cannot open mesh file"):
        capi.Model.from_xml_string(xml)          # no model directory: the relative file cannot be found


def test_cube_mesh_contacts_match_the_box_primitive(capi, orc):
    """Resting on the floor: plane-mesh reports the four bottom corners like plane-box; on top of a box: MPR depth equals
    the box-box depth."""
    for geom, asset in (('type="mesh" mesh="c"', f'<mesh name="c" vertex="{CUBE}"/>'), ('type="box" size="0.1 0.2 0.3"', "")):
        m = capi.Model.from_xml_string(scene(asset, geom, pos="0 0 0.295"))
        o = orc.Oracle(m)
        o.forward()
        n = int(o.get("ncon")[0])
        assert n == 4
        np.testing.assert_allclose(o.get("contact_dist")[:4], -0.005, atol=1e-12)
        p = o.get("contact_pos").reshape(-1, 3)[:4]
        assert sorted(map(tuple, np.round(np.abs(p[:, :2]), 9))) == [(0.1, 0.2)] * 4
    plat = '<geom name="plat" type="box" size="0.5 0.5 0.1" pos="0 0 0.1"/>'
    m = capi.Model.from_xml_string(scene(f'<mesh name="c" vertex="{CUBE}"/>', 'mesh="c"', pos="0.05 0.02 0.497", extra=plat))
    o = orc.Oracle(m)
    o.forward()
    assert int(o.get("ncon")[0]) == 1
    assert abs(o.get("contact_dist")[0] + 0.003) < 1e-6
    assert abs(abs(o.get("contact_frame")[2]) - 1) < 1e-6


def test_mesh_settles_on_the_floor(capi, orc):
    m = capi.Model.from_xml_string(scene(f'<mesh name="c" vertex="{CUBE}"/>', 'mesh="c"', pos="0 0 0.4").replace(
        "<option", '<option timestep="0.002"'))
    o = orc.Oracle(m)
    o.step(1500)
    assert abs(o.get("qvel")).max() < 1e-3 and 0.29 < o.get("qpos")[2] < 0.3001


def test_degenerate_meshes_are_rejected(capi):
    with pytest.raises(capi.B2mjError, match="coplanar"):
        capi.Model.from_xml_string(scene('<mesh name="p" vertex="0 0 0 1 0 0 0 1 0 1 1 0"/>', 'mesh="p"'))
    with pytest.raises(capi.B2mjError, match="unknown mesh"):
        capi.Model.from_xml_string(scene("", 'type="mesh" mesh="nope"'))
