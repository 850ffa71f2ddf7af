"""Independent checks of the MJCF compiler and mj_setConst restatement (csrc/model/*.cpp).

The CUDA path and the CPU oracle are both fed the b2mjModel this compiler produces, so a wrong mass, inverse weight or
pair filter would be invisible to every GPU-vs-oracle test.  These tests recompute those quantities a second way, in
numpy, from first principles (closed-form primitive inertias, a dense inverse of the joint-space inertia assembled from
the oracle's qM, a brute-force enumeration of MuJoCo's collision filter rules) and compare."""
import itertools
import math

import numpy as np
import pytest

from conftest import model_path

MODELS = ["panda_like.xml", "hand_like.xml", "humanoid_like.xml", "bin.xml", "pendulum_scene.xml", "equality_scene.xml",
          "box_stack.xml", "actuated_arm.xml"]


def dense_M(model, qM):
    nv = model.nv
    M = np.zeros((nv, nv))
    for i in range(nv):
        adr, j = model.dof_Madr[i], i
        while j >= 0:
            M[i, j] = M[j, i] = qM[adr]
            adr += 1
            j = model.dof_parentid[j]
    return M


def body_jacobians(model, o, body, point):
    """3 x nv translational and rotational Jacobians of a world point attached to `body` (mj_jac semantics), from the
    oracle's cdof (spatial motion axes about the root subtree's centre of mass)."""
    nv = model.nv
    cdof = o.get("cdof").reshape(nv, 6)
    com = o.get("subtree_com").reshape(model.nbody, 3)[model.body_rootid[body]]
    jp, jr = np.zeros((3, nv)), np.zeros((3, nv))
    b = body
    while b > 0 and model.body_dofnum[b] == 0:
        b = model.body_parentid[b]
    if b == 0:
        return jp, jr
    k = model.body_dofadr[b] + model.body_dofnum[b] - 1
    while k >= 0:
        jr[:, k] = cdof[k, :3]
        jp[:, k] = cdof[k, 3:] + np.cross(cdof[k, :3], point - com)
        k = model.dof_parentid[k]
    return jp, jr


@pytest.mark.parametrize("name", MODELS)
def test_setconst_inverse_weights_from_dense_inverse(name, capi, orc):
    """dof_invweight0, body_invweight0, dof_M0 and stat.meaninertia against a dense inv(M) at qpos0 (mj_setConst /
    set0 semantics: dof weights are diag(inv M), averaged over the 3 translational / rotational dofs of free and ball
    joints; body weights are tr(J inv(M) J') / 3 for the translational and rotational Jacobians at the body's com)."""
    model = capi.Model.from_xml_file(model_path(name))
    if model.nv == 0:
        pytest.skip("no dofs")
    o = orc.Oracle(model)
    o.forward()
    M = dense_M(model, o.get("qM"))
    assert np.all(np.linalg.eigvalsh(M) > 0)
    Minv = np.linalg.inv(M)
    np.testing.assert_allclose(model.dof_M0, np.diag(M), rtol=1e-10)
    assert abs(model.stat.meaninertia - np.mean(np.diag(M))) < 1e-10 * np.mean(np.diag(M))
    want = np.zeros(model.nv)
    for j in range(model.njnt):
        t, d = model.jnt_type[j], model.jnt_dofadr[j]
        if t == 0:
            want[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
            want[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        elif t == 1:
            want[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
        else:
            want[d] = Minv[d, d]
    np.testing.assert_allclose(model.dof_invweight0, want, rtol=1e-8)
    xipos = o.get("xipos").reshape(model.nbody, 3)
    for b in range(1, model.nbody):
        jp, jr = body_jacobians(model, o, b, xipos[b])
        wt = np.trace(jp @ Minv @ jp.T) / 3
        wr = np.trace(jr @ Minv @ jr.T) / 3
        got = model.body_invweight0[b]
        np.testing.assert_allclose(got, [wt, wr], rtol=1e-7, atol=1e-12, err_msg=f"{name} body {b}")


def primitive(kind, size, density):
    if kind == "sphere":
        r = size[0]
        m = density * 4 / 3 * math.pi * r ** 3
        return m, [0.4 * m * r * r] * 3
    if kind == "box":
        a, b, c = size
        m = density * 8 * a * b * c
        return m, [m / 3 * (b * b + c * c), m / 3 * (a * a + c * c), m / 3 * (a * a + b * b)]
    if kind == "capsule":
        r, h = size[0], size[1]      # half-length h of the cylinder
        mc = density * math.pi * r * r * 2 * h
        ms = density * 4 / 3 * math.pi * r ** 3
        m = mc + ms
        iz = mc * r * r / 2 + ms * 0.4 * r * r
        # hemispheres: own inertia about the sphere centre 2/5 ms r^2 (both halves), shifted by h with the 3r/8 offset
        ix = mc * (r * r / 4 + h * h / 3) + ms * (0.4 * r * r + h * h + 0.75 * h * r)
        return m, [ix, ix, iz]
    if kind == "cylinder":
        r, h = size[0], size[1]
        m = density * math.pi * r * r * 2 * h
        return m, [m * (r * r / 4 + h * h / 3)] * 2 + [m * r * r / 2]
    raise ValueError(kind)


@pytest.mark.parametrize("kind,size", [("sphere", (0.07,)), ("box", (0.05, 0.08, 0.11)), ("capsule", (0.04, 0.13)),
                                       ("capsule", (0.09, 0.02))])
def test_inertia_from_geom_closed_forms(kind, size, capi):
    density = 730.0
    xml = f"""<mujoco><worldbody><body name="b" pos="0.1 0.2 0.3">
      <freejoint/><geom type="{kind}" size="{' '.join(str(s) for s in size)}" density="{density}"/>
    </body></worldbody></mujoco>"""
    m = capi.Model.from_xml_string(xml)
    mass, inertia = primitive(kind, size, density)
    assert abs(m.body_mass[1] - mass) < 1e-12 * mass
    np.testing.assert_allclose(sorted(m.body_inertia[1]), sorted(inertia), rtol=1e-12)
    np.testing.assert_allclose(m.body_ipos[1], 0, atol=1e-15)


def test_composite_body_parallel_axis(capi):
    """two spheres on one body: total mass, com and inertia by the parallel-axis theorem"""
    xml = """<mujoco><worldbody><body name="b"><freejoint/>
      <geom type="sphere" size="0.05" pos="0.2 0 0" density="1000"/>
      <geom type="sphere" size="0.08" pos="-0.1 0 0" density="500"/>
    </body></worldbody></mujoco>"""
    m = capi.Model.from_xml_string(xml)
    m1, i1 = primitive("sphere", (0.05,), 1000)
    m2, i2 = primitive("sphere", (0.08,), 500)
    M = m1 + m2
    cx = (m1 * 0.2 + m2 * -0.1) / M
    assert abs(m.body_mass[1] - M) < 1e-12
    np.testing.assert_allclose(m.body_ipos[1], [cx, 0, 0], atol=1e-14)
    ixx = i1[0] + i2[0]
    iyy = i1[0] + m1 * (0.2 - cx) ** 2 + i2[0] + m2 * (-0.1 - cx) ** 2
    np.testing.assert_allclose(sorted(m.body_inertia[1]), sorted([ixx, iyy, iyy]), rtol=1e-12)


def brute_force_pairs(model):
    """MuJoCo's broadphase-independent filter rules (engine_collision_driver.c: mj_collision / canCollide /
    filterBodyPair), enumerated over all geom pairs."""
    filterparent = not (model.opt.disableflags & (1 << 9))
    excl = set(int(s) for s in model.exclude_signature)
    out = set()
    for g1, g2 in itertools.combinations(range(model.ngeom), 2):
        b1, b2 = int(model.geom_bodyid[g1]), int(model.geom_bodyid[g2])
        w1, w2 = int(model.body_weldid[b1]), int(model.body_weldid[b2])
        if w1 == w2:                                        # same (welded) body, incl. both static
            continue
        if ((min(b1, b2) << 16) + max(b1, b2)) in excl:
            continue
        if filterparent and w1 != 0 and w2 != 0:
            p1, p2 = int(model.body_weldid[model.body_parentid[w1]]), int(model.body_weldid[model.body_parentid[w2]])
            if p1 == w2 or p2 == w1:
                continue
        ct1, ca1, ct2, ca2 = (int(model.geom_contype[g1]), int(model.geom_conaffinity[g1]), int(model.geom_contype[g2]),
                              int(model.geom_conaffinity[g2]))
        if not ((ct1 & ca2) or (ct2 & ca1)):
            continue
        # the narrowphase table is upper triangular in geom type: the lower type comes first
        a, b = (g1, g2) if model.geom_type[g1] <= model.geom_type[g2] else (g2, g1)
        out.add((a, b))
    return out


@pytest.mark.parametrize("name", MODELS)
def test_collision_pair_table_matches_brute_force_filter(name, capi):
    model = capi.Model.from_xml_file(model_path(name))
    got = set(zip((int(x) for x in model.collpair_geom1), (int(x) for x in model.collpair_geom2)))
    want = brute_force_pairs(model)
    # plane-plane and other pairs without a narrowphase are dropped by the compiler: compare on supported type pairs
    def supported(p):
        t1, t2 = model.geom_type[p[0]], model.geom_type[p[1]]
        return not (t1 == 0 and t2 == 0)
    want = {p for p in want if supported(p)}
    assert got == want, (sorted(got - want)[:5], sorted(want - got)[:5])
    # slot addresses: cumulative, room for every pair's maximum contact count
    adr = 0
    for k in range(model.ncollpair):
        assert model.collpair_slotadr[k] == adr
        adr += model.collpair_maxcon[k]


def test_contact_parameter_mixing_rules(capi, orc):
    """mj_contactParam: friction = element-wise max, solref / solimp mixed by solmix weights (equal priority), condim =
    max; the higher priority geom wins outright."""
    xml = """<mujoco><option cone="elliptic"/><worldbody>
      <geom name="floor" type="plane" size="1 1 .1" friction="0.6 0.01 0.002" solref="0.02 1" solimp="0.9 0.95 0.001 0.5 2" solmix="1"/>
      <body pos="0 0 0.049"><freejoint/>
        <geom name="a" type="sphere" size="0.05" friction="1.2 0.004 0.0005" solref="0.01 0.8" solimp="0.8 0.9 0.002 0.4 3" solmix="3" condim="4"/></body>
      <body pos="0.5 0 0.049"><freejoint/>
        <geom name="b" type="sphere" size="0.05" friction="0.1 0.1 0.1" solref="0.005 0.5" priority="2" condim="1"/></body>
    </worldbody></mujoco>"""
    m = capi.Model.from_xml_string(xml)
    o = orc.Oracle(m)
    o.forward()
    assert o.get("ncon")[0] == 2
    g2 = o.get("contact_geom2")[:2]
    fr = o.get("contact_friction").reshape(-1, 5)[:2]
    sr = o.get("contact_solref").reshape(-1, 2)[:2]
    si = o.get("contact_solimp").reshape(-1, 5)[:2]
    dim = o.get("contact_dim")[:2]
    ia = list(g2).index(m.name2id(capi.OBJ_GEOM, "a"))
    ib = 1 - ia
    # floor + a: equal priority -> max friction (tangent, tangent, torsion, roll, roll), weighted solref / solimp
    np.testing.assert_allclose(fr[ia], [1.2, 1.2, 0.01, 0.002, 0.002])
    w = 1 / (1 + 3)
    np.testing.assert_allclose(sr[ia], [w * 0.02 + (1 - w) * 0.01, w * 1 + (1 - w) * 0.8])
    np.testing.assert_allclose(si[ia], w * np.array([0.9, 0.95, 0.001, 0.5, 2]) + (1 - w) * np.array([0.8, 0.9, 0.002, 0.4, 3]))
    assert dim[ia] == 4
    # floor + b: b has priority 2 -> its parameters, its condim
    np.testing.assert_allclose(fr[ib], [0.1, 0.1, 0.1, 0.1, 0.1])
    np.testing.assert_allclose(sr[ib], [0.005, 0.5])
    assert dim[ib] == 1
