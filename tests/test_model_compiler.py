"""MJCF compiler facts pinned by the reference's tests and by closed-form values."""
import numpy as np
import pytest


def test_pendulum_qpos0_pinned(load_model):
    # reference ros_interface_test.cpp:290-298
    m = load_model("pendulum_scene.xml")
    assert (m.nq, m.nv, m.nbody, m.njnt, m.ngeom) == (13, 11, 6, 4, 6)
    np.testing.assert_array_equal(m.qpos0, [1, 0, 0, 0, 0, 0, 1, 0, 0.06, 1, 0, 0, 0])
    assert m.opt.timestep == 0.001 and m.opt.cone == 1 and m.opt.solver == 2 and m.opt.integrator == 0
    np.testing.assert_array_equal(m.opt.gravity[:], [0, 0, -9.81])


def test_enum_values_pinned(capi):
    # mujoco_env_fixture.h:145-156, GeomType.msg:2-9, EqualityConstraintType.msg:2-5
    import re, os
    from conftest import ROOT
    h = open(os.path.join(ROOT, "include", "b2mj.h")).read()
    for name, val in (("B2MJ_NEQDATA", 11), ("B2MJ_NIMP", 5), ("B2MJ_NREF", 2)):
        assert re.search(rf"#define {name} {val}\b", h)
    assert "B2MJ_EQ_CONNECT = 0, B2MJ_EQ_WELD = 1, B2MJ_EQ_JOINT = 2, B2MJ_EQ_TENDON = 3" in h
    assert "B2MJ_GEOM_PLANE = 0, B2MJ_GEOM_HFIELD = 1, B2MJ_GEOM_SPHERE = 2, B2MJ_GEOM_CAPSULE = 3" in h


def test_capsule_inertia_closed_form(capi, load_model):
    # density 1000 capsules
    m = load_model("pendulum_scene.xml")
    b = m.name2id(capi.OBJ_BODY, "base_link")
    assert m.body_mass[b] == pytest.approx(5.428672, rel=1e-6)
    assert sorted(m.body_inertia[b]) == pytest.approx([0.00944589, 0.11002712, 0.11002712], rel=1e-6)
    np.testing.assert_allclose(m.body_ipos[b], [0, 0, 0.8], atol=1e-12)
    b = m.name2id(capi.OBJ_BODY, "middle_link")
    assert m.body_mass[b] == pytest.approx(1.776047, rel=1e-6)
    b = m.name2id(capi.OBJ_BODY, "end_link")
    assert m.body_mass[b] == pytest.approx(0.284838, rel=2e-6)
    assert sorted(m.body_inertia[b]) == pytest.approx([0.00005563, 0.00125362, 0.00125362], rel=2e-4)
    b = m.name2id(capi.OBJ_BODY, "body_ball")
    assert m.body_mass[b] == pytest.approx(0.1)
    np.testing.assert_allclose(m.body_inertia[b], 1e-4, rtol=1e-12)
    b = m.name2id(capi.OBJ_BODY, "immovable")
    np.testing.assert_allclose(m.body_inertia[b], [4.369e-5, 4.028e-5, 1.407e-5])


def test_equality_compile_facts(capi, load_model):
    # reference ros_interface_test.cpp:769-829
    m = load_model("equality_scene.xml")
    assert m.neq == 4
    w = m.name2id(capi.OBJ_EQUALITY, "weld_eq")
    assert m.eq_type[w] == 1 and m.eq_obj2id[w] == 0 and m.eq_active[w] == 1
    np.testing.assert_allclose(m.eq_solref[w], [0.3, 0.9])
    np.testing.assert_allclose(m.eq_solimp[w], [0.8, 0.95, 0.002, 0.4, 2])
    d = m.eq_data[w]
    # anchor (3) | relpose pos (3) | relpose quat normalised (4) | torquescale
    np.testing.assert_allclose(d[3:6], [1.1, 1.2, 1.3])
    q = np.array([0.358, -0.003, -0.886, 0.295])
    np.testing.assert_allclose(d[6:10], q / np.linalg.norm(q), atol=1e-12)
    assert d[10] == pytest.approx(0.9)
    j = m.name2id(capi.OBJ_EQUALITY, "joint_eq")
    assert m.eq_type[j] == 2
    np.testing.assert_allclose(m.eq_data[j][:5], [0.5, 0.25, 0.76, 0.66, 1])
    t = m.name2id(capi.OBJ_EQUALITY, "tendon_eq")
    assert m.eq_type[t] == 3
    c = m.name2id(capi.OBJ_EQUALITY, "connect_eq")
    assert m.eq_type[c] == 0 and m.eq_obj2id[c] == 0
    np.testing.assert_allclose(m.eq_solimp[c], [0.8, 0.95, 0.002, 0.4, 1])


def test_name_lookup_roundtrip(capi, load_model):
    m = load_model("panda_like.xml")
    for j in range(m.njnt):
        name = m.id2name(capi.OBJ_JOINT, j)
        assert name and m.name2id(capi.OBJ_JOINT, name) == j
    assert m.name2id(capi.OBJ_JOINT, "no_such_joint") == -1


def test_panda_like_shape(load_model):
    m = load_model("panda_like.xml")
    assert (m.nq, m.nv, m.nu, m.nbody) == (9, 9, 8, 12)
    assert m.opt.solver == 0 and m.opt.integrator == 0 and m.opt.cone == 0 and m.opt.timestep == 0.002
    # every candidate pair references valid geoms, ordered, no duplicates
    pairs = list(zip(m.collpair_geom1.tolist(), m.collpair_geom2.tolist()))
    assert len(set(pairs)) == len(pairs) == m.ncollpair


def test_set_const_is_idempotent(load_model, capi):
    from conftest import model_path
    m = capi.Model.from_xml_file(model_path("panda_like.xml"))
    a = {k: np.array(getattr(m, k)) for k in ("dof_invweight0", "body_invweight0", "body_subtreemass", "actuator_acc0")}
    mi = m.stat.meaninertia
    m.set_const()
    for k, v in a.items():
        np.testing.assert_allclose(getattr(m, k), v, rtol=1e-13, atol=0)
    assert m.stat.meaninertia == pytest.approx(mi, rel=1e-13)


def test_unsupported_physics_attributes_are_refused_not_ignored(capi):
    """Attributes that would change the dynamics but are not implemented must fail the compile instead of being dropped
    (round-1 verdict: "silently ignored" options): the ellipsoid fluid model, shell inertia, tendon armature."""
    base = ('<mujoco><worldbody><body pos="0 0 1"><joint name="j" type="hinge" %s/><geom size="0.1" %s/></body></worldbody>'
            '%s</mujoco>')
    for joint, geom, extra, word in (("", 'shellinertia="true"', "", "shellinertia"),
                                     ("", 'fluidshape="ellipsoid"', "", "fluidshape"),
                                     ("", 'fluidcoef="0.5 0.25 1.5 1 1"', "", "fluidcoef"),
                                     ('actuatorfrcrange="-1 1"', "", "", "actuatorfrcrange"),
                                     ("", "", '<tendon><fixed armature="1"><joint joint="j" coef="1"/></fixed></tendon>',
                                      "armature")):
        with pytest.raises(capi.B2mjError, match=word):
            capi.Model.from_xml_string(base % (joint, geom, extra))
    capi.Model.from_xml_string(base % ("", 'fluidshape="none"', ""))  # the default value is fine
    # enable flags that change the dynamics or the sensor values are refused when switched on; energy / fwdinv only add
    # outputs and are accepted; a misspelt flag is an error
    flagged = '<mujoco><option><flag %s/></option><worldbody><body><freejoint/><geom size="0.1"/></body></worldbody></mujoco>'
    for flag in ("sensornoise", "multiccd"):
        with pytest.raises(capi.B2mjError, match=flag):
            capi.Model.from_xml_string(flagged % f'{flag}="enable"')
        capi.Model.from_xml_string(flagged % f'{flag}="disable"')
    capi.Model.from_xml_string(flagged % 'energy="enable" fwdinv="enable"')
    with pytest.raises(capi.B2mjError, match="unknown option flag"):
        capi.Model.from_xml_string(flagged % 'gravty="disable"')


def test_joint_springdamper_sets_stiffness_and_damping_from_the_effective_inertia(capi, orc):
    """springdamper="tau zeta": k = I / (tau zeta)^2, b = 2 I / tau with I = ndof / sum(dof_invweight0) at qpos0
    (mjCModel::AutoSpringDamper), overriding the joint's own stiffness / damping.  Checked on a pendulum whose effective
    inertia is known in closed form, and dynamically: the oracle's free response decays with time constant tau."""
    tau, zeta = 0.25, 0.5
    xml = ('<mujoco><option gravity="0 0 0" timestep="0.0005"/><worldbody><body><joint name="j" axis="0 1 0" stiffness="3" '
           f'damping="7" springdamper="{tau} {zeta}"/><geom size="0.05" pos="0 0 -0.4" mass="2"/></body>'
           '<body pos="1 0 0"><freejoint/><geom size="0.1" mass="3"/></body></worldbody></mujoco>')
    m = capi.Model.from_xml_string(xml)
    inertia = 2 * 0.4 ** 2 + 0.4 * 2 * 0.05 ** 2  # point mass at 0.4 m + the sphere's own 2/5 m r^2
    np.testing.assert_allclose(1 / m.dof_invweight0[0], inertia, rtol=1e-12)
    np.testing.assert_allclose(m.jnt_stiffness[0], inertia / (tau * zeta) ** 2, rtol=1e-12)
    np.testing.assert_allclose(m.dof_damping[0], 2 * inertia / tau, rtol=1e-12)
    assert m.jnt_stiffness[1] == 0 and not m.dof_damping[1:].any()  # joints without the attribute are untouched
    # free joint: one damping value on all six dofs from the averaged weight
    mf = capi.Model.from_xml_string('<mujoco><worldbody><body><joint type="free" springdamper="0.1 1"/><geom size="0.1" mass="3"/>'
                                    '</body></worldbody></mujoco>')
    avg = 6 / mf.dof_invweight0[:6].sum()
    np.testing.assert_allclose(mf.dof_damping[:6], 2 * avg / 0.1, rtol=1e-12)
    np.testing.assert_allclose(mf.jnt_stiffness[0], avg / 0.01, rtol=1e-12)
    # free response from q0: envelope exp(-t / tau), so the amplitude one time constant later is below q0 / e
    o = orc.Oracle(m)
    q = o.get("qpos").copy()
    q[0] = 0.2
    o.set("qpos", q)
    n = int(round(tau / 0.0005))
    peak = 0.0
    for k in range(3 * n):
        o.step()
        if k >= n:
            peak = max(peak, abs(o.get("qpos")[0]))
    assert 0.2 * np.exp(-3) < peak < 0.2 * np.exp(-1) * 1.3


def test_actuator_shortcuts_equal_their_general_form(capi):
    """<intvelocity>, <damper>, <cylinder> are shorthands for <general> (MuJoCo XML reference, actuator section): every
    actuator array must equal the spelled-out form, which the GPU path tests (affine gain / bias, integrator, filter)."""
    body = ('<mujoco><worldbody><body><joint name="j" axis="0 1 0"/><geom size="0.1"/></body></worldbody><actuator>%s'
            '</actuator></mujoco>')
    pairs = [
        ('<intvelocity joint="j" kp="30" actrange="-1 1"/>',
         '<general joint="j" dyntype="integrator" gainprm="30" biastype="affine" biasprm="0 -30 0" actrange="-1 1"/>'),
        ('<damper joint="j" kv="4" ctrlrange="0 2"/>',
         '<general joint="j" gaintype="affine" gainprm="0 0 -4" ctrlrange="0 2"/>'),
        ('<cylinder joint="j" timeconst="0.3" diameter="0.2" bias="1 -2 -0.5"/>',
         f'<general joint="j" dyntype="filter" dynprm="0.3" gainprm="{np.pi / 4 * 0.04!r}" biastype="affine" biasprm="1 -2 -0.5"/>'),
        ('<cylinder joint="j" area="0.7"/>',
         '<general joint="j" dyntype="filter" dynprm="1" gainprm="0.7" biastype="affine"/>'),
    ]
    fields = ("actuator_dyntype", "actuator_gaintype", "actuator_biastype", "actuator_dynprm", "actuator_gainprm",
              "actuator_biasprm", "actuator_ctrllimited", "actuator_ctrlrange", "actuator_actlimited", "actuator_actrange",
              "actuator_actadr", "actuator_trnid", "actuator_gear")
    for short, general in pairs:
        a, b = capi.Model.from_xml_string(body % short), capi.Model.from_xml_string(body % general)
        assert a.na == b.na
        for f in fields:
            np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f"{short}: {f}")
    for bad, word in (('<damper joint="j" kv="4"/>', "control range"), ('<damper joint="j" kv="4" ctrlrange="-1 1"/>', "negative"),
                      ('<damper joint="j" kv="-1" ctrlrange="0 1"/>', "negative"), ('<muscle joint="j"/>', "muscle")):
        with pytest.raises(capi.B2mjError, match=word):
            capi.Model.from_xml_string(body % bad)


def test_compiler_mass_options_and_statistic_overrides(capi):
    """compiler settotalmass / boundmass / boundinertia / inertiagrouprange and <statistic> overrides change the
    dynamics (masses, inertias, the solver's meaninertia scale) and were dropped silently before round 2c."""
    body = ('<worldbody><body pos="0 0 1"><joint type="hinge"/><geom size="0.1" group="2"/><geom size="0.05" pos="0.2 0 0"/>'
            '<body pos="0 0 2"><joint type="hinge"/><geom size="0.1"/></body></body></worldbody>')
    mk = lambda c, extra="": capi.Model.from_xml_string(f"<mujoco><compiler {c}/>{extra}{body}</mujoco>")  # noqa: E731
    m0 = mk("")
    v1, v2 = 4000 / 3 * np.pi * 0.1 ** 3, 4000 / 3 * np.pi * 0.05 ** 3  # density 1000
    np.testing.assert_allclose(m0.body_mass, [0, v1 + v2, v1], rtol=1e-12)
    m = mk('settotalmass="5"')
    np.testing.assert_allclose(m.body_mass.sum(), 5, rtol=1e-12)
    np.testing.assert_allclose(m.body_mass[1:] / m0.body_mass[1:], 5 / m0.body_mass.sum(), rtol=1e-12)
    np.testing.assert_allclose(m.body_inertia[1:] / m0.body_inertia[1:], 5 / m0.body_mass.sum(), rtol=1e-12)
    m = mk('inertiagrouprange="0 1"')  # the group-2 sphere no longer counts: the body is the small sphere alone
    np.testing.assert_allclose(m.body_mass[1], v2, rtol=1e-12)
    np.testing.assert_allclose(m.body_ipos[1], [0.2, 0, 0], atol=1e-15)
    m = mk('boundmass="6" boundinertia="0.5"')
    np.testing.assert_allclose(m.body_mass, [0, 6, 6])
    np.testing.assert_allclose(m.body_inertia[1:], 0.5)
    assert m.stat.meaninertia > m0.stat.meaninertia
    m = mk("", '<statistic meaninertia="2.5" extent="3" center="0 0 1"/>')
    assert (m.stat.meaninertia, m.stat.extent, list(m.stat.center)) == (2.5, 3.0, [0, 0, 1])


def test_override_flag_rewrites_every_contact_parameter(capi, orc):
    """<flag override="enable">: every contact takes o_margin (gap 0), o_solref, o_solimp -- applied by the compiler to
    all geoms and explicit pairs (the mixing rules reproduce a shared value), so the step needs no run-time switch."""
    xml = ('<mujoco><option o_margin="0.01" o_solref="0.05 0.8" o_solimp="0.8 0.9 0.002"><flag override="%s"/></option>'
           '<worldbody><geom name="g1" type="plane" size="1 1 .1" solref="0.01 1" margin="0.003"/><body pos="0 0 0.105"><freejoint/>'
           '<geom name="g2" size="0.1" solref="0.02 1.2"/></body><body pos="1 0 0.105"><freejoint/><geom size="0.1" gap="0.001"/>'
           '</body></worldbody><contact><pair geom1="g1" geom2="g2" solref="0.03 1" margin="0.2"/></contact></mujoco>')
    on, off = capi.Model.from_xml_string(xml % "enable"), capi.Model.from_xml_string(xml % "disable")
    assert (on.opt.enableflags, off.opt.enableflags) == (1, 0)
    o = orc.Oracle(on)
    o.forward()
    n = int(o.get("ncon")[0])
    assert n == 2  # the explicit pair and the dynamic plane-sphere pair
    np.testing.assert_allclose(o.get("contact_solref")[:2 * n].reshape(n, 2), np.tile([0.05, 0.8], (n, 1)))
    np.testing.assert_allclose(o.get("contact_solimp")[:5 * n].reshape(n, 5), np.tile([0.8, 0.9, 0.002, 0.5, 2], (n, 1)))
    np.testing.assert_allclose(o.get("contact_includemargin")[:n], 0.01)
    o2 = orc.Oracle(off)
    o2.forward()
    np.testing.assert_allclose(o2.get("contact_solref")[:2], [0.03, 1])
    np.testing.assert_allclose(o2.get("contact_includemargin")[:1], 0.2)


def test_frame_elements_are_pure_coordinate_transforms(capi):
    """<frame> (MuJoCo 3 model files): everything inside takes the frame's pose -- geoms, sites, joint anchors and axes,
    child bodies, nested frames, fromto geoms -- and its childclass; the frame itself leaves nothing in the model.  The
    expected model is the same scene with every pose composed here with numpy and written out flat."""
    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])

    def rot(q, v):
        return qmul(qmul(q, np.array([0, *v])), q * [1, -1, -1, -1])[1:]

    def compose(F, pos, quat=(1, 0, 0, 0)):
        return F[0] + rot(F[1], np.asarray(pos, float)), qmul(F[1], np.asarray(quat, float))

    def fmt(v):
        return " ".join(repr(float(x)) for x in v)
    s, c = np.sin(np.pi / 4), np.cos(np.pi / 4)
    F1 = (np.array([0, 0, 1.0]), np.array([c, 0, 0, s]))          # pos 0 0 1, 90 deg about z
    F2 = (np.array([0, .5, 0]), np.array([c, s, 0, 0]))           # inside body a: 90 deg about x
    F3 = compose(F2, [.1, .1, .1])                                # nested in F2, no rotation of its own
    qa = np.array([np.cos(np.pi / 12), 0, np.sin(np.pi / 12), 0])  # body a: 30 deg about y
    framed = f"""<mujoco><default><default class="c"><geom size="0.03" friction="0.5 0.01 0.001"/></default></default><worldbody>
      <frame pos="0 0 1" quat="{fmt(F1[1])}">
        <geom name="g0" type="box" size=".1 .2 .3" pos="1 0 0"/>
        <body name="a" pos="1 0 0" quat="{fmt(qa)}">
          <joint name="ja" axis="1 0 0"/><geom name="ga" size="0.1"/>
          <frame pos="0 .5 0" quat="{fmt(F2[1])}" childclass="c">
            <joint name="jb" type="slide" axis="0 0 1" pos="0 0 .1"/>
            <geom name="gb" type="capsule" fromto="0 0 0 0 0 .2"/>
            <site name="sb" pos=".1 0 0" quat="{fmt(qa)}"/>
            <frame pos=".1 .1 .1"><body name="b" pos="0 0 .2"><joint name="jc" axis="0 1 0"/><geom name="gc" size=".05"/></body></frame>
          </frame>
          <body name="c" pos="0 0 -.3"><joint name="jd" axis="0 0 1"/><geom name="gd" size=".04"/></body>
        </body>
      </frame></worldbody></mujoco>"""
    g0, a = compose(F1, [1, 0, 0]), compose(F1, [1, 0, 0], qa)
    sb, b = compose(F2, [.1, 0, 0], qa), compose(F3, [0, 0, .2])
    ft = np.concatenate([compose(F2, [0, 0, 0])[0], compose(F2, [0, 0, .2])[0]])
    flat = f"""<mujoco><worldbody>
      <geom name="g0" type="box" size=".1 .2 .3" pos="{fmt(g0[0])}" quat="{fmt(g0[1])}"/>
      <body name="a" pos="{fmt(a[0])}" quat="{fmt(a[1])}">
        <joint name="ja" axis="1 0 0"/><geom name="ga" size="0.1"/>
        <joint name="jb" type="slide" axis="{fmt(rot(F2[1], [0, 0, 1]))}" pos="{fmt(compose(F2, [0, 0, .1])[0])}"/>
        <geom name="gb" type="capsule" fromto="{fmt(ft)}" size="0.03" friction="0.5 0.01 0.001"/>
        <site name="sb" pos="{fmt(sb[0])}" quat="{fmt(sb[1])}"/>
        <body name="b" pos="{fmt(b[0])}" quat="{fmt(b[1])}"><joint name="jc" axis="0 1 0"/><geom name="gc" size=".05" friction="0.5 0.01 0.001"/></body>
        <body name="c" pos="0 0 -.3"><joint name="jd" axis="0 0 1"/><geom name="gd" size=".04"/></body>
      </body></worldbody></mujoco>"""
    mf, me = capi.Model.from_xml_string(framed), capi.Model.from_xml_string(flat)
    assert (mf.nbody, mf.njnt, mf.ngeom, mf.nsite) == (me.nbody, me.njnt, me.ngeom, me.nsite) == (4, 4, 5, 1)
    for f in ("body_parentid", "jnt_bodyid", "jnt_type", "geom_bodyid", "geom_type", "site_bodyid"):
        np.testing.assert_array_equal(getattr(mf, f), getattr(me, f), err_msg=f)
    for f in ("body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "jnt_pos", "jnt_axis", "geom_pos",
              "geom_size", "geom_friction", "site_pos", "site_quat", "dof_invweight0", "body_invweight0"):
        np.testing.assert_allclose(getattr(mf, f), getattr(me, f), rtol=1e-12, atol=1e-14, err_msg=f)
    # q and -q are the same rotation (the fromto capsule comes out with the other sign)
    np.testing.assert_allclose(np.abs(np.sum(mf.geom_quat.reshape(-1, 4) * me.geom_quat.reshape(-1, 4), axis=1)), 1, atol=1e-14)
    assert [mf.name2id(capi.OBJ_BODY, n) for n in "abc"] == [1, 2, 3]  # document order, depth first
    for bad, word in (('<frame><inertial pos="0 0 0" mass="1" diaginertia="1 1 1"/></frame>', "inertial"),
                      ('<replicate count="3"><geom size=".1"/></replicate>', "replicate")):
        with pytest.raises(capi.B2mjError, match=word):
            capi.Model.from_xml_string(f'<mujoco><worldbody><body>{bad}</body></worldbody></mujoco>')


def test_site_fromto_matches_the_geom_rule(capi):
    """<site fromto="..."> (capsule / cylinder / box / ellipsoid sites): midpoint, z axis along the segment and half
    length, exactly as for a geom with the same attribute."""
    x = ('<mujoco><worldbody><body><geom type="{0}" fromto="0.1 0 0 0.1 0.2 0.3" size="0.02"/>'
         '<site type="{0}" fromto="0.1 0 0 0.1 0.2 0.3" size="0.02"/></body></worldbody></mujoco>')
    for kind in ("capsule", "cylinder", "box", "ellipsoid"):
        m = capi.Model.from_xml_string(x.format(kind))
        np.testing.assert_array_equal(m.site_pos, m.geom_pos)
        np.testing.assert_array_equal(m.site_quat, m.geom_quat)
        n = 2 if kind in ("capsule", "cylinder") else 3
        np.testing.assert_array_equal(m.site_size.ravel()[:n], m.geom_size.ravel()[:n])
    with pytest.raises(capi.B2mjError, match="fromto"):
        capi.Model.from_xml_string(x.format("sphere"))


def test_unknown_attributes_on_physics_elements_are_errors(capi):
    """MuJoCo validates elements against its schema; a misspelt attribute must not be dropped silently (the model would
    simulate something else).  Checked for body / inertial / joint / freejoint / geom / site / frame."""
    base = '<mujoco><worldbody><body {0}><inertial pos="0 0 0" mass="1" diaginertia="1 1 1" {1}/><joint {2}/><geom size="0.1" {3}/><site {4}/></body></worldbody></mujoco>'
    capi.Model.from_xml_string(base.format("", "", "", "", ""))
    for k, (attr, where) in enumerate((('gravcmp="1"', "body"), ('fulinertia="1 1 1 0 0 0"', "inertial"), ('dampng="2"', "joint"),
                                       ('fricton="1 0 0"', "geom"), ('fromt="0 0 0 1 0 0"', "site"))):
        args = [""] * 5
        args[k] = attr
        with pytest.raises(capi.B2mjError, match=f"<{where}>: unknown attribute '{attr.split('=')[0]}'"):
            capi.Model.from_xml_string(base.format(*args))
    with pytest.raises(capi.B2mjError, match="unknown attribute 'damping'"):
        capi.Model.from_xml_string('<mujoco><worldbody><body><freejoint damping="1"/><geom size=".1"/></body></worldbody></mujoco>')
    # every schema attribute that only matters to rendering or bookkeeping is accepted
    capi.Model.from_xml_string('<mujoco><worldbody><body name="b" user="1 2"><joint group="2" user="3"/><geom size="0.1" material="m" '
                               'rgba="1 0 0 1" user="1"/><site material="m" rgba="0 1 0 1" group="3" user="4"/></body></worldbody></mujoco>')


def test_primitive_fitted_to_a_mesh_is_refused(capi):
    """<geom type="box" mesh="..."> asks MuJoCo to fit the primitive to the mesh (size is then ignored): not implemented,
    so it must not compile to a box of the given size."""
    xml = ('<mujoco><asset><mesh name="m" vertex="0 0 0 1 0 0 0 1 0 0 0 1"/></asset><worldbody><body>'
           '<geom type="%s" mesh="m" size=".1 .1 .1"/></body></worldbody></mujoco>')
    capi.Model.from_xml_string(xml % "mesh")
    for kind in ("box", "sphere", "capsule"):
        with pytest.raises(capi.B2mjError, match="fitting a primitive"):
            capi.Model.from_xml_string(xml % kind)
