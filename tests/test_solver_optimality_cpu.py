"""Algorithm-independent check of the oracle's constraint stage: whatever solver ran, the
answer has to be the minimiser of MuJoCo's convex problem  1/2 (a - a_smooth)' M (a - a_smooth) + s(J a - aref).
Its optimality conditions need no solver code to state:
    stationarity   M (qacc - qacc_smooth) = J' f
    row laws       equality:       f = -D r                      (r = J qacc - aref)
                   friction loss:  f = clamp(-D r, -loss, +loss)
                   limit, frictionless and pyramidal contact rows:  f = max(0, -D r)
                   elliptic contacts: f inside the friction cone, f = 0 when the contact separates (top zone),
                                      f = -D r when r lies in the polar cone (bottom zone)
They are evaluated here with numpy from the fields the oracle exposes (efc_J, efc_D, efc_aref, efc_force, qM, ...), on
the bench models in their contact-rich states.  The GPU path is held to the oracle per field; this file holds the oracle
to the problem statement."""
import numpy as np
import pytest

from conftest import model_path

EQUALITY, FRICTION_DOF, FRICTION_TENDON, LIMIT_JOINT, LIMIT_TENDON, FRICTIONLESS, PYRAMIDAL, ELLIPTIC = range(8)
PGS, CG, NEWTON = 0, 1, 2


def dense_mass(m, qM):
    M = np.zeros((m.nv, m.nv))
    for i in range(m.nv):
        adr, j, k = m.dof_Madr[i], i, 0
        while j >= 0:
            M[i, j] = M[j, i] = qM[adr + k]
            j = m.dof_parentid[j]
            k += 1
    return M


CASES = [("panda_like.xml", 0, 450, NEWTON), ("panda_like.xml", 1, 450, NEWTON), ("humanoid_like.xml", 0, 150, NEWTON),
         ("humanoid_like.xml", 1, 150, CG), ("box_stack.xml", 0, 150, NEWTON), ("box_stack.xml", 1, 150, NEWTON),
         ("equality_scene.xml", 0, 60, NEWTON), ("hand_like.xml", 1, 120, NEWTON), ("ROWS", 0, 400, NEWTON), ("ROWS", 0, 400, CG),
         ("CONDIM", 0, 200, NEWTON), ("CONDIM", 1, 200, NEWTON), ("CONDIM", 1, 200, CG)]

# every remaining row type in one scene: dof and tendon friction loss, joint and tendon limits, a frictionless contact
ROWS = """<mujoco><option timestep="0.002"/><worldbody>
  <geom type="plane" size="2 2 .1" condim="1"/>
  <body pos="0 0 1"><joint name="a" axis="0 1 0" range="-0.4 0.4" limited="true" frictionloss="0.3"/>
    <geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.03"/>
    <body pos="0.4 0 0"><joint name="b" axis="0 1 0" frictionloss="0.05"/><geom type="capsule" fromto="0 0 0 0.3 0 0" size="0.03"/></body></body>
  <body pos="1 0 0.099"><freejoint/><geom size="0.1" condim="1"/></body>
</worldbody>
<tendon><fixed name="t" limited="true" range="-0.3 0.5" frictionloss="0.1"><joint joint="a" coef="1"/><joint joint="b" coef="1"/></fixed></tendon>
</mujoco>"""


# contact dimensions 1 / 4 / 6 (torsional and rolling friction rows), margin + gap, solmix / priority mixing, spinning and
# rolling bodies so that the extra friction rows carry force
CONDIM = """<mujoco><option timestep="0.002" impratio="3"/><worldbody>
  <geom name="floor" type="plane" size="3 3 .1" condim="3" friction="0.9 0.02 0.01" margin="0.004" gap="0.001" solmix="2"/>
  <body pos="0 0 0.1"><freejoint/><geom size="0.1" condim="4" friction="0.6 0.03 0.002" priority="1"/></body>
  <body pos="0.5 0 0.1"><freejoint/><geom size="0.1" condim="6" friction="0.7 0.01 0.004" solref="0.01 0.8" solmix="0.5"/></body>
  <geom type="box" size="0.3 0.3 0.1" pos="1 0 0.1" condim="1"/>
  <body pos="1 0 0.25"><freejoint/><geom type="box" size="0.08 0.06 0.05" condim="1"/></body>
  <body pos="1.5 0 0.05"><freejoint/><geom type="capsule" size="0.05 0.1" euler="90 0 0" condim="6" margin="0.01" gap="0.003"/></body>
  <body pos="0.5 0.5 0.32"><freejoint/><geom size="0.08" condim="4"/></body>
  <body pos="0.5 0.5 0.1"><freejoint/><geom size="0.12" condim="6" friction="0.5 0.02 0.003"/></body>
</worldbody></mujoco>"""


def load_case(name, capi):
    if name == "ROWS":
        return capi.Model.from_xml_string(ROWS), None
    if name == "ROWS_DIRECT":  # negative solref = stiffness and damping given directly; a power-3 impedance curve
        return capi.Model.from_xml_string(ROWS.replace('frictionloss="0.3"', 'frictionloss="0.3" solreflimit="-800 -30" '
                                                       'solimplimit="0.8 0.97 0.01 0.3 3"')), None
    if name == "CONDIM":
        m = capi.Model.from_xml_string(CONDIM)
        v = np.zeros(m.nv)
        for b in range(6):  # every body slides, spins about the vertical and rolls
            v[6 * b:6 * b + 6] = [0.3, -0.2, 0, 1.0, -2.0, 6.0]
        return m, v
    return capi.Model.from_xml_file(model_path(name)), None  # options are edited: not the shared cached model


@pytest.mark.parametrize("name,cone,settle,solver", CASES)
def test_solution_satisfies_the_optimality_conditions(name, cone, settle, solver, orc, capi):
    m, v0 = load_case(name, capi)
    m.opt.cone, m.opt.solver = cone, solver
    m.opt.tolerance, m.opt.iterations = 1e-14, 200
    o = orc.Oracle(m)
    rng = np.random.default_rng(4)
    o.set("qpos", m.qpos0 + rng.uniform(-0.05, 0.05, m.nq) * (v0 is None))
    if v0 is not None:
        o.set("qvel", v0)
    for s in range(settle):
        if m.nu and s % 25 == 0:
            lo, hi = m.actuator_ctrlrange.reshape(-1, 2).T
            o.set("ctrl", np.where(hi > lo, rng.uniform(np.minimum(lo, hi), np.maximum(lo, hi)), rng.uniform(-1, 1, m.nu)))
        o.step(1)
    o.forward()
    nefc, nv = int(o.get("nefc")[0]), m.nv
    assert nefc > 0
    J = o.get("efc_J")[:nefc * nv].reshape(nefc, nv)
    D, aref, f = o.get("efc_D")[:nefc], o.get("efc_aref")[:nefc], o.get("efc_force")[:nefc]
    typ, ids, loss = o.get("efc_type")[:nefc], o.get("efc_id")[:nefc], o.get("efc_frictionloss")[:nefc]
    qacc, M = o.get("qacc"), dense_mass(m, o.get("qM"))
    scale = max(1.0, np.abs(f).max())
    eps = 2e-7 if solver == NEWTON else 1e-5  # CG stops at a looser point than Newton's quadratic convergence reaches
    # stationarity, and the solver's own report of it
    np.testing.assert_allclose(M @ (qacc - o.get("qacc_smooth")), J.T @ f, atol=eps * scale * max(1, np.abs(J).max()))
    np.testing.assert_allclose(o.get("qfrc_constraint"), J.T @ f, atol=1e-9 * scale * max(1, np.abs(J).max()))
    r = J @ qacc - aref
    tol = eps * scale
    seen = set()
    i = 0
    while i < nefc:
        t = int(typ[i])
        seen.add(t)
        if t == EQUALITY:
            assert abs(f[i] + D[i] * r[i]) < tol, (i, "equality")
        elif t in (FRICTION_DOF, FRICTION_TENDON):
            assert abs(f[i] - np.clip(-D[i] * r[i], -loss[i], loss[i])) < tol, (i, "friction loss")
        elif t in (LIMIT_JOINT, LIMIT_TENDON, FRICTIONLESS, PYRAMIDAL):
            assert abs(f[i] - max(0.0, -D[i] * r[i])) < tol, (i, "one-sided row", t)
        else:  # elliptic contact: rows i .. i + dim - 1 belong to contact ids[i]
            c = int(ids[i])
            dim = int(o.get("contact_dim")[c])
            mu = o.get("contact_friction")[5 * c:5 * c + 5][:dim - 1]
            fn, ft = f[i], f[i + 1:i + dim]
            assert fn > -tol and np.sqrt(np.sum((ft / mu) ** 2)) <= fn + tol, (i, "force outside the friction cone", fn, ft)
            rn, rt = r[i], r[i + 1:i + dim]
            # zones of MuJoCo's elliptic cost in the scaled coordinates  N = mu0 r_n,  T = |mu_j r_j| * (per-row scale);
            # with the regularised friction mu0 the test below is the one that needs no knowledge of that scaling:
            # a separating contact whose tangential motion is small carries no force at all
            if rn > 0 and np.all(np.abs(rt) < 1e-12):
                assert np.abs(f[i:i + dim]).max() < tol, (i, "top zone")
            i += dim
            continue
        i += 1
    assert seen, name
    if name == "ROWS":
        assert seen >= {FRICTION_DOF, FRICTION_TENDON, LIMIT_JOINT, FRICTIONLESS}, seen
    if name == "CONDIM":
        dims = set(int(d) for d in o.get("contact_dim")[:int(o.get("ncon")[0])])
        assert dims >= {1, 4, 6} and FRICTIONLESS in seen, (dims, seen)
        assert np.abs(f).max() > 1 and nefc >= (30 if cone == 0 else 18), nefc


def test_the_conditions_reject_a_wrong_answer(capi, orc):
    """The checker itself: an under-converged PGS answer (2 sweeps) violates stationarity or a row law by far more than
    the tolerance the test above applies."""
    m = capi.Model.from_xml_file(model_path("box_stack.xml"))
    m.opt.solver, m.opt.iterations = PGS, 2
    o = orc.Oracle(m)
    o.step(150)
    o.set("qacc_warmstart", np.zeros(m.nv))
    o.forward()
    nefc, nv = int(o.get("nefc")[0]), m.nv
    J = o.get("efc_J")[:nefc * nv].reshape(nefc, nv)
    D, aref, f = o.get("efc_D")[:nefc], o.get("efc_aref")[:nefc], o.get("efc_force")[:nefc]
    r = J @ o.get("qacc") - aref
    worst = np.abs(f - np.maximum(0.0, -D * r)).max()
    assert worst > 1e-3 * max(1.0, np.abs(f).max())


def impedance(solimp, x):
    """MuJoCo's documented impedance curve d(r) (XML reference, solimp): x = pos - margin."""
    d0, dw = np.clip(solimp[:2], 1e-4, 0.9999)
    width, mid, power = max(0.0, solimp[2]), np.clip(solimp[3], 1e-4, 0.9999), max(1.0, solimp[4])
    if d0 == dw or width <= 1e-15:
        return 0.5 * (d0 + dw)
    x = abs(x) / width
    if x >= 1 or x <= 0:
        return dw if x >= 1 else d0
    if power == 1:
        y = x
    elif x <= mid:
        y = x ** power / mid ** (power - 1)
    else:
        y = 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
    return d0 + y * (dw - d0)


def stiffness_damping(solref, solimp, dt):
    dmax = np.clip(solimp[1], 1e-4, 0.9999)
    if solref[0] > 0 and solref[1] > 0:
        tc = max(solref[0], 2 * dt)  # refsafe
        return 1 / (dmax * dmax * tc * tc * solref[1] * solref[1]), 2 / (dmax * tc)
    return -solref[0] / (dmax * dmax), -solref[1] / dmax


@pytest.mark.parametrize("name,cone,settle", [("panda_like.xml", 0, 450), ("humanoid_like.xml", 1, 150), ("box_stack.xml", 0, 150),
                                               ("equality_scene.xml", 0, 60), ("ROWS", 0, 400), ("ROWS_DIRECT", 0, 400),
                                               ("CONDIM", 1, 200)])
def test_constraint_parameters_follow_the_documented_formulas(name, cone, settle, orc, capi):
    """Row M6 against MuJoCo's published definitions (Computation chapter, "solver parameters"): reference acceleration
    aref = -B vel - K d (pos - margin), regulariser R = (1 - d) / d * diagApprox, D = 1 / R, with (K, B) from solref
    (time constant clamped at two time steps) and the impedance d from the solimp curve -- recomputed here with numpy
    from the per-row inputs the oracle exposes."""
    m, v0 = load_case(name, capi)
    m.opt.cone = cone
    o = orc.Oracle(m)
    rng = np.random.default_rng(9)
    o.set("qpos", m.qpos0 + rng.uniform(-0.05, 0.05, m.nq) * (v0 is None))
    if v0 is not None:
        o.set("qvel", v0)
    o.step(settle)
    o.forward()
    nefc = int(o.get("nefc")[0])
    typ, ids = o.get("efc_type")[:nefc], o.get("efc_id")[:nefc]
    pos, margin, vel = o.get("efc_pos")[:nefc], o.get("efc_margin")[:nefc], o.get("efc_vel")[:nefc]
    kbip = o.get("efc_KBIP")[:4 * nefc].reshape(nefc, 4)
    aref, R, D, diag = (o.get(k)[:nefc] for k in ("efc_aref", "efc_R", "efc_D", "efc_diagApprox"))
    np.testing.assert_allclose(aref, -kbip[:, 1] * vel - kbip[:, 0] * kbip[:, 2] * (pos - margin), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(D * R, 1, rtol=1e-12)
    dt, checked, first_of_contact = m.opt.timestep, 0, {}
    for i in range(nefc):
        t, k = int(typ[i]), int(ids[i])
        if t == LIMIT_JOINT:
            solref, solimp = m.jnt_solref.reshape(-1, 2)[k], m.jnt_solimp.reshape(-1, 5)[k]
        elif t == EQUALITY:
            solref, solimp = m.eq_solref.reshape(-1, 2)[k], m.eq_solimp.reshape(-1, 5)[k]
        elif t in (FRICTIONLESS, PYRAMIDAL, ELLIPTIC):
            if t == ELLIPTIC and first_of_contact.setdefault(k, i) != i:
                continue  # tangential rows of an elliptic contact: scaled from the normal row, not from this formula
            solref, solimp = o.get("contact_solref")[2 * k:2 * k + 2], o.get("contact_solimp")[5 * k:5 * k + 5]
        else:
            continue
        d = impedance(solimp, pos[i] - margin[i])
        K, B = stiffness_damping(solref, solimp, dt)
        np.testing.assert_allclose(kbip[i, :3], [K, B, d], rtol=1e-12, err_msg=f"row {i} type {t}")
        if t != PYRAMIDAL:  # pyramid edges carry an extra friction-dependent factor
            np.testing.assert_allclose(R[i], max(1e-15, (1 - d) / d * diag[i]), rtol=1e-12, err_msg=f"row {i} type {t}")
        checked += 1
    assert checked >= 2, (name, checked)


def integrate_pos(m, q, v, eps):
    """qpos advanced along the generalised velocity v (mj_integratePos: free-joint linear velocity in the world frame,
    angular velocities of free and ball joints in the body frame, applied on the right of the quaternion)."""
    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])

    def rotate(quat, w):
        ang = np.linalg.norm(w) * eps
        if abs(ang) < 1e-300:
            return quat
        out = qmul(quat, np.array([np.cos(ang / 2), *(np.sin(ang / 2) * w / np.linalg.norm(w))]))
        return out / np.linalg.norm(out)
    q = q.copy()
    for j in range(m.njnt):
        qa, da, t = m.jnt_qposadr[j], m.jnt_dofadr[j], m.jnt_type[j]
        if t == 0:  # free
            q[qa:qa + 3] += eps * v[da:da + 3]
            q[qa + 3:qa + 7] = rotate(q[qa + 3:qa + 7], v[da + 3:da + 6])
        elif t == 1:  # ball
            q[qa:qa + 4] = rotate(q[qa:qa + 4], v[da:da + 3])
        else:
            q[qa] += eps * v[da]
    return q


@pytest.mark.parametrize("name,cone,settle", [("ROWS", 1, 400), ("humanoid_like.xml", 1, 150), ("hand_like.xml", 1, 120),
                                               ("equality_scene.xml", 0, 60), ("panda_like.xml", 1, 450),
                                               ("humanoid_like.xml", 0, 150), ("panda_like.xml", 0, 450)])
def test_constraint_jacobian_is_the_derivative_of_the_residual(name, cone, settle, orc, capi):
    """efc_J against central differences of efc_pos along random generalised velocities: joint limits, connect / weld /
    joint / tendon equalities, frictionless contacts and the normal rows of elliptic contacts (distance of the nearest points --
    collision and contact Jacobian together).  Involves no formula from the oracle's own derivation."""
    m, v0 = load_case(name, capi)
    m.opt.cone = cone
    o = orc.Oracle(m)
    rng = np.random.default_rng(12)
    o.set("qpos", m.qpos0 + rng.uniform(-0.05, 0.05, m.nq) * (v0 is None))
    o.step(settle)
    q0 = o.get("qpos").copy()
    o.set("qvel", np.zeros(m.nv))

    def rows(q):
        o.set("qpos", q)
        o.forward()
        n = int(o.get("nefc")[0])
        geoms = [(int(o.get("contact_geom1")[c]), int(o.get("contact_geom2")[c])) for c in range(int(o.get("ncon")[0]))]
        return (o.get("efc_type")[:n].copy(), o.get("efc_id")[:n].copy(), o.get("efc_pos")[:n].copy(),
                o.get("efc_J")[:n * m.nv].reshape(n, m.nv).copy(), geoms)
    typ, ids, _, J, geoms = rows(q0)
    eq_type = m.eq_type
    want, first = [], {}
    for i in range(len(typ)):
        t, k = int(typ[i]), int(ids[i])
        if t == LIMIT_JOINT or t == FRICTIONLESS or (t == EQUALITY and eq_type[k] in (0, 1, 2, 3)):  # connect, weld, joint, tendon
            want.append(i)
        elif t == ELLIPTIC and first.setdefault(k, i) == i:
            # sphere-box / capsule-box are constructions of our own: once the sphere centre or the
            # capsule axis is inside the box the reported depth saturates at the radius and stops following the motion
            # (seen on hand_like: a finger capsule sunk into the palm box) -- not held to this check
            types = {int(m.geom_type[g]) for g in geoms[k]}
            if 6 in types and 0 not in types:
                continue
            want.append(i)
        elif t == PYRAMIDAL and first.setdefault(k, i) == i:
            # edges come in pairs J_n +- mu J_t: the mean of a pair is the normal row, its residual the distance
            types = {int(m.geom_type[g]) for g in geoms[k]}
            if 6 in types and 0 not in types:
                continue
            want.append(i)
            J[i] = 0.5 * (J[i] + J[i + 1])
    assert len(want) >= 2, (name, len(want))
    if cone == 0 and name == "humanoid_like.xml":
        assert any(int(typ[i]) == PYRAMIDAL for i in want)
    eps, checked = 1e-6, 0
    for trial in range(4):
        v = rng.uniform(-1, 1, m.nv)
        tp, ip, pp, _, gp = rows(integrate_pos(m, q0, v, eps))
        tm, im, pm, _, gm = rows(integrate_pos(m, q0, v, -eps))
        if not (np.array_equal(tp, typ) and np.array_equal(tm, typ) and np.array_equal(ip, ids) and np.array_equal(im, ids)
                and gp == geoms and gm == geoms):
            continue  # the perturbation changed the active set: rows no longer correspond
        fd = (pp - pm) / (2 * eps)
        np.testing.assert_allclose((J @ v)[want], fd[want], rtol=2e-5, atol=2e-6, err_msg=f"{name} trial {trial}")
        checked += 1
    assert checked >= 2, (name, checked)


@pytest.mark.parametrize("cone", ["elliptic", "pyramidal"])
def test_torsional_and_rolling_friction_saturate_at_mu_times_normal_force(cone, capi, orc):
    """condim 4 / 6 semantics in closed form: a sphere spinning about the contact normal is braked by the torque
    mu_torsion N, a sphere rolling without slipping by the torque mu_roll N (both coefficients are lengths):
    alpha = mu_t N / I  and  a = mu_r N / (1.4 m r) for a solid sphere.
    N is the weight under pyramidal cones.  Under elliptic cones the restated primal cost couples the rows: a saturated
    friction row raises the normal force of that step (the fast-spinning sphere sees up to five times its weight at
    single steps, 15 % over its weight on average here) -- whether libmujoco does the same is part of the unpinned
    parity, so the elliptic case takes N from the solver's own normal forces."""
    xml = f"""<mujoco><option timestep="0.001" cone="{cone}"/><worldbody>
      <geom type="plane" size="5 5 .1" condim="6" friction="1 0.03 0.004"/>
      <body pos="0 0 0.1"><freejoint/><geom size="0.1" condim="4" friction="1 0.03 0.004"/></body>
      <body pos="1 0 0.1"><freejoint/><geom size="0.1" condim="6" friction="1 0.03 0.004"/></body>
    </worldbody></mujoco>"""
    m = capi.Model.from_xml_string(xml)
    o = orc.Oracle(m)
    o.step(300)  # settle on the plane
    v = np.zeros(12)
    v[5] = 40.0              # first sphere: spin about the vertical
    v[6], v[10] = 1.0, 10.0  # second: rolls along +x without slipping (omega_y = v / r)
    o.set("qvel", v)
    o.step(50)
    a = o.get("qvel").copy()
    mass, r, steps = m.body_mass[1], 0.1, 200
    normal = np.zeros(2)
    for _ in range(steps):
        o.step(1)
        if cone == "elliptic":  # the spinning sphere hops (see above): a step without its contact adds no force
            f = o.get("efc_force")
            for c in range(int(o.get("ncon")[0])):
                ball, adr = int(o.get("contact_geom2")[c]) - 1, int(o.get("contact_efc_address")[c])
                normal[ball] += f[adr]
                scaled = f[adr + 1:adr + 6] / [1, 1, 0.03, 0.004, 0.004]
                assert abs(np.linalg.norm(scaled) - f[adr]) < 1e-6 * f[adr], (ball, f[adr:adr + 6])  # on the cone boundary
    b = o.get("qvel").copy()
    N = normal / steps if cone == "elliptic" else np.full(2, mass * 9.81)
    alpha, decel = -(b[5] - a[5]) / (steps * 0.001), -(b[6] - a[6]) / (steps * 0.001)
    tol = 0.01 if cone == "elliptic" else 0.08  # the pyramid only approximates the friction limit
    assert abs(alpha / (0.03 * N[0] / (0.4 * mass * r * r)) - 1) < tol, (alpha, N)
    assert abs(decel / (0.004 * N[1] / (1.4 * mass * r)) - 1) < tol, (decel, N)
    assert abs(b[10] * r - b[6]) < 1e-3  # still rolling without slipping


def test_limit_margins_activate_rows_before_the_limit(capi, orc):
    """joint / tendon margin: the limit row exists as soon as the distance to the limit drops below the margin, with
    efc_pos = that distance and the impedance evaluated at distance - margin."""
    xml = """<mujoco><compiler angle="radian"/><option gravity="0 0 0"/><worldbody>
      <body><joint name="a" axis="0 1 0" range="-0.5 0.5" limited="true" margin="0.05"/><geom size="0.1" pos="0.3 0 0"/>
        <body pos="0.3 0 0"><joint name="b" axis="0 1 0"/><geom size="0.05" pos="0.2 0 0"/></body></body></worldbody>
      <tendon><fixed name="t" limited="true" range="-1 0.6" margin="0.1"><joint joint="a" coef="1"/><joint joint="b" coef="1"/></fixed></tendon>
    </mujoco>"""
    m = capi.Model.from_xml_string(xml)
    for qa, qb, expect in ((0.40, 0.0, []), (0.47, 0.0, [(LIMIT_JOINT, 0.03, 0.05)]), (0.47, 0.1, [(LIMIT_JOINT, 0.03, 0.05), (LIMIT_TENDON, 0.03, 0.1)]),
                           (-0.48, -0.45, [(LIMIT_JOINT, 0.02, 0.05), (LIMIT_TENDON, 0.07, 0.1)])):
        o = orc.Oracle(m)
        o.set("qpos", [qa, qb])
        o.forward()
        n = int(o.get("nefc")[0])
        got = [(int(o.get("efc_type")[i]), o.get("efc_pos")[i], o.get("efc_margin")[i]) for i in range(n)]
        assert len(got) == len(expect), (qa, qb, got)
        for g, e in zip(got, expect):
            assert g[0] == e[0]
            np.testing.assert_allclose(g[1:], e[1:], atol=1e-12)
        for i in range(n):
            solimp = (m.jnt_solimp.reshape(-1, 5)[0] if got[i][0] == LIMIT_JOINT else m.tendon_solimp_lim.reshape(-1, 5)[0])
            np.testing.assert_allclose(o.get("efc_KBIP")[4 * i + 2], impedance(solimp, got[i][1] - got[i][2]), rtol=1e-12)
