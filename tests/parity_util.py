"""Helpers shared by the GPU parity tests: seeded inputs, field-by-field diffs against the CPU oracle, and the
state-injected single-step check (the north star's "per step" clause: both sides start every step from the same
state, so chaotic growth in contact-rich scenes cannot hide or fake a mismatch)."""
import numpy as np

TOL = 1e-5
STATE_FIELDS = ("qpos", "qvel", "act", "qacc_warmstart", "time")


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / (1.0 + np.abs(b)))) if a.size else 0.0


def perturbed(model, nenv, seed, amp=0.1):
    rng = np.random.default_rng(seed)
    qpos = np.tile(model.qpos0, (nenv, 1))
    qvel = np.zeros((nenv, model.nv))
    for j in range(model.njnt):
        t, qa, da = model.jnt_type[j], model.jnt_qposadr[j], model.jnt_dofadr[j]
        if t == 0:
            qpos[:, qa:qa + 2] += rng.uniform(-amp, amp, (nenv, 2))
            q = qpos[:, qa + 3:qa + 7] + rng.uniform(-amp, amp, (nenv, 4))
            qpos[:, qa + 3:qa + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
            qvel[:, da:da + 6] = rng.uniform(-amp, amp, (nenv, 6))
        elif t == 1:
            q = qpos[:, qa:qa + 4] + rng.uniform(-amp, amp, (nenv, 4))
            qpos[:, qa:qa + 4] = q / np.linalg.norm(q, axis=1, keepdims=True)
            qvel[:, da:da + 3] = rng.uniform(-amp, amp, (nenv, 3))
        else:
            qpos[:, qa] += rng.uniform(-amp, amp, nenv)
            qvel[:, da] = rng.uniform(-amp, amp, nenv)
    return qpos, qvel


def ctrl_sample(model, rng, nenv):
    if not model.nu:
        return np.zeros((nenv, 0))
    lo, hi = model.actuator_ctrlrange[:, 0], model.actuator_ctrlrange[:, 1]
    bad = hi <= lo
    lo, hi = np.where(bad, -1.0, lo), np.where(bad, 1.0, hi)
    return rng.uniform(lo, hi, (nenv, model.nu))


def compare_forward_fields(capi, model, sim, oracles, skip=(), tol=1e-8, tag=""):
    """Every mjData field of every env after a forward pass on both sides.  Integer fields exact, floats
    relative to the field's scale.  Returns the worst relative error seen."""
    ncon_g, nefc_g = sim.get("ncon")[:, 0], sim.get("nefc")[:, 0]
    skip = set(skip) | {"efc_AR", "warning", "solver_iter"}
    if not any(t in (1, 4, 5, 31, 32, 0) for t in model.sensor_type):  # only defined after mj_rnePostConstraint
        skip |= {"cacc", "cfrc_int", "cfrc_ext"}
    worst = 0.0
    cache = {}
    for e, o in enumerate(oracles):
        assert o.get("ncon")[0] == ncon_g[e], f"{tag} env {e}: ncon {ncon_g[e]} vs oracle {o.get('ncon')[0]}"
        assert o.get("nefc")[0] == nefc_g[e], f"{tag} env {e}: nefc {nefc_g[e]} vs oracle {o.get('nefc')[0]}"
        for fname in capi.FIELD_NAMES:
            n, is_int = model.field_size(capi.field_id(fname))
            if fname in skip or n <= 0:
                continue
            if fname not in cache:
                cache[fname] = sim.get(fname)
            gv, ov = cache[fname][e], o.get(fname)
            if fname.startswith("contact_"):
                k = (n // model.nconmax) * ncon_g[e]
                gv, ov = gv[:k], ov[:k]
            elif fname.startswith("efc_"):
                k = (n // model.njmax) * nefc_g[e]
                gv, ov = gv[:k], ov[:k]
            if not ov.size:
                continue
            if is_int:
                np.testing.assert_array_equal(gv, ov, err_msg=f"{tag}:{fname} env {e}")
            else:
                scale = 1e-12 + np.max(np.abs(ov))
                err = float(np.max(np.abs(gv - ov)) / scale)
                worst = max(worst, err)
                assert err < tol, f"{tag}:{fname} env {e}: {err:.3e}"
    return worst


def make_oracles(orc, model, qpos, qvel):
    out = []
    for e in range(qpos.shape[0]):
        o = orc.Oracle(model)
        o.set("qpos", qpos[e])
        o.set("qvel", qvel[e])
        out.append(o)
    return out


def injected_steps(model, sim, oracles, nsteps, rng, tol=TOL, tag="", extra_inputs=None, check_every=1):
    """Run nsteps on the GPU batch with fresh random controls.  Before every checked step the GPU state is copied into
    the oracles, both sides take one step, and qpos / qvel / qacc / act must agree to tol.  Returns (worst error,
    max nefc seen)."""
    worst, max_nefc = 0.0, 0
    nenv = len(oracles)
    for s in range(nsteps):
        check = (s % check_every) == 0
        if check:
            st = {k: sim.get(k) for k in STATE_FIELDS if model.field_size_by_name(k) > 0}
        ctrl = ctrl_sample(model, rng, nenv)
        if model.nu:
            sim.set("ctrl", ctrl)
        inputs = extra_inputs(s) if extra_inputs else {}
        for k, v in inputs.items():
            sim.set(k, v)
        sim.step(1)
        if not check:
            continue
        out = {k: sim.get(k) for k in ("qpos", "qvel", "qacc") + (("act",) if model.na else ())}
        nefc = sim.get("nefc")[:, 0]
        max_nefc = max(max_nefc, int(nefc.max()))
        for e, o in enumerate(oracles):
            for k, v in st.items():
                o.set(k, v[e])
            if model.nu:
                o.set("ctrl", ctrl[e])
            for k, v in inputs.items():
                o.set(k, v[e])
            o.step(1)
            assert int(o.get("nefc")[0]) == int(nefc[e]), f"{tag} step {s} env {e}: nefc {nefc[e]} vs {o.get('nefc')[0]}"
            for k, v in out.items():
                err = rel(v[e], o.get(k))
                worst = max(worst, err)
                assert err < tol, f"{tag} step {s} env {e} field {k}: {err:.3e} (nefc {nefc[e]})"
    return worst, max_nefc
