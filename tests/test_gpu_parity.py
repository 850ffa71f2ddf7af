"""GPU parity tests proper: the CUDA batched step (through the C-ABI) against the CPU oracle on the
same seeded inputs, against the committed golden fixtures, and through size-independent properties
at the BASELINE batch size.  Tolerance: 1e-5 relative per step in FP64 (BASELINE.json north_star);
time and contact-pair indexing bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

TOL = 1e-5


def rel(a, b):
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
    return rng.uniform(lo, hi, (nenv, model.nu))


@pytest.fixture(scope="module")
def BatchSim():
    from mujoco_ros_pkgs_b200.batch import BatchSim as B

    return B


@pytest.mark.parametrize("name", ["panda_like.xml", "pendulum_scene.xml", "equality_scene.xml", "box_stack.xml",
                                  "hand_like.xml", "humanoid_like.xml", "bin.xml"])
def test_forward_fields_match_oracle(name, load_model, orc, capi, BatchSim):
    """Every mjData field after mj_forward, env by env (per-stage diff)."""
    model = load_model(name)
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 11, 0.02 if name in ("hand_like.xml", "humanoid_like.xml", "bin.xml") else 0.1)
    ctrl = ctrl_sample(model, np.random.default_rng(5), nenv)
    sim = BatchSim(model, nenv)
    sim.keep_intermediates(True)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    if model.nu:
        sim.set("ctrl", ctrl)
    sim.forward()
    ncon_g, nefc_g = sim.get("ncon")[:, 0], sim.get("nefc")[:, 0]
    skip = {"efc_AR", "xfrc_applied", "warning", "solver_iter", "efc_state"}
    if not any(t in (1, 4, 5, 31, 32, 0) for t in model.sensor_type):  # only defined after mj_rnePostConstraint
        skip |= {"cacc", "cfrc_int", "cfrc_ext"}
    for e in range(nenv):
        o = orc.Oracle(model)
        o.set("qpos", qpos[e])
        o.set("qvel", qvel[e])
        if model.nu:
            o.set("ctrl", ctrl[e])
        o.forward()
        assert o.get("ncon")[0] == ncon_g[e] and o.get("nefc")[0] == nefc_g[e]
        for fname in capi.FIELD_NAMES:
            n, is_int = model.field_size(capi.field_id(fname))
            if fname in skip or n <= 0:
                continue
            gv, ov = sim.get(fname)[e], o.get(fname)
            if fname.startswith("contact_"):
                k = (n // model.nconmax) * ncon_g[e]
                gv, ov = gv[:k], ov[:k]
            elif fname.startswith("efc_"):
                k = (n // model.njmax) * nefc_g[e]
                gv, ov = gv[:k], ov[:k]
            if is_int:
                np.testing.assert_array_equal(gv, ov, err_msg=f"{name}:{fname} env {e}")
            else:
                scale = 1e-12 + np.max(np.abs(ov)) if ov.size else 1.0
                assert np.max(np.abs(gv - ov)) / scale < 1e-8 if ov.size else True, f"{name}:{fname} env {e}"


@pytest.mark.parametrize("name,nsteps", [("panda_like.xml", 1000), ("pendulum_scene.xml", 300), ("equality_scene.xml", 300),
                                         ("box_stack.xml", 600), ("hand_like.xml", 300), ("humanoid_like.xml", 60),
                                         ("bin.xml", 120)])
def test_rollout_matches_oracle(name, nsteps, load_model, orc, BatchSim):
    """State divergence vs the oracle < 1e-5 per step and over the rollout; time bit-exact;
    contact-pair indices identical (checked at every 50th step)."""
    model = load_model(name)
    nenv = 4 if name == "bin.xml" else 16
    qpos, qvel = perturbed(model, nenv, 21, 0.02 if name in ("hand_like.xml", "humanoid_like.xml", "bin.xml") else 0.1)
    rng = np.random.default_rng(9)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = []
    for e in range(nenv):
        o = orc.Oracle(model)
        o.set("qpos", qpos[e])
        o.set("qvel", qvel[e])
        oracles.append(o)
    worst = 0.0
    for s in range(1, nsteps + 1):
        ctrl = ctrl_sample(model, rng, nenv)
        if model.nu:
            sim.set("ctrl", ctrl)
        sim.step(1)
        for e, o in enumerate(oracles):
            if model.nu:
                o.set("ctrl", ctrl[e])
            o.step(1)
        if s % 50 == 0 or s == 1 or s == nsteps:
            gq, gv, gt = sim.get("qpos"), sim.get("qvel"), sim.get("time")[:, 0]
            oq = np.stack([o.get("qpos") for o in oracles])
            ov = np.stack([o.get("qvel") for o in oracles])
            worst = max(worst, rel(gq, oq), rel(gv, ov))
            np.testing.assert_array_equal(gt, [o.time for o in oracles])
            assert worst < TOL, f"{name}: state diverged at step {s}: {worst:.3e}"
    sim.keep_intermediates(True)
    sim.forward()
    ncon = sim.get("ncon")[:, 0]
    g1, g2 = sim.get("contact_geom1"), sim.get("contact_geom2")
    for e, o in enumerate(oracles):
        o.forward()
        assert o.get("ncon")[0] == ncon[e]
        np.testing.assert_array_equal(g1[e][:ncon[e]], o.get("contact_geom1")[:ncon[e]])
        np.testing.assert_array_equal(g2[e][:ncon[e]], o.get("contact_geom2")[:ncon[e]])


@pytest.mark.parametrize("name", ["panda_like", "pendulum_scene", "equality_scene", "box_stack", "hand_like",
                                  "humanoid_like", "bin"])
def test_golden_trajectories(name, load_model, BatchSim):
    """Committed fixtures (tools/make_golden.py): same inputs in every env, same trajectory out."""
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    model = load_model(f"{name}.xml")
    nenv = 5
    sim = BatchSim(model, nenv)
    sim.set("qpos", np.tile(g["qpos_init"], (nenv, 1)))
    sim.set("qvel", np.tile(g["qvel_init"], (nenv, 1)))
    for k in range(g["qpos"].shape[0]):
        if model.nu:
            sim.set("ctrl", np.tile(g["ctrl"][k], (nenv, 1)))
        sim.step(int(g["stride"]))
        q, v = sim.get("qpos"), sim.get("qvel")
        assert rel(q, np.tile(g["qpos"][k], (nenv, 1))) < TOL, (name, k)
        assert rel(v, np.tile(g["qvel"][k], (nenv, 1))) < TOL, (name, k)
        np.testing.assert_array_equal(q, np.tile(q[0], (nenv, 1)))  # envs are bitwise replicas


def test_time_and_guards(load_model, capi, BatchSim):
    # mujoco_env_test.cpp:198-200,219-221 (time), :255-275 (n <= 0 refused)
    model = load_model("pendulum_scene.xml")
    sim = BatchSim(model, 3)
    sim.step(1)
    np.testing.assert_array_equal(sim.get("time")[:, 0], model.opt.timestep)
    sim.step(99)
    assert np.all(np.abs(sim.get("time")[:, 0] - 100 * model.opt.timestep) < 1e-6)
    with pytest.raises(capi.B2mjError):
        sim.step(0)
    with pytest.raises(capi.B2mjError):
        sim.step(-3)
    with pytest.raises(capi.B2mjError):
        sim.step_end()  # without step_begin


def test_reset_all_and_masked(load_model, BatchSim):
    # mujoco_env_test.cpp:507,519-526; masked form = per-env mj_resetData
    model = load_model("pendulum_scene.xml")
    nenv = 6
    sim = BatchSim(model, nenv)
    qpos, qvel = perturbed(model, nenv, 3)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    sim.step(25)
    before_q = sim.get("qpos")
    mask = np.array([1, 0, 1, 0, 0, 1], dtype=np.uint8)
    sim.reset(mask)
    q, v, t = sim.get("qpos"), sim.get("qvel"), sim.get("time")[:, 0]
    for e in range(nenv):
        if mask[e]:
            np.testing.assert_array_equal(q[e], model.qpos0)
            np.testing.assert_array_equal(v[e], 0)
            assert t[e] == 0
        else:
            np.testing.assert_array_equal(q[e], before_q[e])
            assert t[e] > 0
    sim.reset()
    np.testing.assert_array_equal(sim.get("qpos"), np.tile(model.qpos0, (nenv, 1)))
    np.testing.assert_array_equal(sim.get("time"), 0)


def test_hanging_pendulum_bitwise_static(load_model, BatchSim):
    # mujoco_sensors_test.cpp:389-391,584: sensor GT variance exactly 0 over 1001 single steps
    model = load_model("pendulum_scene.xml")
    sim = BatchSim(model, 4)
    s0 = None
    for _ in range(1001):
        sim.step(1)
        s = sim.get("sensordata")
        s0 = s if s0 is None else s0
        np.testing.assert_array_equal(s, s0)
    np.testing.assert_array_equal(sim.get("qvel")[:, :5], 0.0)
    assert np.all(sim.get("ncon")[:, 0] == 1)


def test_split_step_equals_step(load_model, BatchSim):
    # control hook placement (mujoco_env.h:242-246): begin -> host writes ctrl -> end == one step
    model = load_model("panda_like.xml")
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 4)
    rng = np.random.default_rng(2)
    a, b = BatchSim(model, nenv), BatchSim(model, nenv)
    for s in (a, b):
        s.set("qpos", qpos)
        s.set("qvel", qvel)
    for _ in range(30):
        ctrl = ctrl_sample(model, rng, nenv)
        a.set("ctrl", ctrl)
        a.step(1)
        b.step_begin()
        b.set("ctrl", ctrl)  # what a controlCallback would do
        b.step_end()
    assert rel(a.get("qpos"), b.get("qpos")) < 1e-12
    assert rel(a.get("qvel"), b.get("qvel")) < 1e-12


def test_multi_step_launch_equals_single_steps(load_model, BatchSim):
    model = load_model("pendulum_scene.xml")
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 8)
    a, b = BatchSim(model, nenv), BatchSim(model, nenv)
    for s in (a, b):
        s.set("qpos", qpos)
        s.set("qvel", qvel)
    a.step(40)
    for _ in range(40):
        b.step(1)
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))
    np.testing.assert_array_equal(a.get("qvel"), b.get("qvel"))


def test_set_device_matches_host_set(load_model, BatchSim):
    import torch

    model = load_model("panda_like.xml")
    nenv = 32
    sim = BatchSim(model, nenv)
    ctrl = ctrl_sample(model, np.random.default_rng(1), nenv)
    t = torch.from_numpy(ctrl).cuda()
    sim.set_device("ctrl", t.data_ptr(), model.nu)
    sim.sync()
    np.testing.assert_array_equal(sim.get("ctrl"), ctrl)


def test_bad_state_triggers_reset_and_warning(load_model, BatchSim):
    # mj_checkPos semantics: NaN qpos -> warning counter + mj_resetData for that env only
    model = load_model("pendulum_scene.xml")
    nenv = 4
    sim = BatchSim(model, nenv)
    q = np.tile(model.qpos0, (nenv, 1))
    q[2, 4] = np.nan
    sim.set("qpos", q)
    sim.step(1)
    w = sim.get("warning")
    assert w[2, 4] == 1 and w[[0, 1, 3]].sum() == 0  # B2MJ_WARN_BADQPOS = 4
    assert np.all(np.isfinite(sim.get("qpos")))


def test_full_batch_properties(load_model, BatchSim):
    """BASELINE size (4096 envs): replicas of one input stay bitwise identical across the batch and
    equal the 1-env run (batch independence), quaternions stay unit, no warnings."""
    model = load_model("panda_like.xml")
    nenv = 4096
    qpos, qvel = perturbed(model, 4, 31)
    rng = np.random.default_rng(6)
    big, small = BatchSim(model, nenv), BatchSim(model, 4)
    big.set("qpos", np.tile(qpos, (nenv // 4, 1)))
    big.set("qvel", np.tile(qvel, (nenv // 4, 1)))
    small.set("qpos", qpos)
    small.set("qvel", qvel)
    for _ in range(50):
        c = ctrl_sample(model, rng, 4)
        big.set("ctrl", np.tile(c, (nenv // 4, 1)))
        small.set("ctrl", c)
        big.step(1)
        small.step(1)
    q = big.get("qpos").reshape(nenv // 4, 4, model.nq)
    np.testing.assert_array_equal(q, np.broadcast_to(q[0], q.shape))
    np.testing.assert_array_equal(q[0], small.get("qpos"))
    assert big.get("warning").sum() == 0


@pytest.mark.parametrize("chunk", ["0", "7"])
def test_rollout_equals_stepping(chunk, load_model, BatchSim, monkeypatch):
    """b2mj_rollout (one launch, device ctrl stream, trajectory out) == nsteps x (set ctrl; b2mj_step), for both
    schedules: static one-env-per-warp (chunk 0) and the ticketed persistent grid (chunks of 7 steps)."""
    import torch

    monkeypatch.setenv("B2MJ_ROLLOUT_CHUNK", chunk)

    model = load_model("panda_like.xml")
    nenv, K = 64, 30
    qpos, qvel = perturbed(model, nenv, 12)
    rng = np.random.default_rng(4)
    ctrl = np.stack([ctrl_sample(model, rng, nenv) for _ in range(K)])
    a, b = BatchSim(model, nenv), BatchSim(model, nenv)
    for s in (a, b):
        s.set("qpos", qpos)
        s.set("qvel", qvel)
    cdev = torch.from_numpy(ctrl).cuda()
    tq = torch.zeros(K, nenv, model.nq, dtype=torch.float64, device="cuda")
    tv = torch.zeros(K, nenv, model.nv, dtype=torch.float64, device="cuda")
    ts = torch.zeros(K, nenv, model.nsensordata, dtype=torch.float64, device="cuda")
    a.rollout(K, cdev.data_ptr(), tq.data_ptr(), tv.data_ptr(), ts.data_ptr())
    a.sync()
    for k in range(K):
        b.set("ctrl", ctrl[k])
        b.step(1)
        np.testing.assert_array_equal(tq[k].cpu().numpy(), b.get("qpos"))
        np.testing.assert_array_equal(tv[k].cpu().numpy(), b.get("qvel"))
        np.testing.assert_array_equal(ts[k].cpu().numpy(), b.get("sensordata"))
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))
    np.testing.assert_array_equal(a.get("time"), b.get("time"))


def test_contact_rich_single_steps_match_oracle_across_nefc(load_model, orc, BatchSim):
    """PGS parity for every constraint count the C2 workload produces, not just the typical few rows: the batch
    is driven into contact with random controls, then envs are picked so that each nefc value seen (in particular
    17..25: past the one-row-per-lane half, and 18/19 where AR plus its row constants no longer fit the on-chip
    window) gets a state-injected single step on the oracle.  Tolerance 1e-5 relative on qacc / qvel / qpos."""
    model = load_model("panda_like.xml")
    nenv = 2048
    rng = np.random.default_rng(5)
    qpos, qvel = perturbed(model, nenv, 77, 0.1)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    picked = {}
    for block in range(12):
        for _ in range(100):
            sim.set("ctrl", ctrl_sample(model, rng, nenv))
            sim.step(1)
        # state BEFORE the probe step
        st = {k: sim.get(k) for k in ("qpos", "qvel", "qacc_warmstart", "time")}
        ctrl = ctrl_sample(model, rng, nenv)
        sim.set("ctrl", ctrl)
        sim.step(1)
        nefc = sim.get("nefc")[:, 0]
        out = {k: sim.get(k) for k in ("qpos", "qvel", "qacc")}
        iters = sim.get("solver_iter")[:, 0]
        for e in range(nenv):
            n = int(nefc[e])
            if n >= 6 and len(picked.get(n, [])) < 2:
                picked.setdefault(n, []).append((e, {k: v[e].copy() for k, v in st.items()}, ctrl[e].copy(),
                                                 {k: v[e].copy() for k, v in out.items()}, int(iters[e])))
    assert picked, "the workload produced no contact-rich env"
    seen = sorted(picked)
    assert seen[-1] >= 17, f"expected constraint counts beyond 16 rows, saw {seen}"
    worst = 0.0
    for n in seen:
        for e, st, ctrl, out, it in picked[n]:
            o = orc.Oracle(model)
            o.set("qpos", st["qpos"])
            o.set("qvel", st["qvel"])
            o.set("qacc_warmstart", st["qacc_warmstart"])
            o.set("ctrl", ctrl)
            o.step(1)
            assert o.get("nefc")[0] == n
            err = max(rel(out["qacc"], o.get("qacc")), rel(out["qvel"], o.get("qvel")), rel(out["qpos"], o.get("qpos")))
            worst = max(worst, err)
            assert err < TOL, f"nefc={n} env={e} iters={it}: single-step mismatch {err:.3e}"
    print(f"contact-rich single steps: nefc values {seen}, worst rel err {worst:.2e}")


def test_runtime_model_mutation_matches_oracle(capi, orc, BatchSim):
    """The reference's mutating services (callbacks.cpp:462-738: set_gravity, set_body_state mass, set_geom_properties
    friction / size -> mj_setConst, equality parameters) as host-side model edits + b2mj_model_set_const +
    b2mj_model_update on a live handle: the batch must then step exactly like an oracle created from the edited model."""
    from conftest import model_path

    model = capi.Model.from_xml_file(model_path("pendulum_scene.xml"))
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 3, 0.2)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    sim.step(50)
    # ---- edit: gravity (set_gravity), a body mass + inertia (set_body_state), geom friction / size (set_geom_properties)
    model.opt.gravity[0] = 1.5
    model.opt.gravity[2] = -4.0
    b = model.name2id(capi.OBJ_BODY, "end_link")
    model.body_mass[b] *= 2.0
    model.body_inertia[b] *= 2.0
    g = model.name2id(capi.OBJ_GEOM, "ball")
    model.geom_friction[g, 0] = 0.3
    model.geom_size[g, 0] = 0.06
    model.geom_rbound[g] = 0.06
    model.set_const()              # mj_setConst: subtree masses, dof_M0, invweights
    sim.model_update(model)
    st = {k: sim.get(k) for k in ("qpos", "qvel", "qacc_warmstart", "time")}
    sim.step(100)
    gq, gv = sim.get("qpos"), sim.get("qvel")
    worst = 0.0
    for e in range(nenv):
        o = orc.Oracle(model)
        for k in ("qpos", "qvel", "qacc_warmstart", "time"):
            o.set(k, st[k][e])
        o.step(100)
        worst = max(worst, rel(gq[e], o.get("qpos")), rel(gv[e], o.get("qvel")))
    assert worst < TOL, worst
    # and it differs from the un-edited dynamics (the update really took effect)
    ref = capi.Model.from_xml_file(model_path("pendulum_scene.xml"))
    o = orc.Oracle(ref)
    for k in ("qpos", "qvel", "qacc_warmstart", "time"):
        o.set(k, st[k][0])
    o.step(100)
    assert rel(gq[0], o.get("qpos")) > 1e-3
    # size fields cannot change on a live handle
    import ctypes

    bad = capi.Model.from_xml_file(model_path("panda_like.xml"))
    with pytest.raises(Exception):
        sim.model_update(bad)


def test_step_host_equals_set_step_get(load_model, BatchSim):
    """b2mj_step_host (one call: ctrl up, step, state down, one sync) == b2mj_set + b2mj_step + b2mj_get, bitwise."""
    model = load_model("pendulum_scene.xml" if False else "panda_like.xml")
    nenv = 96
    qpos, qvel = perturbed(model, nenv, 31)
    rng = np.random.default_rng(8)
    a, b = BatchSim(model, nenv), BatchSim(model, nenv)
    for s_ in (a, b):
        s_.set("qpos", qpos)
        s_.set("qvel", qvel)
    hq = np.empty((nenv, model.nq)); hv = np.empty((nenv, model.nv)); hs = np.empty((nenv, model.nsensordata))
    for k in range(25):
        ctrl = np.ascontiguousarray(ctrl_sample(model, rng, nenv))
        a.step_host(1, ctrl, hq, hv, hs)
        b.set("ctrl", ctrl)
        b.step(1)
        np.testing.assert_array_equal(hq, b.get("qpos"))
        np.testing.assert_array_equal(hv, b.get("qvel"))
        np.testing.assert_array_equal(hs, b.get("sensordata"))
    a.step_host(3)   # no transfers at all: keeps ctrl
    b.step(3)
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))
    with pytest.raises(Exception):
        a.step_host(0)


def test_step_host_zero_copy_with_pinned_buffers(load_model, BatchSim):
    """Pinned host buffers take the zero-copy path of b2mj_step_host (the step kernel reads ctrl from and writes qpos /
    qvel / sensordata to mapped host memory); pageable buffers take the copy path.  Both must be bitwise the plain
    set / step / get sequence, and the streamed controls must stick as the env's ctrl."""
    import torch

    model = load_model("panda_like.xml")
    nenv = 300   # more than one CTA wave worth of 2-env CTAs is not needed; partial last CTA on purpose
    qpos, qvel = perturbed(model, nenv, 33)
    rng = np.random.default_rng(9)
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()  # noqa: E731
    a, b, c = BatchSim(model, nenv), BatchSim(model, nenv), BatchSim(model, nenv)
    for s_ in (a, b, c):
        s_.set("qpos", qpos)
        s_.set("qvel", qvel)
    pq, pv, ps, pc = pin(nenv, model.nq), pin(nenv, model.nv), pin(nenv, model.nsensordata), pin(nenv, model.nu)
    hq = np.empty((nenv, model.nq)); hv = np.empty((nenv, model.nv)); hs = np.empty((nenv, model.nsensordata))
    for k in range(40):
        ctrl = np.ascontiguousarray(ctrl_sample(model, rng, nenv))
        np.copyto(pc, ctrl)
        pq.fill(np.nan); pv.fill(np.nan); ps.fill(np.nan)
        a.step_host(1, pc, pq, pv, ps)          # zero-copy
        c.step_host(1, ctrl, hq, hv, hs)        # copies
        b.set("ctrl", ctrl)
        b.step(1)
        for z, h, name in ((pq, hq, "qpos"), (pv, hv, "qvel"), (ps, hs, "sensordata")):
            ref = b.get(name)
            np.testing.assert_array_equal(z, ref, err_msg=f"zero-copy {name} step {k}")
            np.testing.assert_array_equal(h, ref, err_msg=f"copy {name} step {k}")
    np.testing.assert_array_equal(a.get("ctrl"), b.get("ctrl"))   # streamed controls were written back to the record
    a.step(2); b.step(2)                                          # and are what a plain step uses next
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))
    # outputs only (no ctrl): still zero-copy, ctrl kept
    a.step_host(1, None, pq, pv, None); b.step(1)
    np.testing.assert_array_equal(pq, b.get("qpos"))


def test_launch_order_does_not_change_results(tmp_path):
    """The heaviest-first launch order (b2k_order_kernel) only permutes which warp runs which env: a contact-rich
    4096-env run (two waves: the order is refreshed after every step, on the side stream for this small model) must be
    bitwise identical with the synchronous refresh (B2MJ_ORDER_SYNC=1) and with the reordering disabled
    (B2MJ_NO_REORDER=1); separate processes, the switches are read once."""
    import subprocess
    import sys

    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from mujoco_ros_pkgs_b200 import _capi\n"
        "from mujoco_ros_pkgs_b200.batch import BatchSim\n"
        "m = _capi.Model.from_xml_file(%r)\n"
        "rng = np.random.default_rng(2)\n"
        "n = 4096\n"
        "sim = BatchSim(m, n)\n"
        "sim.set('qpos', np.tile(m.qpos0, (n, 1)) + rng.uniform(-0.3, 0.3, (n, m.nq)))\n"
        "lo, hi = m.actuator_ctrlrange[:, 0], m.actuator_ctrlrange[:, 1]\n"
        "for k in range(700):\n"
        "    if k %% 50 == 0: sim.set('ctrl', rng.uniform(lo, hi, (n, m.nu)))\n"
        "    sim.step(1)\n"
        "sim.step(40)\n"
        "np.save(sys.argv[1], np.concatenate([sim.get('qpos'), sim.get('qvel'), sim.get('nefc').astype(float)], axis=1))\n"
    ) % (ROOT, os.path.join(ROOT, "mujoco_ros_pkgs_b200", "models", "panda_like.xml"))
    outs = []
    for tag, extra in (("on", {}), ("sync", {"B2MJ_ORDER_SYNC": "1"}), ("off", {"B2MJ_NO_REORDER": "1"})):
        out = str(tmp_path / f"state_{tag}.npy")
        env = dict(os.environ, **extra)
        subprocess.run([sys.executable, "-c", code, out], check=True, env=env, timeout=600)
        outs.append(np.load(out))
    assert outs[0][:, -1].max() >= 8, "the run should reach contact-rich states"
    np.testing.assert_array_equal(outs[0], outs[2])
    np.testing.assert_array_equal(outs[1], outs[2])
