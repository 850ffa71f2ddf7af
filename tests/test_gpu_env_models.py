"""Per-env model variants: each env steps its own edited copy of the model -- what the reference's
mutating services do to its single env (set_body_state mass: callbacks.cpp:210-370, set_geom_properties friction / size
-> mj_setConst: :508-592, set_gravity: :462-506, set_equality_constraint_parameters: :641-738), batched as domain
randomisation.  Parity: one CPU oracle per env, built from that env's own variant."""
import numpy as np
import pytest

from conftest import model_path
from parity_util import TOL, compare_forward_fields, perturbed, rel

pytestmark = pytest.mark.gpu


def make_variants(capi, name, edits):
    out = []
    for edit in edits:
        m = capi.Model.from_xml_file(model_path(name))
        edit(m)
        m.set_const()
        out.append(m)
    return out


def test_pendulum_scene_variants_match_their_own_oracles(capi, orc):
    from mujoco_ros_pkgs_b200.batch import BatchSim

    def v0(m):
        pass

    def v1(m):  # heavier end link + lower gravity
        b = m.name2id(capi.OBJ_BODY, "end_link")
        m.body_mass[b] *= 2.5
        m.body_inertia[b] *= 2.5
        m.opt.gravity[2] = -3.7

    def v2(m):  # slippery, larger ball; sideways gravity
        g = m.name2id(capi.OBJ_GEOM, "ball")
        m.geom_friction[g, 0] = 0.2
        m.geom_size[g, 0] = 0.07
        m.geom_rbound[g] = 0.07
        m.opt.gravity[0] = 2.0

    def v3(m):  # joint damping switched on, stiffer contact
        m.dof_damping[:] = 0.3
        g = m.name2id(capi.OBJ_GEOM, "ball")
        m.geom_solref[g, 0] = 0.01

    variants = make_variants(capi, "pendulum_scene.xml", [v0, v1, v2, v3])
    nenv = 12
    env_model = np.array([0, 1, 2, 3, 3, 2, 1, 0, 1, 1, 2, 3], dtype=np.int32)
    qpos, qvel = perturbed(variants[0], nenv, 7, 0.2)
    sim = BatchSim(variants[0], nenv)
    sim.set_env_models(variants, env_model)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = []
    for e in range(nenv):
        o = orc.Oracle(variants[env_model[e]])
        o.set("qpos", qpos[e])
        o.set("qvel", qvel[e])
        oracles.append(o)
    worst = 0.0
    for s in range(300):
        sim.step(1)
        for o in oracles:
            o.step(1)
        if s % 50 == 49:
            gq, gv = sim.get("qpos"), sim.get("qvel")
            for e, o in enumerate(oracles):
                worst = max(worst, rel(gq[e], o.get("qpos")), rel(gv[e], o.get("qvel")))
            assert worst < TOL, (s, worst)
    # the variants really differ from each other
    gq = sim.get("qpos")
    assert rel(gq[0], gq[1]) > 1e-3 and rel(gq[0], gq[2]) > 1e-3 and rel(gq[0], gq[3]) > 1e-4
    # envs sharing a variant and ... (different initial states, so just finite)
    assert np.all(np.isfinite(gq))
    # every mjData field after a forward pass, env by env against its own variant
    sim.keep_intermediates(True)
    sim.forward()
    for o in oracles:
        o.forward()
    for v in range(4):
        sel = [e for e in range(nenv) if env_model[e] == v]
        sub_sim = _EnvSubset(sim, sel)
        compare_forward_fields(capi, variants[v], sub_sim, [oracles[e] for e in sel], skip={"xfrc_applied"}, tag=f"variant {v}")
    # back to one shared model
    sim.set_env_models([], None)
    sim.keep_intermediates(False)
    sim.reset()
    sim.set("qpos", qpos)
    sim.step(20)
    ref = BatchSim(variants[0], nenv)
    ref.set("qpos", qpos)
    ref.step(20)
    np.testing.assert_array_equal(sim.get("qpos"), ref.get("qpos"))


class _EnvSubset:
    """view of a BatchSim restricted to some envs, for compare_forward_fields"""

    def __init__(self, sim, envs):
        self.sim, self.envs = sim, envs

    def get(self, name):
        return self.sim.get(name)[self.envs]


def test_panda_equality_and_mass_randomisation_batch(capi, orc):
    """C2-sized use: 256 envs over 8 variants of the Panda (link masses x U(0.7, 1.3), finger friction, gravity tilt)
    under PGS with random controls: injected single steps against per-variant oracles."""
    from mujoco_ros_pkgs_b200.batch import BatchSim
    from parity_util import ctrl_sample

    rng = np.random.default_rng(3)

    def randomise(k):
        def f(m):
            r = np.random.default_rng(100 + k)
            m.body_mass[1:] *= r.uniform(0.7, 1.3, m.nbody - 1)
            m.body_inertia[1:] *= r.uniform(0.7, 1.3, (m.nbody - 1, 1))
            m.geom_friction[:, 0] *= r.uniform(0.5, 1.5)
            m.opt.gravity[0] = r.uniform(-1, 1)
        return f

    variants = make_variants(capi, "panda_like.xml", [randomise(k) for k in range(8)])
    nenv = 256
    env_model = rng.integers(0, 8, nenv).astype(np.int32)
    qpos, qvel = perturbed(variants[0], nenv, 5, 0.1)
    sim = BatchSim(variants[0], nenv)
    sim.set_env_models(variants, env_model)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    for s in range(400):
        if s % 25 == 0:
            sim.set("ctrl", ctrl_sample(variants[0], rng, nenv))
        sim.step(1)
    check = list(range(0, nenv, 9))
    oracles = {e: orc.Oracle(variants[env_model[e]]) for e in check}
    worst = 0.0
    for s in range(20):
        st = {k: sim.get(k) for k in ("qpos", "qvel", "qacc_warmstart", "time")}
        ctrl = ctrl_sample(variants[0], rng, nenv)
        sim.set("ctrl", ctrl)
        sim.step(1)
        gq, gv, ga = sim.get("qpos"), sim.get("qvel"), sim.get("qacc")
        for e, o in oracles.items():
            for k, v in st.items():
                o.set(k, v[e])
            o.set("ctrl", ctrl[e])
            o.step(1)
            worst = max(worst, rel(gq[e], o.get("qpos")), rel(gv[e], o.get("qvel")), rel(ga[e], o.get("qacc")))
    assert worst < TOL, worst
    print(f"panda variants: worst injected-step error {worst:.2e}")


def test_env_models_reject_size_and_topology_changes(capi):
    from mujoco_ros_pkgs_b200.batch import BatchSim

    a = capi.Model.from_xml_file(model_path("pendulum_scene.xml"))
    b = capi.Model.from_xml_file(model_path("panda_like.xml"))
    sim = BatchSim(a, 4)
    with pytest.raises(capi.B2mjError):
        sim.set_env_models([a, b], [0, 1, 0, 1])
    c = capi.Model.from_xml_file(model_path("pendulum_scene.xml"))
    c.qpos0[0] += 0.1
    with pytest.raises(capi.B2mjError):
        sim.set_env_models([a, c], [0, 1, 0, 1])
    with pytest.raises(capi.B2mjError):
        sim.set_env_models([a], [0, 1, 0, 0])
