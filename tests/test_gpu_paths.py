"""GPU parity for the device code paths round 1 left untested (VERDICT r1 "What's weak" 3, 4): every solver x cone x
integrator combination the reference exposes (viewer.cpp:579-609) on models that reach each PGS size class (<= 32 rows
owner-computes, 33..64 two-slot register form, > 64 matrix-free), the CG solver, elliptic cones under PGS (QCQP block
updates), RK4 with PGS, activation dynamics (na > 0), xfrc_applied, mocap bodies, the BADQVEL / BADQACC resets,
efc_state, and the reference's own five MJCF files compiled verbatim.  All through the C-ABI against the CPU oracle;
tolerance 1e-5 relative per step (BASELINE.json north_star), integer fields exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from parity_util import (TOL, compare_forward_fields, ctrl_sample, injected_steps, make_oracles, perturbed, rel)

pytestmark = pytest.mark.gpu

PGS, CG, NEWTON = 0, 1, 2
STATE = ("qpos", "qvel", "act", "qacc_warmstart", "time", "ctrl")
PYR, ELL = 0, 1
EULER, RK4, IMPLICIT, IMPLICITFAST = 0, 1, 2, 3


@pytest.fixture(scope="module")
def BatchSim():
    from mujoco_ros_pkgs_b200.batch import BatchSim as B

    return B


def variant(capi, name, solver=None, cone=None, integrator=None):
    from conftest import model_path

    m = capi.Model.from_xml_file(model_path(name))
    if solver is not None:
        m.opt.solver = solver
    if cone is not None:
        m.opt.cone = cone
    if integrator is not None:
        m.opt.integrator = integrator
    return m


# (model, solver, cone, integrator, settle steps, checked steps, perturbation, envs)
MATRIX = [
    ("panda_like.xml", PGS, ELL, None, 450, 40, 0.1, 16),          # PGS + elliptic cones, <= 32 rows (regT<1> cone path)
    ("panda_like.xml", CG, None, None, 450, 40, 0.1, 16),          # CG
    ("panda_like.xml", NEWTON, ELL, None, 450, 40, 0.1, 16),
    ("box_stack.xml", PGS, ELL, None, 150, 40, 0.1, 8),            # stacked boxes: 24 cone rows
    ("box_stack.xml", PGS, PYR, None, 150, 40, 0.1, 8),            # pyramidal: 33..64 rows -> regT<2>
    ("box_stack.xml", CG, None, None, 150, 40, 0.1, 8),
    ("equality_scene.xml", PGS, None, None, 50, 40, 0.1, 8),       # equality rows (lo = -inf) under PGS
    ("equality_scene.xml", CG, None, None, 50, 40, 0.1, 8),
    ("humanoid_like.xml", PGS, PYR, None, 120, 30, 0.02, 8),       # 33..64 rows
    ("humanoid_like.xml", PGS, ELL, None, 120, 30, 0.02, 8),
    ("humanoid_like.xml", CG, None, None, 120, 30, 0.02, 8),
    ("hand_like.xml", PGS, PYR, RK4, 100, 20, 0.02, 8),            # RK4 + PGS, 33..64 rows
    ("hand_like.xml", PGS, ELL, RK4, 100, 20, 0.02, 8),            # RK4 + PGS + elliptic
    ("hand_like.xml", CG, None, RK4, 100, 20, 0.02, 8),
    ("hand_like.xml", NEWTON, PYR, EULER, 100, 20, 0.02, 8),
    ("bin.xml", PGS, ELL, None, 100, 10, 0.02, 4),                 # > 64 rows: matrix-free PGS with cone blocks
    ("bin.xml", PGS, PYR, None, 100, 10, 0.02, 4),                 # > 64 scalar rows
    ("bin.xml", CG, None, None, 100, 10, 0.02, 4),
    ("bin.xml", NEWTON, PYR, None, 100, 10, 0.02, 4),
    # implicit-in-velocity integrators (stages_implicit.cuh): dense inverse path (nv <= 16), sparse L'DL path, dense LU
    ("panda_like.xml", None, None, IMPLICITFAST, 450, 40, 0.1, 16),
    ("panda_like.xml", None, None, IMPLICIT, 450, 40, 0.1, 16),
    ("humanoid_like.xml", None, None, IMPLICITFAST, 120, 30, 0.02, 8),
    ("humanoid_like.xml", None, None, IMPLICIT, 120, 30, 0.02, 8),
    ("hand_like.xml", None, None, IMPLICITFAST, 100, 20, 0.02, 8),
    ("hand_like.xml", None, None, IMPLICIT, 100, 20, 0.02, 8),
    ("actuated_arm.xml", None, None, IMPLICITFAST, 50, 40, 0.2, 8),  # tendon transmission, affine gain / bias, na > 0
    ("actuated_arm.xml", None, None, IMPLICIT, 50, 40, 0.2, 8),
    ("bin.xml", None, None, IMPLICIT, 100, 6, 0.02, 2),              # 20 free bodies: nv = 120 dense LU
]


@pytest.mark.parametrize("name,solver,cone,integ,settle,nchk,amp,nenv", MATRIX)
def test_solver_cone_integrator_matrix(name, solver, cone, integ, settle, nchk, amp, nenv, capi, orc, BatchSim):
    model = variant(capi, name, solver, cone, integ)
    tag = f"{name}[sol{model.opt.solver} cone{model.opt.cone} int{model.opt.integrator}]"
    qpos, qvel = perturbed(model, nenv, 41, amp)
    rng = np.random.default_rng(17)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    # drive the batch into its contact regime with piecewise-constant random controls
    for s in range(settle):
        if model.nu and s % 25 == 0:
            sim.set("ctrl", ctrl_sample(model, rng, nenv))
        sim.step(1)
    assert np.all(np.isfinite(sim.get("qpos")))
    oracles = make_oracles(orc, model, sim.get("qpos"), sim.get("qvel"))
    # CG stops on its own tolerance (1e-8 scaled) or at the 100-iteration cap without having converged; either way
    # its iterate is sensitive to summation order, so both sides agree only to about the solver's own accuracy
    # (measured 1e-13 .. 1e-5 across these cases); PGS and Newton cases use the north-star tolerance unchanged.
    cg = model.opt.solver == CG
    worst, max_nefc = injected_steps(model, sim, oracles, nchk, rng, tol=1e-4 if cg else TOL, tag=tag)
    # all fields after a forward pass from the state the batch has reached
    st = {k: sim.get(k) for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "ctrl") if model.field_size_by_name(k) > 0}
    sim.keep_intermediates(True)
    sim.forward()
    for e, o in enumerate(oracles):
        for k, v in st.items():
            o.set(k, v[e])
        o.forward()
    wf = compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tol=1e-6 if cg else 1e-8, tag=tag)
    print(f"{tag}: injected-step worst {worst:.2e}, forward fields worst {wf:.2e}, max nefc {max_nefc}")


IMPLICIT_ARM = """
<mujoco>
  <compiler angle="radian"/>
  <option timestep="0.002" integrator="{integ}" gravity="0 0 -9.81"><flag contact="disable"/></option>
  <worldbody>
    <body pos="0 0 1">
      <joint name="j1" type="hinge" axis="0 1 0" damping="0.7"/>
      <geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.04" density="800"/>
      <body pos="0.4 0 0">
        <joint name="j2" type="hinge" axis="0 0 1" damping="0.3"/>
        <geom type="capsule" fromto="0 0 0 0.3 0.1 0" size="0.03" density="800"/>
        <body pos="0.3 0.1 0">
          <joint name="j3" type="ball" damping="0.05"/>
          <geom type="box" size="0.05 0.08 0.03" pos="0.05 0 0.02" density="900"/>
        </body>
      </body>
    </body>
    <body pos="0 1 1">
      <freejoint/>
      <geom type="box" size="0.1 0.2 0.05" density="500"/>
      <body pos="0.2 0 0">
        <joint name="k1" type="slide" axis="1 0 0" damping="0.2"/>
        <geom type="sphere" size="0.05" pos="0.1 0.05 0" density="700"/>
      </body>
    </body>
  </worldbody>
  <actuator>
    <velocity joint="j1" kv="3.5"/>
    <position joint="j2" kp="20" kv="1.5"/>
    <general joint="k1" gaintype="affine" gainprm="2 0 -0.8" biastype="affine" biasprm="0 -1 -0.4"/>
  </actuator>
</mujoco>
"""


@pytest.mark.parametrize("integ", ["implicit", "implicitfast"])
def test_implicit_integrators_free_running(integ, capi, orc, BatchSim):
    """Ball + free + slide joints with velocity-dependent actuators (every term of mjd_smooth_vel), contact free, so a
    300-step free-running comparison is meaningful; also through the split step (control hook between the halves)."""
    model = capi.Model.from_xml_string(IMPLICIT_ARM.format(integ=integ))
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 9, 0.3)
    qvel *= 10
    rng = np.random.default_rng(2)
    sim, split = BatchSim(model, nenv), BatchSim(model, nenv)
    for s_ in (sim, split):
        s_.set("qpos", qpos)
        s_.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)
    worst = 0.0
    for s in range(300):
        if s % 20 == 0:
            ctrl = rng.uniform(-1, 1, (nenv, model.nu))
        sim.set("ctrl", ctrl)
        sim.step(1)
        split.step_begin()
        split.set("ctrl", ctrl)
        split.step_end()
        for e, o in enumerate(oracles):
            o.set("ctrl", ctrl[e])
            o.step(1)
        if s % 25 == 24:
            for k in ("qpos", "qvel"):
                g, g2 = sim.get(k), split.get(k)
                np.testing.assert_array_equal(g, g2)
                for e, o in enumerate(oracles):
                    worst = max(worst, rel(g[e], o.get(k)))
            assert worst < 1e-8, (s, worst)
    print(f"{integ}: free-running 300 steps worst {worst:.2e}")


@pytest.mark.parametrize("name,integ", [("humanoid_like.xml", EULER), ("hand_like.xml", RK4), ("bin.xml", EULER)])
def test_fluid_forces(name, integ, capi, orc, BatchSim):
    """option density / viscosity / wind (viewer.cpp:597-600): inertia-box fluid model in the passive stage."""
    model = variant(capi, name, integrator=integ)
    model.opt.density, model.opt.viscosity = 40.0, 0.9
    model.opt.wind[0], model.opt.wind[1], model.opt.wind[2] = 1.5, -0.7, 0.3
    nenv = 4
    qpos, qvel = perturbed(model, nenv, 11, 0.05)
    qvel *= 20
    sim = BatchSim(model, nenv)
    sim.keep_intermediates(True)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)
    sim.forward()
    for o in oracles:
        o.forward()
    gp = sim.get("qfrc_passive")
    for e, o in enumerate(oracles):
        assert np.max(np.abs(o.get("qfrc_passive"))) > 1e-3
        assert rel(gp[e], o.get("qfrc_passive")) < 1e-11, (e, rel(gp[e], o.get("qfrc_passive")))
    rng = np.random.default_rng(5)
    worst, _ = injected_steps(model, sim, oracles, 30, rng, tag=f"fluid:{name}")
    print(f"fluid {name}: injected-step worst {worst:.2e}")


NOSLIP = [
    ("panda_like.xml", PGS, PYR, 450, 30, 0.1, 16),
    ("panda_like.xml", PGS, ELL, 450, 30, 0.1, 16),
    ("panda_like.xml", NEWTON, ELL, 450, 30, 0.1, 16),
    ("box_stack.xml", PGS, PYR, 150, 30, 0.1, 8),
    ("box_stack.xml", NEWTON, ELL, 150, 30, 0.1, 8),
    ("box_stack.xml", CG, PYR, 150, 30, 0.1, 8),
    ("humanoid_like.xml", NEWTON, PYR, 120, 25, 0.02, 8),      # sparse L'DL path, friction-loss rows
    ("humanoid_like.xml", PGS, ELL, 120, 25, 0.02, 8),
    ("hand_like.xml", NEWTON, ELL, 100, 15, 0.02, 8),           # RK4: the noslip pass runs in every sub-step
    ("bin.xml", NEWTON, ELL, 100, 6, 0.02, 2),                  # nv = 120: single-warp path instead of the team
]


@pytest.mark.parametrize("name,solver,cone,settle,nchk,amp,nenv", NOSLIP)
def test_noslip_pass(name, solver, cone, settle, nchk, amp, nenv, capi, orc, BatchSim):
    """opt.noslip_iterations > 0 (viewer.cpp:590-591): the unregularised friction sweep after the main solver."""
    model = variant(capi, name, solver, cone)
    model.opt.noslip_iterations = 4
    model.opt.noslip_tolerance = 1e-12
    tag = f"noslip:{name}[sol{solver} cone{cone}]"
    qpos, qvel = perturbed(model, nenv, 41, amp)
    rng = np.random.default_rng(17)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    for s in range(settle):
        if model.nu and s % 25 == 0:
            sim.set("ctrl", ctrl_sample(model, rng, nenv))
        sim.step(1)
    assert np.all(np.isfinite(sim.get("qpos")))
    oracles = make_oracles(orc, model, sim.get("qpos"), sim.get("qvel"))
    worst, max_nefc = injected_steps(model, sim, oracles, nchk, rng, tol=1e-4 if solver == CG else TOL, tag=tag)
    # the pass must have run: solver_iter counts main + noslip iterations on both sides
    st = {k: sim.get(k) for k in STATE if model.field_size_by_name(k) > 0}
    sim.keep_intermediates(True)
    sim.forward()
    it = sim.get("solver_iter")[:, 0]
    ff = sim.get("efc_force")
    nefc = sim.get("nefc")[:, 0]
    for e, o in enumerate(oracles):
        for k, v in st.items():
            o.set(k, v[e])
        o.forward()
        if solver != CG:
            assert int(o.get("solver_iter")[0]) == int(it[e]), (e, it[e], o.get("solver_iter"))
        n = int(nefc[e])
        assert rel(ff[e][:n], o.get("efc_force")[:n]) < (1e-3 if solver == CG else 1e-6), (e, n)
    print(f"{tag}: injected-step worst {worst:.2e}, max nefc {max_nefc}")


CONVEX_PILE = """
<mujoco>
  <option timestep="0.002" solver="Newton" cone="elliptic"/>
  <worldbody>
    <geom type="plane" size="3 3 0.1"/>
    <geom type="box" size="0.3 0.3 0.05" pos="0 0 0.05"/>
    <geom type="cylinder" size="0.15 0.1" pos="0.6 0 0.1"/>
    <body pos="0 0 0.4"><freejoint/><geom type="cylinder" size="0.08 0.12" density="600"/></body>
    <body pos="0.05 0.02 0.75" euler="40 20 0"><freejoint/><geom type="ellipsoid" size="0.07 0.1 0.13" density="600"/></body>
    <body pos="-0.1 0.05 1.1" euler="0 80 10"><freejoint/><geom type="cylinder" size="0.06 0.15" density="600"/></body>
    <body pos="0.6 0.02 0.5" euler="10 10 0"><freejoint/><geom type="ellipsoid" size="0.1 0.08 0.06" density="600"/></body>
    <body pos="0.62 -0.03 0.8"><freejoint/><geom type="sphere" size="0.07" density="600"/></body>
    <body pos="0.0 0.5 0.3" euler="0 90 0"><freejoint/><geom type="capsule" size="0.05 0.12" density="600"/></body>
    <body pos="0.02 0.52 0.6" euler="30 0 30"><freejoint/><geom type="ellipsoid" size="0.09 0.06 0.05" density="600"/></body>
    <body pos="-0.5 -0.5 0.3" euler="20 30 0"><freejoint/><geom type="box" size="0.08 0.06 0.05" density="600"/></body>
    <body pos="-0.52 -0.48 0.6" euler="70 0 0"><freejoint/><geom type="cylinder" size="0.05 0.1" density="600"/></body>
  </worldbody>
</mujoco>
"""


def test_convex_pairs(capi, orc, BatchSim):
    """Cylinders and ellipsoids against planes, boxes, spheres, capsules and each other: plane-cylinder, plane-convex
    and the MPR test (stages_convex.cuh).  Both sides run the same portal refinement, so contacts agree to round-off
    (measured: distance 1e-15, normal 4e-12, per-step state 2e-9); the bounds below leave room for an iterate that lands
    on the other side of a branch under FMA contraction, which would end one refinement (<= mpr_tolerance) apart."""
    model = capi.Model.from_xml_string(CONVEX_PILE)
    nenv = 6
    qpos, qvel = perturbed(model, nenv, 23, 0.02)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    for _ in range(6):
        sim.step(50)
        assert np.all(np.isfinite(sim.get("qpos")))
    oracles = make_oracles(orc, model, sim.get("qpos"), sim.get("qvel"))
    st = {k: sim.get(k) for k in STATE if model.field_size_by_name(k) > 0}
    sim.keep_intermediates(True)
    sim.forward()
    ncon = sim.get("ncon")[:, 0]
    g1, g2 = sim.get("contact_geom1"), sim.get("contact_geom2")
    dist, pos, frame = sim.get("contact_dist"), sim.get("contact_pos"), sim.get("contact_frame")
    seen = set()
    worst_d = worst_n = 0.0
    for e, o in enumerate(oracles):
        for k, v in st.items():
            o.set(k, v[e])
        o.forward()
        n = int(o.get("ncon")[0])
        assert n == int(ncon[e]), (e, n, ncon[e])
        np.testing.assert_array_equal(g1[e][:n], o.get("contact_geom1")[:n])
        np.testing.assert_array_equal(g2[e][:n], o.get("contact_geom2")[:n])
        for c in range(n):
            seen.add((int(model.geom_type[g1[e][c]]), int(model.geom_type[g2[e][c]])))
        worst_d = max(worst_d, float(np.max(np.abs(dist[e][:n] - o.get("contact_dist")[:n]))))
        worst_n = max(worst_n, float(np.max(np.abs(frame[e][:9 * n].reshape(n, 9)[:, :3] - o.get("contact_frame")[:9 * n].reshape(n, 9)[:, :3]))))
        assert np.max(np.abs(pos[e][:3 * n] - o.get("contact_pos")[:3 * n])) < 5e-3
    assert worst_d < 2e-6 and worst_n < 1e-3, (worst_d, worst_n)
    assert (0, 5) in seen and (0, 4) in seen and len([p for p in seen if 4 in p or 5 in p]) >= 5, seen
    rng = np.random.default_rng(1)
    sim.keep_intermediates(False)
    worst, _ = injected_steps(model, sim, oracles, 40, rng, tol=TOL, tag="convex pile")
    print(f"convex pile: pair types {sorted(seen)}, contact dist worst {worst_d:.1e}, normal worst {worst_n:.1e}, step worst {worst:.1e}")


def test_mesh_geoms(capi, orc, BatchSim):
    """Convex meshes (hull vertices, stages_convex.cuh): plane-mesh multi-point contacts and MPR against boxes,
    cylinders, spheres and other meshes."""
    import itertools
    cube = " ".join(f"{x} {y} {z}" for x, y, z in itertools.product((-0.08, 0.08), (-0.1, 0.1), (-0.06, 0.06)))
    rng0 = np.random.default_rng(4)
    blob = rng0.normal(size=(60, 3))
    blob = blob / np.linalg.norm(blob, axis=1, keepdims=True) * [0.09, 0.07, 0.11]
    blob_s = " ".join(f"{x:.17g}" for x in blob.ravel())
    xml = f"""<mujoco><option timestep="0.002" solver="Newton"/>
      <asset><mesh name="cube" vertex="{cube}"/><mesh name="tet" vertex="0 0 0 0.2 0 0 0 0.2 0 0 0 0.2"/>
             <mesh name="blob" vertex="{blob_s}"/></asset>
      <worldbody>
        <geom type="plane" size="3 3 0.1"/>
        <geom type="box" size="0.3 0.3 0.05" pos="0 0 0.05"/>
        <geom type="mesh" mesh="blob" pos="0.7 0 0.1"/>
        <body pos="0 0 0.3"><freejoint/><geom type="mesh" mesh="cube" density="600"/></body>
        <body pos="0.03 0.02 0.6" euler="20 30 0"><freejoint/><geom type="mesh" mesh="tet" density="600"/></body>
        <body pos="0.7 0.02 0.45" euler="10 0 40"><freejoint/><geom type="mesh" mesh="cube" density="600"/></body>
        <body pos="-0.6 0 0.3" euler="50 20 0"><freejoint/><geom type="mesh" mesh="blob" density="600"/></body>
        <body pos="-0.58 0.03 0.6"><freejoint/><geom type="cylinder" size="0.06 0.08" density="600"/></body>
        <body pos="0 0.7 0.3" euler="0 45 0"><freejoint/><geom type="mesh" mesh="tet" density="600"/></body>
        <body pos="0.02 0.72 0.55"><freejoint/><geom type="sphere" size="0.07" density="600"/></body>
      </worldbody></mujoco>"""
    model = capi.Model.from_xml_string(xml)
    nenv = 6
    qpos, qvel = perturbed(model, nenv, 29, 0.02)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)
    seen = set()
    for chunk in range(6):   # the piles topple over time: compare the contact sets at several moments
        sim.keep_intermediates(False)
        sim.step(50)
        assert np.all(np.isfinite(sim.get("qpos")))
        st = {k: sim.get(k) for k in STATE if model.field_size_by_name(k) > 0}
        sim.keep_intermediates(True)
        sim.forward()
        ncon = sim.get("ncon")[:, 0]
        g1, g2, dist = sim.get("contact_geom1"), sim.get("contact_geom2"), sim.get("contact_dist")
        for e, o in enumerate(oracles):
            for k, v in st.items():
                o.set(k, v[e])
            o.forward()
            n = int(o.get("ncon")[0])
            assert n == int(ncon[e]), (chunk, e, n, ncon[e])
            np.testing.assert_array_equal(g1[e][:n], o.get("contact_geom1")[:n])
            np.testing.assert_array_equal(g2[e][:n], o.get("contact_geom2")[:n])
            if n:
                assert np.max(np.abs(dist[e][:n] - o.get("contact_dist")[:n])) < 2e-6
            for c in range(n):
                seen.add((int(model.geom_type[g1[e][c]]), int(model.geom_type[g2[e][c]])))
    assert (0, 7) in seen and (7, 7) in seen and (6, 7) in seen, seen
    sim.keep_intermediates(False)
    worst, _ = injected_steps(model, sim, oracles, 40, np.random.default_rng(1), tol=TOL, tag="mesh pile")
    print(f"mesh pile: pair types {sorted(seen)}, step worst {worst:.1e}")


SPATIAL = """
<mujoco>
  <compiler angle="radian"/>
  <option timestep="0.002" integrator="{integ}"/>
  <worldbody>
    <geom type="plane" size="3 3 0.1"/>
    <site name="anchor" pos="0 0 1.5"/>
    <site name="anchor2" pos="0.5 0 1.5"/>
    <body name="l1" pos="0 0 1">
      <joint name="j1" type="hinge" axis="0 1 0" damping="0.05"/>
      <geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.03"/>
      <site name="mid" pos="0.25 0 0.05"/>
      <body name="l2" pos="0.5 0 0">
        <joint name="j2" type="ball" damping="0.02"/>
        <joint name="j3" type="slide" axis="1 0 0" range="-0.1 0.1" damping="1"/>
        <geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.03"/>
        <site name="tip" pos="0.4 0 0.02"/>
        <site name="tip2" pos="0.2 0.05 0"/>
      </body>
    </body>
    <body pos="0.3 0.6 0.8"><freejoint/><geom type="box" size="0.05 0.05 0.05"/><site name="boxtop" pos="0 0 0.05"/></body>
  </worldbody>
  <tendon>
    <spatial name="cable" stiffness="40" damping="1.5" limited="true" range="0 1.05"><site site="anchor"/><site site="mid"/><site site="tip"/></spatial>
    <spatial name="block" frictionloss="0.3">
      <site site="anchor2"/><site site="tip"/><pulley divisor="2"/><site site="anchor2"/><site site="tip2"/>
      <pulley divisor="2"/><site site="mid"/><site site="tip2"/>
    </spatial>
    <spatial name="leash" stiffness="200" springlength="0 0.6"><site site="anchor2"/><site site="boxtop"/></spatial>
  </tendon>
  <actuator><motor name="pull" tendon="cable" gear="3" ctrlrange="-2 2"/><position tendon="block" kp="20" kv="0.5" ctrlrange="0.5 1.5"/></actuator>
  <sensor><tendonpos tendon="cable"/><tendonvel tendon="block"/><tendonlimitfrc tendon="cable"/><actuatorfrc actuator="pull"/></sensor>
</mujoco>
"""


@pytest.mark.parametrize("integ", ["Euler", "RK4", "implicitfast"])
def test_spatial_tendons(integ, capi, orc, BatchSim):
    """Site paths with pulleys driving springs, dampers, limits, friction loss, actuators and sensors."""
    model = capi.Model.from_xml_string(SPATIAL.format(integ=integ))
    assert model.ntendon == 3 and model.nwrap == 13
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 13, 0.2)
    sim = BatchSim(model, nenv)
    sim.keep_intermediates(True)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)
    sim.forward()
    for o in oracles:
        o.forward()
    wf = compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tol=1e-9, tag="spatial")
    sim.keep_intermediates(False)
    worst, max_nefc = injected_steps(model, sim, oracles, 150, np.random.default_rng(3), tag=f"spatial:{integ}")
    assert max_nefc >= 2
    print(f"spatial tendons [{integ}]: forward fields worst {wf:.1e}, injected-step worst {worst:.1e}, max nefc {max_nefc}")


RANGE_SCENE = """
<mujoco>
  <option timestep="0.002"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1"/>
    <geom type="sphere" size="0.25" pos="0.6 0 0.3"/>
    <geom type="capsule" size="0.1 0.3" pos="0 0.7 0.4" euler="0 70 20"/>
    <geom type="cylinder" size="0.2 0.4" pos="-0.7 0 0.4" euler="30 0 0" contype="0" conaffinity="0"/>
    <geom type="ellipsoid" size="0.1 0.2 0.3" pos="0 -0.7 0.5" euler="0 40 0" contype="0" conaffinity="0"/>
    <geom type="box" size="0.2 0.3 0.1" pos="0.5 0.5 0.8" euler="20 30 40"/>
    <geom type="box" size="0.5 0.5 0.01" pos="0 0 0.2" rgba="1 0 0 0" contype="0" conaffinity="0"/>
    <body name="probe" pos="0 0 1.2">
      <freejoint/>
      <geom type="sphere" size="0.08"/>
      <site name="s0" pos="0 0 0" euler="180 0 0"/>
      <site name="s1" pos="0.02 0 0" euler="120 0 0"/>
      <site name="s2" pos="0 0.02 0" euler="0 120 0"/>
      <site name="s3" pos="0 0 0.02" euler="200 30 0"/>
      <site name="s4" pos="0 0 0" euler="0 0 0"/>
      <site name="s5" pos="0 0 0" euler="150 40 10"/>
    </body>
  </worldbody>
  <sensor>
    <rangefinder site="s0"/><rangefinder site="s1"/><rangefinder site="s2"/><rangefinder site="s3" cutoff="0.9"/>
    <rangefinder site="s4"/><rangefinder site="s5"/>
  </sensor>
</mujoco>
"""


def test_rangefinder(capi, orc, BatchSim):
    """mjSENS_RANGEFINDER (the sensor plugin publishes it as a scalar, mujoco_sensor_handler_plugin.cpp:331): rays from a
    tumbling free body against every primitive geom type, transparent geoms and the site's own body excluded."""
    model = capi.Model.from_xml_string(RANGE_SCENE)
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 3, 0.3)
    qvel *= 15
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)
    hits = 0
    for s in range(150):
        sim.step(1)
        for o in oracles:
            o.step(1)
        if s % 10 == 9:
            g = sim.get("sensordata")
            for e, o in enumerate(oracles):
                want = o.get("sensordata")
                assert np.array_equal(g[e] < 0, want < 0), (s, e, g[e], want)
                assert rel(g[e], want) < 1e-9, (s, e, g[e], want)
                hits += int(np.sum(want >= 0))
    assert hits > 100


@pytest.mark.parametrize("integ", [EULER, RK4])
def test_activation_dynamics(integ, capi, orc, BatchSim):
    """na > 0: integrator / filter activation dynamics, actlimited clamp, affine gain + bias, forcerange, tendon
    transmission -- act and act_dot against the oracle, free-running for 400 steps (smooth dynamics, no chaos)."""
    model = variant(capi, "actuated_arm.xml", integrator=integ)
    assert model.na == 4
    nenv = 16
    qpos, qvel = perturbed(model, nenv, 5, 0.2)
    rng = np.random.default_rng(3)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    act0 = rng.uniform(-0.3, 0.3, (nenv, model.na))
    sim.set("act", act0)
    oracles = make_oracles(orc, model, qpos, qvel)
    for e, o in enumerate(oracles):
        o.set("act", act0[e])
    worst = 0.0
    for s in range(400):
        if s % 10 == 0:
            ctrl = ctrl_sample(model, rng, nenv) * 1.5   # beyond ctrlrange: exercises the ctrl clamp
        sim.set("ctrl", ctrl)
        sim.step(1)
        for e, o in enumerate(oracles):
            o.set("ctrl", ctrl[e])
            o.step(1)
        if s % 20 == 19:
            for k in ("qpos", "qvel", "act", "act_dot", "sensordata"):
                g = sim.get(k)
                for e, o in enumerate(oracles):
                    worst = max(worst, rel(g[e], o.get(k)))
            assert worst < TOL, (s, worst)
    assert np.abs(sim.get("act")).max() > 0.05
    # forward fields incl. actuator_force / qfrc_actuator / actuator_moment
    sim.keep_intermediates(True)
    sim.forward()
    for o in oracles:
        o.forward()
    compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tag=f"actuated_arm int{integ}")


def test_xfrc_applied_matches_oracle(capi, orc, BatchSim):
    """Cartesian wrenches on bodies (mjData.xfrc_applied, written by plugins inside controlCallback,
    plugin_utils.h:89-95) enter qfrc_smooth through J'; also the accelerometer path of the sensors."""
    model = variant(capi, "humanoid_like.xml")
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 9, 0.02)
    rng = np.random.default_rng(12)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)

    def wrench(s):
        return {"xfrc_applied": rng.uniform(-20, 20, (nenv, 6 * model.nbody))}

    worst, _ = injected_steps(model, sim, oracles, 40, rng, tag="humanoid xfrc", extra_inputs=wrench)
    x = rng.uniform(-20, 20, (nenv, 6 * model.nbody))
    sim.set("xfrc_applied", x)
    st = {k: sim.get(k) for k in ("qpos", "qvel", "qacc_warmstart", "time", "ctrl")}
    sim.keep_intermediates(True)
    sim.forward()
    for e, o in enumerate(oracles):
        for k, v in st.items():
            o.set(k, v[e])
        o.set("xfrc_applied", x[e])
        o.forward()
    compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tag="humanoid xfrc")
    # and the wrench really acts: qacc differs from the wrench-free forward pass
    qa = sim.get("qacc")
    sim.set("xfrc_applied", np.zeros_like(x))
    sim.forward()
    assert rel(qa, sim.get("qacc")) > 1e-3
    print(f"xfrc injected-step worst {worst:.2e}")


def ref_fixture():
    return np.load(os.path.join(GOLDEN, "ref_models.npz"))


@pytest.mark.parametrize("name", ["pendulum_world", "empty_world", "equality_world", "mocap_world", "sensors_world"])
def test_reference_xml_verbatim(name, capi, orc, BatchSim):
    """The reference's own MJCF files (bytes frozen by tools/make_ref_model_fixtures.py), compiled verbatim: forward
    fields at qpos0 against a live oracle and the 200-step trajectory against the frozen one."""
    g = ref_fixture()
    model = capi.Model.from_xml_string(bytes(g[f"{name}__xml"]).decode())
    nenv = 3
    sim = BatchSim(model, nenv)
    sim.keep_intermediates(True)
    sim.forward()
    oracles = [orc.Oracle(model) for _ in range(nenv)]
    for o in oracles:
        o.forward()
    compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tag=name)
    sim.keep_intermediates(False)
    tq, tv = g[f"{name}__qpos"], g[f"{name}__qvel"]
    for k in range(tq.shape[0]):
        sim.step(1)
        if k % 20 == 19 or k == 0:
            assert rel(sim.get("qpos"), np.tile(tq[k], (nenv, 1))) < TOL, (name, k)
            assert rel(sim.get("qvel"), np.tile(tv[k], (nenv, 1))) < TOL, (name, k)
    np.testing.assert_array_equal(sim.get("time")[:, 0], g[f"{name}__time"][0])


def test_mocap_bodies_drag_welded_box(capi, orc, BatchSim):
    """mocap_world.xml (mujoco_ros_mocap_plugin/assets): the plugin writes d->mocap_pos / mocap_quat every control
    callback (mocap_plugin.cpp); a weld ties a free box to mocap2.  Per-env mocap trajectories, injected-step parity."""
    g = ref_fixture()
    model = capi.Model.from_xml_string(bytes(g["mocap_world__xml"]).decode())
    assert model.nmocap == 2
    nenv = 6
    sim = BatchSim(model, nenv)
    oracles = [orc.Oracle(model) for _ in range(nenv)]
    rng = np.random.default_rng(4)
    base_pos = sim.get("mocap_pos").copy()
    phase = rng.uniform(0, 6.28, (nenv, 1))

    def move(s):
        p = base_pos.copy()
        p[:, 3:4] += 0.2 * np.sin(0.02 * s + phase)        # mocap2 x
        p[:, 5:6] += 0.1 * (1 - np.cos(0.015 * s + phase))  # mocap2 z
        q = np.tile([1.0, 0, 0, 0, 1.0, 0, 0, 0], (nenv, 1))
        ang = 0.3 * np.sin(0.01 * s + phase[:, 0])
        q[:, 4], q[:, 7] = np.cos(ang / 2), np.sin(ang / 2)
        return {"mocap_pos": p, "mocap_quat": q}

    worst, max_nefc = injected_steps(model, sim, oracles, 150, rng, tag="mocap_world", extra_inputs=move, check_every=3)
    q = sim.get("qpos")
    assert np.abs(q[:, 0] - 0.3).max() > 0.02, "the welded box should have been dragged"
    assert max_nefc >= 6
    print(f"mocap injected-step worst {worst:.2e}, max nefc {max_nefc}")


def test_bad_qvel_and_qacc_trigger_reset(capi, BatchSim):
    """mj_checkVel / mj_checkAcc: a non-finite or huge value resets THAT env (mj_resetData) and bumps its warning
    counter; the others are untouched."""
    model = variant(capi, "pendulum_scene.xml")
    nenv = 5
    sim = BatchSim(model, nenv)
    v = np.zeros((nenv, model.nv))
    v[1, 2] = np.inf
    v[3, 0] = 2e10          # > mjMAXVAL
    sim.set("qvel", v)
    sim.step(1)
    w = sim.get("warning")
    assert w[1, 5] == 1 and w[3, 5] == 1 and w[[0, 2, 4]].sum() == 0   # B2MJ_WARN_BADQVEL = 5
    assert np.all(np.isfinite(sim.get("qvel")))
    # BADQACC: an absurd applied force makes qacc exceed mjMAXVAL on one env
    sim.reset()
    f = np.zeros((nenv, model.nv))
    f[2, 5] = 1e14
    sim.set("qfrc_applied", f)
    sim.step(1)
    w = sim.get("warning")
    assert w[2, 6] == 1 and w[[0, 1, 3, 4]].sum() == 0                  # B2MJ_WARN_BADQACC = 6
    t = sim.get("time")[:, 0]
    assert t[2] == model.opt.timestep   # reset to t = 0, then the step is taken from the reset state
    q = sim.get("qpos")
    assert np.all(np.isfinite(q)) and np.all(np.isfinite(sim.get("qacc")))


@pytest.mark.parametrize("name,solver", [("panda_like.xml", PGS), ("panda_like.xml", NEWTON), ("humanoid_like.xml", NEWTON),
                                         ("humanoid_like.xml", CG), ("box_stack.xml", PGS), ("bin.xml", NEWTON)])
def test_efc_state_matches_oracle(name, solver, capi, orc, BatchSim):
    """efc_state (satisfied / quadratic / linear / cone per row) after the solve, integer-exact."""
    model = variant(capi, name, solver)
    nenv = 8 if name != "bin.xml" else 4
    qpos, qvel = perturbed(model, nenv, 23, 0.02 if name != "panda_like.xml" else 0.1)
    rng = np.random.default_rng(8)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    for s in range(200 if name != "bin.xml" else 100):
        if model.nu and s % 25 == 0:
            sim.set("ctrl", ctrl_sample(model, rng, nenv))
        sim.step(1)
    st = {k: sim.get(k) for k in ("qpos", "qvel", "qacc_warmstart", "time", "ctrl") if model.field_size_by_name(k) > 0}
    sim.keep_intermediates(True)
    sim.forward()
    nefc = sim.get("nefc")[:, 0]
    gs = sim.get("efc_state")
    total = 0
    for e in range(nenv):
        o = orc.Oracle(model)
        for k, v in st.items():
            o.set(k, v[e])
        o.forward()
        assert o.get("nefc")[0] == nefc[e]
        np.testing.assert_array_equal(gs[e][:nefc[e]], o.get("efc_state")[:nefc[e]], err_msg=f"{name} env {e}")
        total += int(nefc[e])
    assert total > 0


def test_damping_switched_on_by_model_update(capi, orc, BatchSim):
    """a handle created with zero joint damping must step correctly after b2mj_model_update turns damping
    on (the implicit-damping inverse needs arena space that is reserved at create time)."""
    from conftest import model_path

    model = capi.Model.from_xml_file(model_path("panda_like.xml"))
    model.dof_damping[:] = 0
    nenv = 8
    qpos, qvel = perturbed(model, nenv, 2, 0.1)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    sim.step(20)
    model.dof_damping[:] = 2.0
    sim.model_update(model)
    st = {k: sim.get(k) for k in ("qpos", "qvel", "qacc_warmstart", "time")}
    sim.step(50)
    gq, gv = sim.get("qpos"), sim.get("qvel")
    for e in range(nenv):
        o = orc.Oracle(model)
        for k, v in st.items():
            o.set(k, v[e])
        o.step(50)
        assert rel(gq[e], o.get("qpos")) < TOL and rel(gv[e], o.get("qvel")) < TOL


def test_reset_abandons_split_step(capi, BatchSim):
    model = variant(capi, "pendulum_scene.xml")
    sim = BatchSim(model, 2)
    sim.step_begin()
    sim.reset()
    with pytest.raises(capi.B2mjError):
        sim.step_end()


def test_unsupported_options_are_rejected(capi, BatchSim):
    """What the step kernel does not implement must be refused, not ignored: fluid forces under an implicit integrator
    (their velocity derivatives are missing from qDeriv), unknown integrators / solvers."""
    for integ in (IMPLICIT, IMPLICITFAST):
        m = variant(capi, "pendulum_scene.xml", integrator=integ)
        m.opt.density = 1.2
        with pytest.raises(capi.B2mjError, match="fluid"):
            BatchSim(m, 2)
    m = variant(capi, "pendulum_scene.xml")
    m.opt.integrator = 7
    with pytest.raises(capi.B2mjError, match="integrator"):
        BatchSim(m, 2)
    m = variant(capi, "pendulum_scene.xml")
    m.opt.solver = 5
    with pytest.raises(capi.B2mjError, match="solver"):
        BatchSim(m, 2)
    m = variant(capi, "pendulum_scene.xml")   # fluid forces with Euler are supported (test_fluid_forces)
    m.opt.density = 1.2
    BatchSim(m, 2).step(1)


@pytest.mark.parametrize("name,nenv,amp", [("humanoid_like.xml", 6, 0.02), ("hand_like.xml", 6, 0.02), ("bin.xml", 3, 0.02)])
def test_thousand_step_per_step_parity(name, nenv, amp, capi, orc, BatchSim):
    """North star: "<1e-5 rel per step ... over 1000 steps".  The contact-rich configs are chaotic (a free-running
    1e-13 difference grows past 1e-5 within a few hundred steps on either side), so the 1000-step claim is checked the
    way the north star words it -- per step: every step starts from the batch's own state on both sides.  The
    free-running divergence of a second, never-corrected set of oracles is printed for the record."""
    model = variant(capi, name)
    qpos, qvel = perturbed(model, nenv, 61, amp)
    rng = np.random.default_rng(29)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)
    free = make_oracles(orc, model, qpos, qvel)
    worst, max_nefc, div = 0.0, 0, []
    rng_free = np.random.default_rng(29)
    for block in range(10):
        # identical control streams for the injected and the free-running oracles
        state = rng.bit_generator.state
        w, mn = injected_steps(model, sim, oracles, 100, rng, tag=name)
        rng_free.bit_generator.state = state
        for s in range(100):
            ctrl = ctrl_sample(model, rng_free, nenv)
            for e, o in enumerate(free):
                if model.nu:
                    o.set("ctrl", ctrl[e])
                o.step(1)
        worst, max_nefc = max(worst, w), max(max_nefc, mn)
        gq = sim.get("qpos")
        div.append(max(rel(gq[e], o.get("qpos")) for e, o in enumerate(free)))
    np.testing.assert_array_equal(sim.get("time")[:, 0], [o.time for o in free])
    print(f"{name}: 1000 injected steps worst {worst:.2e}, max nefc {max_nefc}; free-running divergence per 100 steps "
          + " ".join(f"{d:.1e}" for d in div))


def test_rk4_split_step_hooks_fire_in_every_substep(capi, orc, BatchSim):
    """mj_RungeKutta makes four forward passes and mjcb_control fires in each (plugin_utils.h:89-97).  The split step
    yields to the host four times per RK4 step (b2mj_step_end -> B2MJ_AGAIN); with a controller that reads the
    sub-step state (time and qvel) the batch must follow an oracle whose control callback does the same, and with a
    constant control the split step must equal the fused one bitwise."""
    model = variant(capi, "actuated_arm.xml", integrator=RK4)
    nenv = 6
    qpos, qvel = perturbed(model, nenv, 15, 0.2)
    a, b = BatchSim(model, nenv), BatchSim(model, nenv)
    for s_ in (a, b):
        s_.set("qpos", qpos)
        s_.set("qvel", qvel)
    ctrl = np.random.default_rng(1).uniform(-1, 1, (nenv, model.nu))
    a.set("ctrl", ctrl)
    b.set("ctrl", ctrl)
    for _ in range(25):
        a.step(1)
        b.step_begin()
        n = 1
        while b.step_end() == 1:
            n += 1
        assert n == 4
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))
    np.testing.assert_array_equal(a.get("act"), b.get("act"))
    np.testing.assert_array_equal(a.get("time"), b.get("time"))

    # state-dependent controller evaluated inside every sub-step
    def controller(time, qv):
        return np.stack([0.8 * np.sin(40.0 * time + k) - 0.5 * qv[:, k % qv.shape[1]] for k in range(model.nu)], axis=1)

    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    oracles = make_oracles(orc, model, qpos, qvel)
    calls = [0]

    def cb(o):
        calls[0] += 1
        c = controller(np.array([o.time]), o.get("qvel")[None, :])[0]
        o.set("ctrl", c)

    for o in oracles:
        o.set_callbacks(control=cb)
    for _ in range(60):
        sim.step_begin()
        while True:
            sim.set("ctrl", controller(sim.get("time")[:, 0], sim.get("qvel")))
            if sim.step_end() != 1:
                break
        for o in oracles:
            o.step(1)
    assert calls[0] == 60 * 4 * nenv
    gq, gv = sim.get("qpos"), sim.get("qvel")
    for e, o in enumerate(oracles):
        assert rel(gq[e], o.get("qpos")) < TOL and rel(gv[e], o.get("qvel")) < TOL, e
    np.testing.assert_array_equal(sim.get("time")[:, 0], [o.time for o in oracles])


def test_collision_function_override_table(capi, BatchSim):
    """Batched counterpart of MujocoEnv::registerCollisionFunction (mujoco_env.cpp:163-176): overriding the narrowphase
    of a geom-type pair.  NONE on (plane, sphere): the pendulum scene's ball falls through the floor in exact free
    fall; BOUNDING_SPHERES on (plane, box): a box comes to rest one bounding radius above the plane instead of one
    half-height; reset restores the built-in functions."""
    PLANE, SPHERE, BOX = 0, 2, 6
    model = variant(capi, "pendulum_scene.xml")
    nenv = 3
    a, b = BatchSim(model, nenv), BatchSim(model, nenv)
    b.register_collision_function(PLANE, SPHERE, 1)   # B2MJ_COLLFN_NONE
    for s_ in (a, b):
        s_.step(400)
    assert np.all(a.get("ncon")[:, 0] == 1) and np.all(b.get("ncon")[:, 0] == 0)
    ball = model.jnt_qposadr[model.name2id(capi.OBJ_JOINT, "ball_freejoint")]
    dof = model.jnt_dofadr[model.name2id(capi.OBJ_JOINT, "ball_freejoint")]
    g, h, k = 9.81, model.opt.timestep, 400
    np.testing.assert_allclose(b.get("qvel")[:, dof + 2], -g * h * k, rtol=1e-12)
    np.testing.assert_allclose(b.get("qpos")[:, ball + 2], model.qpos0[ball + 2] - g * h * h * k * (k + 1) / 2, rtol=1e-12)
    assert a.get("qpos")[0, ball + 2] > 0.04      # the unmodified batch rests on the floor
    b.reset_collision_functions()
    b.reset()
    b.step(400)
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))
    # bounding spheres for plane-box
    xml = """<mujoco><option timestep="0.002"/><worldbody><geom type="plane" size="2 2 .1"/>
      <body pos="0 0 0.5"><freejoint/><geom type="box" size="0.1 0.1 0.05" density="500"/></body></worldbody></mujoco>"""
    m = capi.Model.from_xml_string(xml)
    c, d = BatchSim(m, 2), BatchSim(m, 2)
    d.register_collision_function(BOX, PLANE, 2)      # argument order does not matter; B2MJ_COLLFN_BOUNDING_SPHERES
    for s_ in (c, d):
        s_.step(1500)
    rb = float(m.geom_rbound[1])
    assert abs(c.get("qpos")[0, 2] - 0.05) < 2e-3
    assert abs(d.get("qpos")[0, 2] - rb) < 2e-3 and rb > 0.14
    with pytest.raises(capi.B2mjError):
        d.register_collision_function(PLANE, PLANE, 2)
