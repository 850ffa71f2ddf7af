"""Batched plugin data paths on the GPU: robot_hw (DefaultRobotHWSim read/write) and sensor readout
(MujocoRosSensorsPlugin::lastStageCallback arithmetic)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def BatchSim():
    from mujoco_ros_pkgs_b200.batch import BatchSim as B

    return B


def test_sensor_readout_is_float32_of_sensordata_over_cutoff(load_model, BatchSim):
    # mujoco_sensors_test.cpp:326-328,440-442,550,631-635: GT == sensordata[adr+k]/cutoff (cutoff<=0 -> 1)
    model = load_model("pendulum_scene.xml")
    sim = BatchSim(model, 8)
    sim.step(60)
    vals, gt = sim.sensor_readout(want_gt=True)
    sd = sim.get("sensordata")
    expect = sd.copy()
    for i in range(model.nsensor):
        c = model.sensor_cutoff[i] if model.sensor_cutoff[i] > 0 else 1.0
        a, d = model.sensor_adr[i], model.sensor_dim[i]
        expect[:, a:a + d] = sd[:, a:a + d] / c
    np.testing.assert_array_equal(gt, expect.astype(np.float32).astype(np.float64))
    np.testing.assert_array_equal(vals, gt)  # no noise model configured


def test_sensor_noise_statistics_seeded(load_model, BatchSim, capi):
    # mujoco_sensors_test.cpp:335-391: mean / variance of the error match the model, un-flagged dims exact
    model = load_model("pendulum_scene.xml")
    nenv = 512
    sim = BatchSim(model, nenv)
    sid = model.name2id(capi.OBJ_SENSOR, "vel_EE")
    sim.sensor_configure_noise([(sid, (0.0, 1.0, 0.0), (0.025, 0.0, 0.0), 0b011)], seed=7)
    sim.step(1)
    errs = []
    for _ in range(8):
        vals, gt = sim.sensor_readout(want_gt=True)
        a = model.sensor_adr[sid]
        errs.append(vals[:, a:a + 3] - gt[:, a:a + 3])
    e = np.concatenate(errs)
    assert abs(e[:, 0].mean()) < 2e-3 and abs(e[:, 0].var() - 0.000625) < 1e-4
    assert abs(e[:, 1].mean() - 1.0) < 1e-6 and e[:, 1].var() < 1e-12
    np.testing.assert_array_equal(e[:, 2], 0)
    # reproducible: same seed, same stream
    sim2 = BatchSim(model, nenv)
    sim2.sensor_configure_noise([(sid, (0.0, 1.0, 0.0), (0.025, 0.0, 0.0), 0b011)], seed=7)
    sim2.step(1)
    v2, _ = sim2.sensor_readout(want_gt=True)
    sim3 = BatchSim(model, nenv)
    sim3.sensor_configure_noise([(sid, (0.0, 1.0, 0.0), (0.025, 0.0, 0.0), 0b011)], seed=7)
    sim3.step(1)
    v3, _ = sim3.sensor_readout(want_gt=True)
    np.testing.assert_array_equal(v2, v3)


MODES = {"EFFORT": 0, "POSITION": 1, "POSITION_PID": 2, "VELOCITY": 3, "VELOCITY_PID": 4}
REVOLUTE, CONTINUOUS, PRISMATIC = 0, 1, 2


def philox_normals(seed, env, sensor, dim, count):
    """Host replica of plugins.cu::philox_normal (Philox4x32-10 keyed by seed, counter (env, sensor*4+dim, count))."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [env & 0xFFFFFFFF, (sensor * 4 + dim) & 0xFFFFFFFF, count & 0xFFFFFFFF, (count >> 32) & 0xFFFFFFFF]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    u1, u2 = (c[0] + 0.5) / 4294967296.0, (c[1] + 0.5) / 4294967296.0
    return float(np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586476925286766559 * u2))


HW_CASES = {
    # name: (limits per joint or None)
    "no_limit_handles": None,
    "saturation_handles": "sat",
    "soft_limit_handles": "soft",
}


@pytest.mark.parametrize("limit_kind", list(HW_CASES))
@pytest.mark.parametrize("control_every", [1, 5])
def test_robot_hw_matches_reference_restatement(limit_kind, control_every, load_model, BatchSim, capi, orc):
    """Every control mode x joint kind x e-stop phase x limit handle, BITWISE against oracle/orc_plugins.cpp (the
    line-by-line restatement of default_robot_hw_sim.cpp:230-326 + joint_limits_interface + control_toolbox::Pid),
    over 200 steps of the hand (C3's stated path: RK4, robot_hw write every step, read every control period)."""
    model = load_model("hand_like.xml")
    nenv = 6
    names = ["WRJ1", "WRJ0", "FFJ3", "FFJ2", "MFJ3", "MFJ2", "RFJ2", "LFJ2", "THJ4", "THJ2"]
    jids = [model.name2id(capi.OBJ_JOINT, n) for n in names]
    assert min(jids) >= 0
    nj = len(jids)
    modes = [MODES[k] for k in ("EFFORT", "POSITION", "POSITION_PID", "VELOCITY", "VELOCITY_PID",
                                "POSITION_PID", "POSITION_PID", "EFFORT", "VELOCITY_PID", "POSITION_PID")]
    kinds = [REVOLUTE, REVOLUTE, REVOLUTE, REVOLUTE, REVOLUTE, CONTINUOUS, PRISMATIC, CONTINUOUS, REVOLUTE, REVOLUTE]
    lower = [float(model.jnt_range[j, 0]) for j in jids]
    upper = [float(model.jnt_range[j, 1]) for j in jids]
    effort = [0.8, 1.0, 0.6, 1.0, 0.7, 0.9, 0.5, 0.4, 0.6, 0.3]
    pid5 = np.array([[0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [4.0, 1.5, 0.05, 0.2, -0.2], [0, 0, 0, 0, 0],
                     [0.4, 2.0, 0.0, 0.3, -0.3], [3.0, 0.0, 0.02, 0, 0], [5.0, 8.0, 0.01, 0.25, -0.1], [0, 0, 0, 0, 0],
                     [0.3, 0.5, 0.001, 0.1, -0.1], [2.0, 4.0, 0.0, 0.05, -0.05]])
    antiwindup = [0, 0, 1, 0, 0, 0, 1, 0, 1, 0]
    limits = None
    if HW_CASES[limit_kind]:
        soft = HW_CASES[limit_kind] == "soft"
        limits = []
        for k in range(nj):
            mid, half = 0.5 * (lower[k] + upper[k]), 0.5 * (upper[k] - lower[k])
            limits.append(dict(has_position_limits=int(k % 4 != 3), has_velocity_limits=1,
                               has_acceleration_limits=int(k % 2 == 0), has_effort_limits=1, has_soft_limits=int(soft),
                               min_position=lower[k], max_position=upper[k], max_velocity=1.5 + 0.1 * k,
                               max_acceleration=40.0, max_effort=0.9 * effort[k],
                               soft_min_position=mid - 0.8 * half, soft_max_position=mid + 0.8 * half,
                               k_position=20.0 + k, k_velocity=0.5 + 0.05 * k))
    sim = BatchSim(model, nenv)
    sim.robot_hw_configure(jids, modes, effort_limit=effort, pid=pid5, lower=lower, upper=upper, kind=kinds,
                           limits=limits, antiwindup=antiwindup)
    oracles = [orc.Oracle(model) for _ in range(nenv)]
    pid6 = np.concatenate([pid5, np.array(antiwindup, dtype=float)[:, None]], axis=1)
    hws = [orc.RobotHW(o, jids, modes, kinds, lower, upper, effort, pid6, limits) for o in oracles]
    rng = np.random.default_rng(5)
    dt = model.opt.timestep
    cmd = np.zeros((nenv, nj))
    for s in range(200):
        e_stop = 60 <= s < 120 or 150 <= s < 155
        st = {k: sim.get(k) for k in ("qpos", "qvel", "qfrc_applied")}
        for e, o in enumerate(oracles):   # both sides see the batch's own state: isolates the plugin arithmetic
            for k, v in st.items():
                o.set(k, v[e])
        if s % control_every == 0:
            gp, gv, ge = sim.robot_hw_read()
            for e, hw in enumerate(hws):
                hw.read()
                op, ov, oe = hw.state()
                np.testing.assert_array_equal(gp[e], op, err_msg=f"step {s} env {e}: position")
                np.testing.assert_array_equal(gv[e], ov)
                np.testing.assert_array_equal(ge[e], oe)
            # what a controller would produce: targets beyond the limits on purpose
            span = np.array(upper) - np.array(lower)
            cmd = np.where(np.isin(modes, (1, 2)), np.array(lower) + span * rng.uniform(-0.3, 1.3, (nenv, nj)),
                           np.where(np.isin(modes, (3, 4)), rng.uniform(-4, 4, (nenv, nj)), rng.uniform(-1.5, 1.5, (nenv, nj))))
        sim.robot_hw_write(cmd, e_stop=e_stop, period=dt)
        after = {k: sim.get(k) for k in ("qpos", "qvel", "qfrc_applied")}
        for e, hw in enumerate(hws):
            hw.write(cmd[e], e_stop=e_stop, period=dt)
            for k, v in after.items():
                np.testing.assert_array_equal(v[e], oracles[e].get(k), err_msg=f"step {s} env {e}: {k} ({limit_kind})")
        sim.step(1)
    f = sim.get("qfrc_applied")
    assert np.abs(f).max() > 0.05 and np.all(np.isfinite(sim.get("qpos")))


def test_robot_hw_rejects_limits_the_handles_cannot_take(load_model, BatchSim, capi):
    # joint_limits_interface handle constructors throw without velocity (and, for effort, effort) limits
    model = load_model("hand_like.xml")
    sim = BatchSim(model, 2)
    jid = model.name2id(capi.OBJ_JOINT, "FFJ2")
    with pytest.raises(capi.B2mjError):
        sim.robot_hw_configure([jid], [0], limits=[dict(has_position_limits=1, min_position=0, max_position=1)])
    with pytest.raises(capi.B2mjError):
        sim.robot_hw_configure([jid], [1], limits=[dict(has_soft_limits=1, has_position_limits=1)])


def test_sensor_readout_matches_reference_restatement(load_model, BatchSim, capi, orc):
    """lastStageCallback (mujoco_sensor_handler_plugin.cpp:175-437) on every sensor of the humanoid (C4) and of the
    pendulum scene: noise-free values and ground truth BITWISE against oracle/orc_plugins.cpp; noisy values (vector,
    scalar and quaternion sensors, partial flags, packed mean / sigma) to float32 resolution with the oracle fed the
    same Philox normals."""
    for name in ("humanoid_like.xml", "pendulum_scene.xml"):
        model = load_model(name)
        nenv = 5
        sim = BatchSim(model, nenv)
        rng = np.random.default_rng(2)
        if model.nu:
            sim.set("ctrl", rng.uniform(-1, 1, (nenv, model.nu)))
        sim.step(40)
        sd = sim.get("sensordata")
        oracles = [orc.Oracle(model) for _ in range(nenv)]
        for e, o in enumerate(oracles):
            o.set("sensordata", sd[e])
        vals, gt = sim.sensor_readout(want_gt=True)
        for e, o in enumerate(oracles):
            ov, og = orc.sensor_readout(o)
            np.testing.assert_array_equal(vals[e], ov)
            np.testing.assert_array_equal(gt[e], og)
        # noise on a spread of sensors with different flags
        ns = model.nsensor
        flag = np.zeros(ns, dtype=np.int32)
        mean, sigma = np.zeros((ns, 3)), np.zeros((ns, 3))
        models = []
        for i in range(ns):
            fl = (0, 0b001, 0b011, 0b111, 0b101, 0b010, 0b100)[i % 7]
            if fl == 0:
                continue
            flag[i] = fl
            mean[i] = rng.uniform(-0.05, 0.05, 3)
            sigma[i] = rng.uniform(0.01, 0.05, 3)
            models.append((i, tuple(mean[i]), tuple(sigma[i]), fl))
        seed = 99
        sim.sensor_configure_noise(models, seed=seed)
        for count in range(3):
            vals, gt = sim.sensor_readout(want_gt=True)
            for e, o in enumerate(oracles):
                normals = []
                for i in range(ns):
                    if not flag[i]:
                        continue
                    quat = model.sensor_type[i] in (16, 25)
                    if model.sensor_dim[i] == 1 and not quat:
                        normals.append(philox_normals(seed, e, i, 0, count))
                    else:
                        normals += [philox_normals(seed, e, i, k, count) for k in range(3) if flag[i] & (1 << k)]
                ov, og = orc.sensor_readout(o, flag, mean, sigma, np.array(normals + [0.0]))
                np.testing.assert_array_equal(gt[e], og)
                np.testing.assert_allclose(vals[e], ov, rtol=3e-7, atol=1e-9)
                assert np.abs(vals[e] - gt[e]).max() > 1e-3   # the noise really is applied


def test_robot_hw_position_pid_tracks_target(load_model, BatchSim, capi):
    # closed loop sanity: read -> PID -> write -> step drives the joint to its target and respects the effort clamp
    model = load_model("panda_like.xml")
    nenv = 8
    sim = BatchSim(model, nenv)
    jid = model.name2id(capi.OBJ_JOINT, "joint1")
    sim.robot_hw_configure([jid], [2], effort_limit=[50.0], pid=[[200.0, 0.0, 20.0, 0.0, 0.0]], lower=[-2.8],
                           upper=[2.8], kind=[0])
    target = np.full((nenv, 1), 0.4)
    for _ in range(1500):
        sim.robot_hw_read()
        sim.robot_hw_write(target, period=model.opt.timestep)
        sim.step(1)
    assert np.all(np.abs(sim.get("qfrc_applied")[:, model.jnt_dofadr[jid]]) <= 50.0 + 1e-12)
    assert np.all(sim.get("qpos")[:, model.jnt_qposadr[jid]] > 0.05)
