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


def test_robot_hw_effort_and_position_modes(load_model, BatchSim, capi):
    # default_robot_hw_sim.cpp:271-282: EFFORT -> qfrc_applied = cmd (0 on e-stop); POSITION -> qpos=cmd, qvel=0
    model = load_model("panda_like.xml")
    nenv = 16
    sim = BatchSim(model, nenv)
    jids = [model.name2id(capi.OBJ_JOINT, f"joint{k}") for k in (1, 2, 3)]
    sim.robot_hw_configure(jids, [capi_mode("EFFORT"), capi_mode("POSITION"), capi_mode("VELOCITY")],
                           effort_limit=[87, 87, 87], pid=np.zeros((3, 5)), lower=[-2.8, -1.7, -2.8],
                           upper=[2.8, 1.7, 2.8], kind=[0, 0, 0])
    rng = np.random.default_rng(0)
    cmd = rng.uniform(-0.5, 0.5, (nenv, 3))
    sim.robot_hw_write(cmd)
    qf, qp, qv = sim.get("qfrc_applied"), sim.get("qpos"), sim.get("qvel")
    d = [model.jnt_dofadr[j] for j in jids]
    a = [model.jnt_qposadr[j] for j in jids]
    np.testing.assert_array_equal(qf[:, d[0]], cmd[:, 0])
    np.testing.assert_array_equal(qp[:, a[1]], cmd[:, 1])
    np.testing.assert_array_equal(qv[:, d[1]], 0)
    np.testing.assert_array_equal(qv[:, d[2]], cmd[:, 2])
    sim.robot_hw_write(cmd, e_stop=True)
    np.testing.assert_array_equal(sim.get("qfrc_applied")[:, d[0]], 0)
    pos, vel, eff = sim.robot_hw_read()
    np.testing.assert_allclose(pos[:, 1], cmd[:, 1], atol=1e-12)


def capi_mode(name):
    return {"EFFORT": 0, "POSITION": 1, "POSITION_PID": 2, "VELOCITY": 3, "VELOCITY_PID": 4}[name]


def test_robot_hw_position_pid_tracks_target(load_model, BatchSim, capi):
    # default_robot_hw_sim.cpp:284-304: error -> PID -> clamp(effort_limit) -> qfrc_applied
    model = load_model("panda_like.xml")
    nenv = 8
    sim = BatchSim(model, nenv)
    jid = model.name2id(capi.OBJ_JOINT, "joint1")
    sim.robot_hw_configure([jid], [2], effort_limit=[50.0], pid=[[200.0, 0.0, 20.0, 0.0, 0.0]], lower=[-2.8],
                           upper=[2.8], kind=[0])
    target = np.full((nenv, 1), 0.4)
    ctrl = np.tile(model.qpos0[:8] * 0, (nenv, 1))
    for _ in range(1500):
        sim.robot_hw_write(target, period=model.opt.timestep)
        sim.step(1)
    assert np.all(np.abs(sim.get("qfrc_applied")[:, model.jnt_dofadr[jid]]) <= 50.0 + 1e-12)
    # position servo on joint1 (ctrl=0) fights the PID; it must at least move toward the target
    assert np.all(sim.get("qpos")[:, model.jnt_qposadr[jid]] > 0.05)


def test_robot_hw_on_hand_effort_and_pid(load_model, BatchSim, capi, orc):
    """C3 path: the hand driven through robot_hw_write (EFFORT + POSITION_PID) under RK4; the effort part
    is checked against the oracle stepping with the same qfrc_applied."""
    model = load_model("hand_like.xml")
    nenv = 8
    names = ["FFJ2", "MFJ2", "RFJ2", "WRJ0"]
    jids = [model.name2id(capi.OBJ_JOINT, n) for n in names]
    sim = BatchSim(model, nenv)
    sim.robot_hw_configure(jids, [0, 0, 0, 0], effort_limit=[2.0] * 4, pid=np.zeros((4, 5)),
                           lower=[0, 0, 0, -0.7], upper=[1.57, 1.57, 1.57, 0.49], kind=[0, 0, 0, 0])
    rng = np.random.default_rng(3)
    cmd = rng.uniform(-0.2, 0.2, (nenv, 4))
    oracles = [orc.Oracle(model) for _ in range(nenv)]
    for _ in range(40):
        sim.robot_hw_write(cmd)
        sim.step(1)
        for e, o in enumerate(oracles):
            qf = np.zeros(model.nv)
            for k, j in enumerate(jids):
                qf[model.jnt_dofadr[j]] = cmd[e, k]
            o.set("qfrc_applied", qf)
            o.step(1)
    gq = sim.get("qpos")
    oq = np.stack([o.get("qpos") for o in oracles])
    assert np.max(np.abs(gq - oq) / (1 + np.abs(oq))) < 1e-5
