"""CPU checks of oracle/orc_plugins.cpp, the restatement the GPU plugin kernels are compared against bitwise.
The reference's own sources for this path are present (default_robot_hw_sim.cpp, mujoco_sensor_handler_plugin.cpp);
three ROS packages it calls are not, so their restated helpers are pinned here against the known answers of those
packages' own unit tests (ros/angles test/utest.cpp, control_toolbox test/pid_tests.cpp) as far as they are public
knowledge, plus properties that follow from their documentation."""
import ctypes as C
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def olib(orc):
    lib = orc.lib
    lib.orc_angles_shortest_with_limits.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_double)]
    lib.orc_angles_normalize.restype = C.c_double
    lib.orc_angles_normalize.argtypes = [C.c_double]
    lib.orc_pid_run.restype = C.c_double
    lib.orc_pid_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    return lib


def sadwl(olib, a, b, lo, hi):
    out = C.c_double()
    ok = olib.orc_angles_shortest_with_limits(a, b, lo, hi, C.byref(out))
    return bool(ok), out.value


def test_angles_shortest_distance_with_limits_known_answers(olib):
    # ros/angles test/utest.cpp, TEST(Angles, shortestDistanceWithLimits)
    PI = math.pi
    cases = [
        ((-0.5, 0.5, -0.25, 0.25), False, None),
        ((-0.5, 0.5, 0.25, 0.25), False, None),
        ((-0.5, 0.5, 0.25, -0.25), True, -2 * PI + 1.0),
        ((0.5, 0.5, 0.25, -0.25), True, 0.0),
        ((0.5, 0.0, 0.25, -0.25), False, -0.5),
        ((-0.5, 0.0, 0.25, -0.25), False, 0.5),
        ((-0.2, 0.2, 0.25, -0.25), False, -2 * PI + 0.4),
        ((0.2, -0.2, 0.25, -0.25), False, 2 * PI - 0.4),
        ((0.2, 0.0, 0.25, -0.25), False, 2 * PI - 0.2),
        ((-0.2, 0.0, 0.25, -0.25), False, -2 * PI + 0.2),
        ((-0.25, -0.5, 0.25, -0.25), True, -0.25),
        ((-0.25, 0.5, 0.25, -0.25), True, -2 * PI + 0.75),
        ((-0.2500001, 0.5, 0.25, -0.25), True, -2 * PI + 0.5 + 0.2500001),
        ((-0.6, 0.5, -0.25, 0.25), False, None),
        ((-0.5, 0.6, -0.25, 0.25), False, None),
        ((-0.6, 0.75, -0.25, 0.3), False, None),
        ((-0.6, PI * 3.0 / 4.0, -0.25, 0.3), False, None),
        ((-PI, PI, -PI, PI), True, 0.0),
    ]
    for args, ok, val in cases:
        got_ok, got = sadwl(olib, *args)
        assert got_ok == ok, (args, got_ok, got)
        if val is not None:
            assert abs(got - val) < 1e-6, (args, got, val)


def test_angles_normalize(olib):
    # utest.cpp TEST(Angles, normalize): results lie in (-pi, pi]
    PI = math.pi
    for a, want in ((0.0, 0.0), (PI / 2, PI / 2), (PI, PI), (-PI / 2, -PI / 2), (-PI, PI), (3 * PI / 2, -PI / 2),
                    (2 * PI, 0.0), (-3 * PI / 2, PI / 2), (7 * PI / 2, -PI / 2)):
        assert abs(olib.orc_angles_normalize(a) - want) < 1e-12, a


def test_pid_integral_clamp_and_antiwindup(olib):
    """control_toolbox pid_tests.cpp: with i-clamps the integral TERM never exceeds i_max / i_min; with antiwindup
    the integral ERROR itself is clamped so the term leaves the bound immediately when the error changes sign;
    zero dt or non-finite error returns 0 without touching the state; derivative is (e - e_last) / dt."""
    def run(gains, errors, dt=1.0):
        g = np.array(gains, dtype=np.float64)
        e = np.array(errors, dtype=np.float64)
        out = np.zeros(len(errors))
        olib.orc_pid_run(g.ctypes.data, e.ctypes.data, len(errors), dt, out.ctypes.data)
        return out
    # integrationClampTest: i_gain 1, i_max 1, i_min -1, error -10 twice -> command -1... (sign convention: cmd = +i*ie)
    out = run([0, 1.0, 0, 1.0, -1.0, 0], [-10.0, -10.0])
    np.testing.assert_array_equal(out, [-1.0, -1.0])
    # integrationClampZeroGainTest: i_gain 0 -> integral contributes nothing
    out = run([0, 0.0, 0, 1.0, -1.0, 0], [-1.0, -1.0])
    np.testing.assert_array_equal(out, [0.0, 0.0])
    # integrationAntiwindupTest: i_gain 2, bounds +-1, antiwindup: i_error clamped to +-0.5
    out = run([0, 2.0, 0, 1.0, -1.0, 1], [1.0, 1.0, -1.0])
    np.testing.assert_array_equal(out, [1.0, 1.0, -1.0])
    # without antiwindup the raw integral winds up: after +1, +1 the error sum is 2; -1 brings it to 1 -> term 2 -> clamp 1
    out = run([0, 2.0, 0, 1.0, -1.0, 0], [1.0, 1.0, -1.0])
    np.testing.assert_array_equal(out, [1.0, 1.0, 1.0])
    # p + d: derivative of the error over dt
    out = run([2.0, 0, 0.5, 0, 0, 0], [1.0, 3.0, 3.0], dt=0.5)
    np.testing.assert_allclose(out, [2.0 + 0.5 * 2.0, 6.0 + 0.5 * 4.0, 6.0])
    # dt == 0, NaN, inf -> 0
    assert run([1, 1, 1, 1, -1, 0], [1.0], dt=0.0)[0] == 0.0
    np.testing.assert_array_equal(run([1, 0, 0, 0, 0, 0], [np.nan, np.inf, 2.0]), [0.0, 0.0, 2.0])


def test_robot_hw_first_read_unwraps_from_one(orc, load_model, capi):
    """default_robot_hw_sim.cpp:132,238-242: joint_position_ starts at 1.0 and revolute joints accumulate the shortest
    angular distance, so a joint at 1 + 2*pi*k reads back as 1.0 while a prismatic one reads its raw value."""
    model = load_model("hand_like.xml")
    o = orc.Oracle(model)
    q = model.qpos0.copy()
    q[0] = 1.0 + 2 * math.pi
    q[1] = 7.5
    o.set("qpos", q)
    hw = orc.RobotHW(o, [0, 1], [0, 0], [0, 2])
    hw.read()
    pos, vel, eff = hw.state()
    assert abs(pos[0] - 1.0) < 1e-12 and pos[1] == 7.5
    assert eff[0] == 0.0   # overwritten by qfrc_applied (initial 1.0 is only the pre-read value)


def test_robot_hw_literal_transmission_indexing_agrees_on_ordered_models(load_model, capi, orc):
    """The reference indexes its writes with the transmission index (default_robot_hw_sim.cpp:273-321).  When the
    transmissions are the model's first joints in order (hinge / slide only) that equals joint-id indexing -- the
    decision recorded in DESIGN.md (ids everywhere) changes nothing for such models."""
    model = load_model("hand_like.xml")
    nj = 8
    jids = list(range(nj))
    assert all(model.jnt_type[j] in (2, 3) and model.jnt_dofadr[j] == j and model.jnt_qposadr[j] == j for j in jids)
    modes = [0, 1, 2, 3, 4, 0, 2, 4]
    kinds = [0] * nj
    pid6 = np.tile([2.0, 0.5, 0.01, 0.1, -0.1, 0], (nj, 1))
    a, b = orc.Oracle(model), orc.Oracle(model)
    lo, hi = model.jnt_range[:nj, 0].copy(), model.jnt_range[:nj, 1].copy()
    ha = orc.RobotHW(a, jids, modes, kinds, lo, hi, [1.0] * nj, pid6, None, literal_indexing=False)
    hb = orc.RobotHW(b, jids, modes, kinds, lo, hi, [1.0] * nj, pid6, None, literal_indexing=True)
    rng = np.random.default_rng(1)
    for s in range(50):
        cmd = rng.uniform(-0.5, 0.5, nj)
        for o, h in ((a, ha), (b, hb)):
            h.read()
            h.write(cmd, e_stop=False, period=0.002)
            o.step(1)
    np.testing.assert_array_equal(a.get("qpos"), b.get("qpos"))
    np.testing.assert_array_equal(a.get("qfrc_applied"), b.get("qfrc_applied"))


