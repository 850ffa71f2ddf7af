"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by the product
package (mujoco_ros_pkgs_b200 fails loudly without its CUDA library and has no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} missing: run `make oracle`")
lib = C.CDLL(LIB_PATH)

_vp = C.c_void_p
CALLBACK = C.CFUNCTYPE(None, _vp, _vp, _vp)
lib.orc_make_data.restype = _vp
lib.orc_make_data.argtypes = [_vp]
lib.orc_free_data.argtypes = [_vp]
lib.orc_reset_data.argtypes = [_vp, _vp]
lib.orc_forward.argtypes = [_vp, _vp]
lib.orc_reset_keyframe.argtypes = [_vp, _vp, C.c_int]
lib.orc_step.argtypes = [_vp, _vp]
lib.orc_step1.argtypes = [_vp, _vp]
lib.orc_step2.argtypes = [_vp, _vp]
lib.orc_set_callbacks.argtypes = [_vp, _vp, _vp, _vp]
lib.orc_callback_counts.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.orc_get.argtypes = [_vp, _vp, C.c_int, _vp, C.c_int]
lib.orc_set.argtypes = [_vp, _vp, C.c_int, _vp, C.c_int]
lib.orc_rollout.restype = C.c_double
lib.orc_rollout.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp]


class Oracle:
    """One env of the CPU oracle for a compiled model (mujoco_ros_pkgs_b200._capi.Model)."""

    def __init__(self, model):
        from mujoco_ros_pkgs_b200 import _capi

        self._capi = _capi
        self.model = model
        self._d = _vp(lib.orc_make_data(model.ptr))
        self._cbs = []

    def __del__(self):
        try:
            if self._d:
                lib.orc_free_data(self._d)
                self._d = None
        except Exception:
            pass

    def reset(self):
        lib.orc_reset_data(self.model.ptr, self._d)

    def reset_keyframe(self, key: int):
        lib.orc_reset_keyframe(self.model.ptr, self._d, int(key))

    def forward(self):
        lib.orc_forward(self.model.ptr, self._d)

    def step(self, n: int = 1):
        for _ in range(n):
            lib.orc_step(self.model.ptr, self._d)

    def step1(self):
        lib.orc_step1(self.model.ptr, self._d)

    def step2(self):
        lib.orc_step2(self.model.ptr, self._d)

    def set_callbacks(self, control=None, passive=None):
        """control/passive: python callables f(oracle) fired where mjcb_control / mjcb_passive fire."""
        def wrap(f):
            if f is None:
                return None
            cb = CALLBACK(lambda m, d, u: f(self))
            self._cbs.append(cb)
            return cb
        c, p = wrap(control), wrap(passive)
        lib.orc_set_callbacks(self._d, C.cast(c, _vp) if c else None, C.cast(p, _vp) if p else None, None)

    def callback_counts(self):
        a, b = C.c_int(), C.c_int()
        lib.orc_callback_counts(self._d, C.byref(a), C.byref(b))
        return a.value, b.value

    def get(self, name: str) -> np.ndarray:
        f = self._capi.field_id(name)
        n, is_int = self.model.field_size(f)
        out = np.zeros(max(n, 0), dtype=np.int32 if is_int else np.float64)
        if n > 0:
            got = lib.orc_get(self.model.ptr, self._d, f, out.ctypes.data, n)
            assert got == n, (name, got, n)
        return out

    def set(self, name: str, value):
        f = self._capi.field_id(name)
        n, is_int = self.model.field_size(f)
        arr = np.ascontiguousarray(value, dtype=np.int32 if is_int else np.float64).ravel()
        assert arr.size == n, (name, arr.size, n)
        if n > 0:
            rc = lib.orc_set(self.model.ptr, self._d, f, arr.ctypes.data, n)
            assert rc == n

    @property
    def time(self) -> float:
        return float(self.get("time")[0])


def rollout(model, qpos, qvel, nsteps, ctrl=None, nthreads=1, want_sensors=False):
    """CPU baseline driver (reference wrapper-loop semantics). Returns (seconds, qpos, qvel, sensors)."""
    qpos = np.ascontiguousarray(qpos, dtype=np.float64).copy()
    qvel = np.ascontiguousarray(qvel, dtype=np.float64).copy()
    nenv = qpos.shape[0]
    cptr = None
    if ctrl is not None:
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        assert ctrl.shape == (nsteps, nenv, model.nu)
        cptr = ctrl.ctypes.data
    sens = np.zeros((nenv, max(model.nsensordata, 1)), dtype=np.float32) if want_sensors else None
    secs = lib.orc_rollout(model.ptr, nenv, nsteps, qpos.ctypes.data, qvel.ctypes.data, cptr, nthreads,
                           sens.ctypes.data if sens is not None else None)
    return secs, qpos, qvel, sens


# ---- plugin data paths restated from the reference sources (oracle/orc_plugins.cpp) ----
lib.orc_hw_create.restype = _vp
lib.orc_hw_create.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]
lib.orc_hw_free.argtypes = [_vp]
lib.orc_hw_read.argtypes = [_vp, _vp, _vp]
lib.orc_hw_write.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_double]
lib.orc_hw_state.argtypes = [_vp, _vp, _vp, _vp]
lib.orc_sensor_readout.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]


class RobotHW:
    """DefaultRobotHWSim (default_robot_hw_sim.cpp) for ONE env of an Oracle.  pid rows: p, i, d, i_max, i_min,
    antiwindup; limits: list of dicts with b2mjJointLimits field names, or None."""

    def __init__(self, oracle: Oracle, joint_ids, modes, kinds, lower=None, upper=None, effort_limit=None, pid=None,
                 limits=None, literal_indexing=False):
        from mujoco_ros_pkgs_b200 import _capi

        self.o = oracle
        nj = len(joint_ids)
        self.nj = nj

        def arr(x, dt):
            return None if x is None else np.ascontiguousarray(x, dtype=dt)

        a = [arr(joint_ids, np.int32), arr(modes, np.int32), arr(kinds, np.int32), arr(lower, np.float64),
             arr(upper, np.float64), arr(effort_limit, np.float64), arr(pid, np.float64)]
        if a[6] is not None:
            assert a[6].shape == (nj, 6)
        lim = None
        if limits is not None:
            lim = (_capi.B2mjJointLimits * nj)()
            for k, d in enumerate(limits):
                for name, val in d.items():
                    setattr(lim[k], name, val)
        ptr = lambda x: None if x is None else x.ctypes.data
        self._h = _vp(lib.orc_hw_create(oracle.model.ptr, nj, *[ptr(x) for x in a],
                                        C.cast(lim, _vp) if lim is not None else None, int(literal_indexing)))

    def __del__(self):
        try:
            if self._h:
                lib.orc_hw_free(self._h)
                self._h = None
        except Exception:
            pass

    def read(self):
        lib.orc_hw_read(self._h, self.o.model.ptr, self.o._d)

    def write(self, cmd, e_stop=False, period=0.001):
        c = np.ascontiguousarray(cmd, dtype=np.float64)
        assert c.size == self.nj
        lib.orc_hw_write(self._h, self.o.model.ptr, self.o._d, c.ctypes.data, int(e_stop), float(period))

    def state(self):
        outs = [np.zeros(self.nj) for _ in range(3)]
        lib.orc_hw_state(self._h, *[x.ctypes.data for x in outs])
        return outs


def sensor_readout(oracle: Oracle, flag=None, mean=None, sigma=None, normals=None):
    """MujocoRosSensorsPlugin::lastStageCallback arithmetic for one env: returns (values, gt) [nsensordata]."""
    m = oracle.model
    ns = max(m.nsensor, 1)
    flag = np.ascontiguousarray(flag if flag is not None else np.zeros(ns), dtype=np.int32)
    mean = np.ascontiguousarray(mean if mean is not None else np.zeros((ns, 3)), dtype=np.float64)
    sigma = np.ascontiguousarray(sigma if sigma is not None else np.zeros((ns, 3)), dtype=np.float64)
    normals = np.ascontiguousarray(normals if normals is not None else np.zeros(3 * ns), dtype=np.float64)
    v = np.zeros(max(m.nsensordata, 1))
    g = np.zeros(max(m.nsensordata, 1))
    lib.orc_sensor_readout(m.ptr, oracle._d, flag.ctypes.data, mean.ctypes.data, sigma.ctypes.data,
                           normals.ctypes.data, v.ctypes.data, g.ctypes.data)
    return v[:m.nsensordata], g[:m.nsensordata]


class _RolloutArgs(C.Structure):
    _fields_ = [("nenv", C.c_int), ("nsteps", C.c_int), ("nthreads", C.c_int),
                ("qpos", _vp), ("qvel", _vp), ("act", _vp), ("warm", _vp), ("time", _vp),
                ("ctrl", _vp), ("sensor_out", _vp),
                ("hw_njoint", C.c_int), ("hw_joint_id", _vp), ("hw_mode", _vp), ("hw_kind", _vp),
                ("hw_lower", _vp), ("hw_upper", _vp), ("hw_effort", _vp), ("hw_pid6", _vp), ("hw_limits", _vp),
                ("hw_cmd", _vp), ("hw_control_every", C.c_int)]


lib.orc_rollout_ex.restype = C.c_double
lib.orc_rollout_ex.argtypes = [_vp, C.POINTER(_RolloutArgs)]


def rollout_ex(model, state, nsteps, ctrl=None, nthreads=1, want_sensors=False, hw=None, hw_cmd=None,
               hw_control_every=1):
    """CPU baseline driver from a full state snapshot.  state: dict with qpos, qvel and optionally act,
    qacc_warmstart, time ([nenv][...]); updated copies are returned.  hw: dict(joint_ids, modes, kinds, lower, upper,
    effort, pid6) enables the actuator-write path with hw_cmd [nsteps][nenv][nj].  Returns (seconds, state, sensors)."""
    c64 = lambda x: np.ascontiguousarray(x, dtype=np.float64).copy()  # noqa: E731
    st = {k: c64(v) for k, v in state.items() if v is not None and np.size(v)}
    nenv = st["qpos"].shape[0]
    a = _RolloutArgs()
    a.nenv, a.nsteps, a.nthreads = nenv, nsteps, nthreads
    a.qpos, a.qvel = st["qpos"].ctypes.data, st["qvel"].ctypes.data
    a.act = st["act"].ctypes.data if "act" in st else None
    a.warm = st["qacc_warmstart"].ctypes.data if "qacc_warmstart" in st else None
    a.time = st["time"].ctypes.data if "time" in st else None
    keep = []
    if ctrl is not None and model.nu:
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        assert ctrl.shape == (nsteps, nenv, model.nu), ctrl.shape
        a.ctrl = ctrl.ctypes.data
    sens = np.zeros((nenv, max(model.nsensordata, 1)), dtype=np.float32) if want_sensors else None
    a.sensor_out = sens.ctypes.data if sens is not None else None
    if hw is not None:
        nj = len(hw["joint_ids"])
        a.hw_njoint = nj
        for name, key, dt in (("hw_joint_id", "joint_ids", np.int32), ("hw_mode", "modes", np.int32),
                              ("hw_kind", "kinds", np.int32), ("hw_lower", "lower", np.float64),
                              ("hw_upper", "upper", np.float64), ("hw_effort", "effort", np.float64),
                              ("hw_pid6", "pid6", np.float64)):
            arr = np.ascontiguousarray(hw[key], dtype=dt)
            keep.append(arr)
            setattr(a, name, arr.ctypes.data)
        hw_cmd = np.ascontiguousarray(hw_cmd, dtype=np.float64)
        assert hw_cmd.shape == (nsteps, nenv, nj)
        a.hw_cmd = hw_cmd.ctypes.data
        a.hw_control_every = hw_control_every
    secs = lib.orc_rollout_ex(model.ptr, C.byref(a))
    return secs, st, sens


# ---- implicit integrators: dense velocity derivatives after a forward pass (oracle/orc_implicit.cpp) ----
lib.orc_rne_vel_derivative.argtypes = [_vp, _vp, _vp]
lib.orc_smooth_vel_derivative.argtypes = [_vp, _vp, C.c_int, _vp]


def rne_vel_derivative(oracle: Oracle) -> np.ndarray:
    """d qfrc_bias / d qvel, dense [nv][nv], at the oracle's current state (call forward() first)."""
    nv = oracle.model.nv
    out = np.zeros((nv, nv))
    lib.orc_rne_vel_derivative(oracle.model.ptr, oracle._d, out.ctypes.data)
    return out


def smooth_vel_derivative(oracle: Oracle, flg_bias: bool) -> np.ndarray:
    """mjd_smooth_vel: qDeriv on MuJoCo's ancestor/descendant sparsity pattern, dense [nv][nv]."""
    nv = oracle.model.nv
    out = np.zeros((nv, nv))
    lib.orc_smooth_vel_derivative(oracle.model.ptr, oracle._d, int(flg_bias), out.ctypes.data)
    return out
