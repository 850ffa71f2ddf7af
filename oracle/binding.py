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
