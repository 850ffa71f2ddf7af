"""Thin Python view of a b2mj_handle (tests and bench drive the C-ABI through this).

No compute happens in Python: every method is one C-ABI call into libb2mj.so (CUDA).  There is no
CPU fallback — creating a BatchSim without a GPU raises B2mjError(B2MJ_ENODEVICE).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import B2mjError, Model, check, lib

_vp = C.c_void_p
lib.b2mj_create.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(_vp)]
lib.b2mj_destroy.argtypes = [_vp]
lib.b2mj_destroy.restype = None
lib.b2mj_nenv.argtypes = [_vp]
lib.b2mj_set_stream.argtypes = [_vp, _vp]
lib.b2mj_reset.argtypes = [_vp, _vp]
lib.b2mj_reset_keyframe.argtypes = [_vp, C.c_int, _vp]
lib.b2mj_forward.argtypes = [_vp]
lib.b2mj_step.argtypes = [_vp, C.c_int]
lib.b2mj_rollout.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp]
lib.b2mj_step_host.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp]
lib.b2mj_step_begin.argtypes = [_vp]
lib.b2mj_step_end.argtypes = [_vp]
lib.b2mj_sync.argtypes = [_vp]
lib.b2mj_set_keep_intermediates.argtypes = [_vp, C.c_int]
lib.b2mj_get.argtypes = [_vp, C.c_int, _vp, C.c_size_t]
lib.b2mj_set.argtypes = [_vp, C.c_int, _vp, C.c_size_t]
lib.b2mj_set_device.argtypes = [_vp, C.c_int, _vp, C.c_size_t]
lib.b2mj_stage_profile.argtypes = [_vp, C.c_int, C.POINTER(C.c_uint64), C.c_int]
lib.b2mj_env_cycles.argtypes = [_vp, _vp]
lib.b2mj_stage_name.argtypes = [C.c_int]
lib.b2mj_stage_name.restype = C.c_char_p
lib.b2mj_device_ptr.argtypes = [_vp, C.c_int, C.POINTER(_vp), C.POINTER(C.c_size_t)]
lib.b2mj_model_update.argtypes = [_vp, _vp]
lib.b2mj_register_collision_function.argtypes = [_vp, C.c_int, C.c_int, C.c_int]
lib.b2mj_reset_collision_functions.argtypes = [_vp]
lib.b2mj_set_env_models.argtypes = [_vp, C.POINTER(_vp), C.c_int, _vp]
lib.b2mj_launch_info.argtypes = [_vp, C.POINTER(_capi.B2mjLaunchInfo)]
lib.b2mj_robot_hw_configure.argtypes = [_vp, C.POINTER(_capi.B2mjRobotHW)]
lib.b2mj_robot_hw_write.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_double]
lib.b2mj_robot_hw_read.argtypes = [_vp, _vp, _vp, _vp]
lib.b2mj_sensor_configure_noise.argtypes = [_vp, C.POINTER(_capi.B2mjSensorNoise), C.c_int, C.c_uint64]
lib.b2mj_sensor_readout.argtypes = [_vp, _vp, _vp]
lib.b2mj_allgather_publish.argtypes = [_vp, C.c_int, _vp, _vp]
lib.b2mj_allgather_publish_multi.argtypes = [_vp, C.POINTER(C.c_int), C.c_int, _vp, _vp]
lib.b2mj_publish_pack.argtypes = [_vp, C.POINTER(C.c_int), C.c_int, C.POINTER(_vp), C.POINTER(C.c_int)]
lib.b2mj_publish_fused_create.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, _vp]
lib.b2mj_publish_fused_connect.argtypes = [_vp, _vp]
lib.b2mj_step_publish.argtypes = [_vp]
lib.b2mj_publish_fused_wait.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_int)]
lib.b2mj_sensor_readout_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp)]
lib.b2mj_robot_hw_state_ptrs.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]
lib.b2mj_ubench_dfma.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]


class BatchSim:
    def __init__(self, model: Model, nenv: int, device: int = 0):
        self.model = model
        self.nenv = nenv
        self._h = _vp()
        check(lib.b2mj_create(model.ptr, nenv, device, C.byref(self._h)), "b2mj_create")

    def close(self):
        if getattr(self, "_h", None):
            lib.b2mj_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream_ptr: int):
        check(lib.b2mj_set_stream(self._h, _vp(cuda_stream_ptr)), "set_stream")

    def reset(self, mask=None):
        if mask is None:
            check(lib.b2mj_reset(self._h, None), "reset")
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            assert m.size == self.nenv
            check(lib.b2mj_reset(self._h, m.ctypes.data), "reset")

    def reset_keyframe(self, key: int, mask=None):
        """mj_resetDataKeyframe on all (or the masked) envs."""
        if mask is None:
            check(lib.b2mj_reset_keyframe(self._h, int(key), None), "reset_keyframe")
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            assert m.size == self.nenv
            check(lib.b2mj_reset_keyframe(self._h, int(key), m.ctypes.data), "reset_keyframe")

    def forward(self):
        check(lib.b2mj_forward(self._h), "forward")

    def step(self, n: int = 1):
        check(lib.b2mj_step(self._h, n), "step")

    def rollout(self, nsteps: int, ctrl_ptr: int = 0, qpos_ptr: int = 0, qvel_ptr: int = 0, sensor_ptr: int = 0):
        """Fused open-loop rollout; all pointers are DEVICE addresses (0 = not used)."""
        check(lib.b2mj_rollout(self._h, nsteps, _vp(ctrl_ptr or None), _vp(qpos_ptr or None), _vp(qvel_ptr or None),
                               _vp(sensor_ptr or None)), "rollout")

    def step_begin(self):
        check(lib.b2mj_step_begin(self._h), "step_begin")

    def step_end(self) -> int:
        """Returns 0 when the step is complete, 1 (B2MJ_AGAIN) after an RK4 sub-step: run the hooks, call again."""
        return check(lib.b2mj_step_end(self._h), "step_end")

    def step_host(self, nsteps: int, ctrl: np.ndarray = None, qpos: np.ndarray = None, qvel: np.ndarray = None,
                  sensordata: np.ndarray = None):
        """b2mj_step_host: ctrl up, nsteps steps, qpos / qvel / sensordata down, one synchronisation.  Arrays are
        C-contiguous float64 [nenv][n] host buffers (pinned for full-speed DMA); None skips that transfer."""
        def ptr(a, n):
            if a is None:
                return None
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == self.nenv * n, (a.shape, n)
            return _vp(a.ctypes.data)
        m = self.model
        check(lib.b2mj_step_host(self._h, nsteps, ptr(ctrl, m.nu), ptr(qpos, m.nq), ptr(qvel, m.nv),
                                 ptr(sensordata, m.nsensordata)), "step_host")

    def sync(self):
        check(lib.b2mj_sync(self._h), "sync")

    def keep_intermediates(self, on: bool = True):
        check(lib.b2mj_set_keep_intermediates(self._h, int(on)), "keep_intermediates")

    def get(self, name: str) -> np.ndarray:
        f = _capi.field_id(name)
        n, is_int = self.model.field_size(f)
        out = np.zeros((self.nenv, max(n, 0)), dtype=np.int32 if is_int else np.float64)
        check(lib.b2mj_get(self._h, f, out.ctypes.data, out.nbytes), f"get {name}")
        return out

    def set(self, name: str, value):
        f = _capi.field_id(name)
        n, is_int = self.model.field_size(f)
        arr = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.int32 if is_int else np.float64),
                                                   (self.nenv, n)))
        check(lib.b2mj_set(self._h, f, arr.ctypes.data, arr.nbytes), f"set {name}")

    def get_into(self, name: str, out: np.ndarray):
        """b2mj_get into a caller-provided (e.g. pinned) host array."""
        check(lib.b2mj_get(self._h, _capi.field_id(name), out.ctypes.data, out.nbytes), f"get {name}")

    def set_from(self, name: str, arr: np.ndarray):
        """b2mj_set from a caller-provided (e.g. pinned) contiguous host array."""
        check(lib.b2mj_set(self._h, _capi.field_id(name), arr.ctypes.data, arr.nbytes), f"set {name}")

    def set_device(self, name: str, dev_ptr: int, pitch_elems: int = 0):
        check(lib.b2mj_set_device(self._h, _capi.field_id(name), _vp(dev_ptr), pitch_elems), f"set_device {name}")

    def device_ptr(self, name: str):
        p, pitch = _vp(), C.c_size_t()
        check(lib.b2mj_device_ptr(self._h, _capi.field_id(name), C.byref(p), C.byref(pitch)), f"device_ptr {name}")
        return p.value, pitch.value

    def model_update(self, model: Model = None):
        check(lib.b2mj_model_update(self._h, (model or self.model).ptr), "model_update")

    def register_collision_function(self, type1: int, type2: int, collfn: int):
        check(lib.b2mj_register_collision_function(self._h, type1, type2, collfn), "register_collision_function")

    def reset_collision_functions(self):
        check(lib.b2mj_reset_collision_functions(self._h), "reset_collision_functions")

    def set_env_models(self, models, env_model):
        """Per-env model variants: models = list of Model (edited copies), env_model = [nenv] variant index."""
        if not models:
            check(lib.b2mj_set_env_models(self._h, None, 0, None), "set_env_models")
            return
        ptrs = (_vp * len(models))(*[m.ptr for m in models])
        idx = np.ascontiguousarray(env_model, dtype=np.int32)
        assert idx.size == self.nenv
        check(lib.b2mj_set_env_models(self._h, ptrs, len(models), idx.ctypes.data), "set_env_models")

    def launch_info(self) -> dict:
        li = _capi.B2mjLaunchInfo()
        check(lib.b2mj_launch_info(self._h, C.byref(li)), "launch_info")
        return {k: getattr(li, k) for k, _ in li._fields_}

    def stage_profile(self, enable: bool = True) -> dict:
        """Read (then clear / stop) the per-stage cycle counters: {stage: cycles summed over envs}."""
        buf = (C.c_uint64 * 64)()
        n = check(lib.b2mj_stage_profile(self._h, int(enable), buf, 64), "stage_profile")
        return {lib.b2mj_stage_name(i).decode(): int(buf[i]) for i in range(n)}

    def env_cycles(self) -> np.ndarray:
        out = np.zeros(self.nenv, dtype=np.int32)
        check(lib.b2mj_env_cycles(self._h, out.ctypes.data), "env_cycles")
        return out.astype(np.int64) * 1024

    # ---- plugin data paths ----
    def robot_hw_configure(self, joint_ids, modes, effort_limit=None, pid=None, lower=None, upper=None, kind=None,
                           limits=None, antiwindup=None):
        """limits: list of dicts with b2mjJointLimits field names (missing fields 0), one per joint, or None."""
        nj = len(joint_ids)
        self._hw_keep = []

        def arr(x, dt):
            if x is None:
                return None
            a = np.ascontiguousarray(x, dtype=dt)
            self._hw_keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int if dt == np.int32 else C.c_double))
        lim = None
        if limits is not None:
            assert len(limits) == nj
            lim = (_capi.B2mjJointLimits * nj)()
            for k, d in enumerate(limits):
                for name, val in d.items():
                    setattr(lim[k], name, val)
            self._hw_keep.append(lim)
        cfg = _capi.B2mjRobotHW(nj, arr(joint_ids, np.int32), arr(modes, np.int32), arr(effort_limit, np.float64),
                                arr(pid, np.float64), arr(lower, np.float64), arr(upper, np.float64), arr(kind, np.int32),
                                lim, arr(antiwindup, np.int32))
        check(lib.b2mj_robot_hw_configure(self._h, C.byref(cfg)), "robot_hw_configure")
        self._hw_nj = nj

    def robot_hw_write(self, cmd, e_stop=False, period=0.001):
        c = np.ascontiguousarray(cmd, dtype=np.float64)
        assert c.shape == (self.nenv, self._hw_nj)
        check(lib.b2mj_robot_hw_write(self._h, c.ctypes.data, 0, int(e_stop), float(period)), "robot_hw_write")

    def robot_hw_read(self):
        outs = [np.zeros((self.nenv, self._hw_nj)) for _ in range(3)]
        check(lib.b2mj_robot_hw_read(self._h, *[o.ctypes.data for o in outs]), "robot_hw_read")
        return outs

    def robot_hw_write_device(self, cmd_dev_ptr: int, e_stop=False, period=0.001):
        """cmd already on the device ([nenv][njoint] float64): no copy, asynchronous."""
        check(lib.b2mj_robot_hw_write(self._h, _vp(cmd_dev_ptr), 1, int(e_stop), float(period)), "robot_hw_write")

    def robot_hw_refresh(self):
        """readSim on the device only (no copies, no synchronisation)."""
        check(lib.b2mj_robot_hw_read(self._h, None, None, None), "robot_hw_read")

    def sensor_readout_device(self, want_gt=False):
        v, g = _vp(), _vp()
        check(lib.b2mj_sensor_readout_device(self._h, C.byref(v), C.byref(g) if want_gt else None), "sensor_readout_device")
        return v.value, g.value

    def publish_fields(self, names):
        ids = (C.c_int * len(names))(*[_capi.field_id(n) for n in names])
        return ids, len(names)

    def allgather_publish_multi(self, names, comm_ptr, dst_dev_ptr: int):
        ids, n = self.publish_fields(names)
        check(lib.b2mj_allgather_publish_multi(self._h, ids, n, comm_ptr, _vp(dst_dev_ptr)), "allgather_publish_multi")

    def publish_fused_create(self, world: int, rank: int, names) -> bytes:
        """Allocate this rank's gathered slabs for the fused step + publish; returns the 64-byte CUDA IPC handle."""
        ids = (C.c_int * len(names))(*[_capi.field_id(n) for n in names])
        out = (C.c_ubyte * 64)()
        check(lib.b2mj_publish_fused_create(self._h, world, rank, ids, len(names), C.cast(out, _vp)), "publish_fused_create")
        return bytes(out)

    def publish_fused_connect(self, handles):
        """handles: the IPC handles of all ranks in rank order (each 64 bytes)."""
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(lib.b2mj_publish_fused_connect(self._h, C.cast(buf, _vp)), "publish_fused_connect")

    def step_publish(self):
        check(lib.b2mj_step_publish(self._h), "step_publish")

    def publish_fused_wait(self):
        """Stream-ordered wait for every rank's row of the step just published: (device pointer, count per env)."""
        ptr, cnt = _vp(), C.c_int()
        check(lib.b2mj_publish_fused_wait(self._h, C.byref(ptr), C.byref(cnt)), "publish_fused_wait")
        return ptr.value, cnt.value

    def publish_pack(self, names):
        ids, n = self.publish_fields(names)
        slab, row = _vp(), C.c_int()
        check(lib.b2mj_publish_pack(self._h, ids, n, C.byref(slab), C.byref(row)), "publish_pack")
        return slab.value, row.value

    def sensor_configure_noise(self, models, seed=0):
        arr = (_capi.B2mjSensorNoise * max(1, len(models)))()
        for i, (sid, mean, sigma, flag) in enumerate(models):
            arr[i].sensor_id = sid
            for k in range(3):
                arr[i].mean[k] = mean[k]
                arr[i].sigma[k] = sigma[k]
            arr[i].set_flag = flag
        check(lib.b2mj_sensor_configure_noise(self._h, arr, len(models), seed), "sensor_configure_noise")

    def sensor_readout(self, want_gt=True):
        n = self.model.nsensordata
        v = np.zeros((self.nenv, n))
        g = np.zeros((self.nenv, n)) if want_gt else None
        check(lib.b2mj_sensor_readout(self._h, v.ctypes.data, g.ctypes.data if want_gt else None), "sensor_readout")
        return v, g
