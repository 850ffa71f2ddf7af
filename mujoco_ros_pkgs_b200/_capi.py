"""ctypes binding of the C-ABI in include/b2mj.h (libb2mj.so).

Python is test harness and benchmark driver only; the product is the shared library.  There is no
fallback: if the library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2MJ_LIB: developer override to load an experimental build of the same library (tools/ sweeps)
LIB_PATH = os.environ.get("B2MJ_LIB") or os.path.join(_HERE, "libb2mj.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make lib` (or __graft_entry__.build()); "
        "there is no CPU/Python fallback for the batched step"
    )

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


class B2mjOption(C.Structure):
    _fields_ = [
        ("timestep", C.c_double), ("impratio", C.c_double), ("tolerance", C.c_double),
        ("ls_tolerance", C.c_double), ("noslip_tolerance", C.c_double), ("mpr_tolerance", C.c_double),
        ("gravity", C.c_double * 3), ("wind", C.c_double * 3), ("magnetic", C.c_double * 3),
        ("density", C.c_double), ("viscosity", C.c_double), ("o_margin", C.c_double),
        ("o_solref", C.c_double * 2), ("o_solimp", C.c_double * 5),
        ("integrator", C.c_int), ("collision", C.c_int), ("cone", C.c_int), ("jacobian", C.c_int),
        ("solver", C.c_int), ("iterations", C.c_int), ("ls_iterations", C.c_int),
        ("noslip_iterations", C.c_int), ("mpr_iterations", C.c_int), ("disableflags", C.c_int),
        ("enableflags", C.c_int), ("_pad", C.c_int),
    ]


class B2mjStatistic(C.Structure):
    _fields_ = [("meaninertia", C.c_double), ("meanmass", C.c_double), ("meansize", C.c_double),
                ("extent", C.c_double), ("center", C.c_double * 3)]


class B2mjLaunchInfo(C.Structure):
    _fields_ = [("warps_per_cta", C.c_int), ("ctas", C.c_int), ("smem_bytes_per_cta", C.c_int),
                ("arena_doubles_per_env", C.c_int), ("arena_in_smem", C.c_int),
                ("state_record_bytes", C.c_int), ("regs_per_thread", C.c_int), ("launches", C.c_uint64)]


class B2mjJointLimits(C.Structure):
    _fields_ = [("has_position_limits", C.c_int), ("has_velocity_limits", C.c_int),
                ("has_acceleration_limits", C.c_int), ("has_effort_limits", C.c_int),
                ("has_soft_limits", C.c_int), ("angle_wraparound", C.c_int),
                ("min_position", C.c_double), ("max_position", C.c_double), ("max_velocity", C.c_double),
                ("max_acceleration", C.c_double), ("max_effort", C.c_double),
                ("soft_min_position", C.c_double), ("soft_max_position", C.c_double),
                ("k_position", C.c_double), ("k_velocity", C.c_double)]


class B2mjRobotHW(C.Structure):
    _fields_ = [("njoint", C.c_int), ("joint_id", C.POINTER(C.c_int)), ("control_mode", C.POINTER(C.c_int)),
                ("effort_limit", C.POINTER(C.c_double)), ("pid_gains", C.POINTER(C.c_double)),
                ("lower_limit", C.POINTER(C.c_double)), ("upper_limit", C.POINTER(C.c_double)),
                ("joint_kind", C.POINTER(C.c_int)), ("limits", C.POINTER(B2mjJointLimits)),
                ("pid_antiwindup", C.POINTER(C.c_int))]


class B2mjSensorNoise(C.Structure):
    _fields_ = [("sensor_id", C.c_int), ("mean", C.c_double * 3), ("sigma", C.c_double * 3),
                ("set_flag", C.c_int)]


# every symbol include/b2mj.h declares (tests/test_capi_symbols.py checks the header against this list)
EXPORTS = [
    "b2mj_field_size", "b2mj_field_name", "b2mj_field_by_name", "b2mj_model_from_xml_file",
    "b2mj_model_from_xml_string", "b2mj_model_free", "b2mj_model_save_binary", "b2mj_model_load_binary",
    "b2mj_model_from_file", "b2mj_reset_keyframe", "b2mj_model_set_const", "b2mj_name2id", "b2mj_id2name",
    "b2mj_model_narrays", "b2mj_model_array_info", "b2mj_model_nsizes", "b2mj_model_size_info",
    "b2mj_create", "b2mj_destroy", "b2mj_nenv", "b2mj_model", "b2mj_set_stream", "b2mj_reset",
    "b2mj_forward", "b2mj_step", "b2mj_rollout", "b2mj_step_begin", "b2mj_step_end", "b2mj_step_host", "b2mj_sync",
    "b2mj_set_keep_intermediates", "b2mj_get", "b2mj_set", "b2mj_set_device", "b2mj_device_ptr", "b2mj_model_update",
    "b2mj_set_env_models", "b2mj_register_collision_function", "b2mj_reset_collision_functions",
    "b2mj_robot_hw_configure", "b2mj_robot_hw_write", "b2mj_robot_hw_read", "b2mj_robot_hw_state_ptrs", "b2mj_sensor_configure_noise",
    "b2mj_sensor_readout", "b2mj_sensor_readout_device", "b2mj_allgather_publish", "b2mj_allgather_publish_multi",
    "b2mj_publish_pack", "b2mj_publish_fused_create", "b2mj_publish_fused_connect", "b2mj_step_publish",
    "b2mj_publish_fused_wait", "b2mj_ubench_dfma", "b2mj_launch_info", "b2mj_stage_profile", "b2mj_stage_name", "b2mj_env_cycles", "b2mj_last_error", "b2mj_version",
    "b2mj_device_count",
]

_vp = C.c_void_p
lib.b2mj_last_error.restype = C.c_char_p
lib.b2mj_field_name.restype = C.c_char_p
lib.b2mj_field_name.argtypes = [C.c_int]
lib.b2mj_field_by_name.argtypes = [C.c_char_p]
lib.b2mj_field_size.argtypes = [_vp, C.c_int, C.POINTER(C.c_int)]
lib.b2mj_model_from_xml_file.argtypes = [C.c_char_p, C.POINTER(_vp)]
lib.b2mj_model_from_xml_string.argtypes = [C.c_char_p, C.POINTER(_vp)]
lib.b2mj_model_free.argtypes = [_vp]
lib.b2mj_model_free.restype = None
lib.b2mj_model_save_binary.argtypes = [_vp, C.c_char_p]
lib.b2mj_model_load_binary.argtypes = [C.c_char_p, C.POINTER(_vp)]
lib.b2mj_model_from_file.argtypes = [C.c_char_p, C.POINTER(_vp)]
lib.b2mj_model_set_const.argtypes = [_vp]
lib.b2mj_name2id.argtypes = [_vp, C.c_int, C.c_char_p]
lib.b2mj_id2name.argtypes = [_vp, C.c_int, C.c_int]
lib.b2mj_id2name.restype = C.c_char_p
lib.b2mj_model_array_info.argtypes = [_vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_vp)]
lib.b2mj_model_size_info.argtypes = [_vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]


def last_error() -> str:
    return (lib.b2mj_last_error() or b"").decode()


class B2mjError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        raise B2mjError(f"{what} failed ({rc}): {last_error()}")
    return rc


# object types (b2mj.h)
OBJ_BODY, OBJ_XBODY, OBJ_JOINT, OBJ_DOF, OBJ_GEOM, OBJ_SITE = 1, 2, 3, 4, 5, 6
OBJ_EQUALITY, OBJ_TENDON, OBJ_ACTUATOR, OBJ_SENSOR, OBJ_KEY, OBJ_PAIR = 15, 16, 17, 18, 21, 13


class Model:
    """Owning wrapper of a b2mjModel*; arrays are exposed as numpy views by their mjModel names."""

    def __init__(self, ptr: int):
        self._ptr = _vp(ptr)
        self._arrays = {}
        self._sizes = {}
        self._reflect()

    def _reflect(self):
        name, kind, rows, cols, ptr = C.c_char_p(), C.c_int(), C.c_int(), C.c_int(), _vp()
        self._arrays.clear()
        for i in range(lib.b2mj_model_narrays()):
            check(lib.b2mj_model_array_info(self._ptr, i, C.byref(name), C.byref(kind), C.byref(rows),
                                            C.byref(cols), C.byref(ptr)), "array_info")
            n = rows.value * cols.value
            ctype = (C.c_double, C.c_int, C.c_char)[kind.value]
            if n == 0 or not ptr.value:
                arr = np.zeros((rows.value, cols.value) if cols.value > 1 else (rows.value,),
                               dtype=(np.float64, np.int32, np.uint8)[kind.value])
            else:
                buf = (ctype * n).from_address(ptr.value)
                arr = np.frombuffer(buf, dtype=(np.float64, np.int32, np.uint8)[kind.value])
                if cols.value > 1:
                    arr = arr.reshape(rows.value, cols.value)
            self._arrays[name.value.decode()] = arr
        val = C.c_int()
        for i in range(lib.b2mj_model_nsizes()):
            check(lib.b2mj_model_size_info(self._ptr, i, C.byref(name), C.byref(val)), "size_info")
            self._sizes[name.value.decode()] = val.value
        # opt/stat live right after the size ints (+1 pad int)
        nsz = lib.b2mj_model_nsizes() + 1
        off = (nsz * 4 + 7) // 8 * 8
        self.opt = B2mjOption.from_address(self._ptr.value + off)
        self.stat = B2mjStatistic.from_address(self._ptr.value + off + C.sizeof(B2mjOption))

    @classmethod
    def from_xml_file(cls, path: str) -> "Model":
        out = _vp()
        check(lib.b2mj_model_from_xml_file(path.encode(), C.byref(out)), f"load {path}")
        return cls(out.value)

    @classmethod
    def from_xml_string(cls, xml: str) -> "Model":
        out = _vp()
        check(lib.b2mj_model_from_xml_string(xml.encode(), C.byref(out)), "compile xml")
        return cls(out.value)

    @classmethod
    def from_file(cls, path: str) -> "Model":
        """Extension dispatch like the reference's loader (mujoco_env.cpp:771-911): .b2mjb binary, else MJCF XML."""
        out = _vp()
        check(lib.b2mj_model_from_file(path.encode(), C.byref(out)), f"load {path}")
        return cls(out.value)

    @classmethod
    def load_binary(cls, path: str) -> "Model":
        out = _vp()
        check(lib.b2mj_model_load_binary(path.encode(), C.byref(out)), f"load {path}")
        return cls(out.value)

    def save_binary(self, path: str):
        check(lib.b2mj_model_save_binary(self._ptr, path.encode()), f"save {path}")

    def __getattr__(self, k):
        d = self.__dict__
        if "_arrays" in d and k in d["_arrays"]:
            return d["_arrays"][k]
        if "_sizes" in d and k in d["_sizes"]:
            return d["_sizes"][k]
        raise AttributeError(k)

    @property
    def ptr(self):
        return self._ptr

    def set_const(self):
        check(lib.b2mj_model_set_const(self._ptr), "set_const")
        self._reflect()

    def name2id(self, objtype: int, name: str) -> int:
        return lib.b2mj_name2id(self._ptr, objtype, name.encode())

    def id2name(self, objtype: int, i: int) -> str:
        s = lib.b2mj_id2name(self._ptr, objtype, i)
        return s.decode() if s is not None else None

    def field_size(self, field: int):
        is_int = C.c_int()
        n = check(lib.b2mj_field_size(self._ptr, field, C.byref(is_int)), "field_size")
        return n, bool(is_int.value)

    def field_size_by_name(self, name: str) -> int:
        return self.field_size(field_id(name))[0]

    def __del__(self):
        try:
            if self._ptr:
                lib.b2mj_model_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def field_id(name: str) -> int:
    i = lib.b2mj_field_by_name(name.encode())
    if i < 0:
        raise KeyError(name)
    return i


NFIELD = 0
while lib.b2mj_field_name(NFIELD) is not None:
    NFIELD += 1
FIELD_NAMES = [lib.b2mj_field_name(i).decode() for i in range(NFIELD)]
