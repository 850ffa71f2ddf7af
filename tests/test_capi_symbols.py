"""The C-ABI library loads on a CPU-only box and exports every symbol include/b2mj.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "b2mj.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2mj_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_export_list(capi):
    assert header_functions() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(capi):
    for name in header_functions():
        assert hasattr(capi.lib, name), name


def test_version_and_error_string(capi):
    capi.lib.b2mj_version.restype = ctypes.c_int
    assert capi.lib.b2mj_version() == 100
    out = ctypes.c_void_p()
    rc = capi.lib.b2mj_model_from_xml_string(b"<mujoco><worldbody><body><nonsense/></body></worldbody></mujoco>",
                                             ctypes.byref(out))
    assert rc < 0 and capi.last_error() != ""


def test_no_cpu_fallback_without_a_device(capi, load_model, gpu_available):
    """The product path must fail loudly (B2MJ_ENODEVICE) when there is no GPU: no CPU fallback."""
    if gpu_available:
        return
    from mujoco_ros_pkgs_b200.batch import BatchSim

    try:
        BatchSim(load_model("pendulum_scene.xml"), 4)
    except capi.B2mjError as ex:
        assert "-7" in str(ex) or "no CUDA device" in str(ex)
    else:
        raise AssertionError("BatchSim must not be constructible without a CUDA device")


def test_product_does_not_link_the_oracle():
    """oracle/ is test infrastructure: the product library and package never reference it."""
    pkg = os.path.join(ROOT, "mujoco_ros_pkgs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle/" not in text.replace("test oracle", ""), f
    mk = open(os.path.join(ROOT, "Makefile")).read()
    lib_rule = mk[mk.index("$(LIB):"):]
    assert "oracle" not in lib_rule.split("clean:")[0]


def test_field_reflection(capi, load_model):
    m = load_model("pendulum_scene.xml")
    n, is_int = m.field_size(capi.field_id("qpos"))
    assert (n, is_int) == (13, False)
    n, is_int = m.field_size(capi.field_id("ncon"))
    assert (n, is_int) == (1, True)
    assert capi.lib.b2mj_field_by_name(b"not_a_field") == -1
