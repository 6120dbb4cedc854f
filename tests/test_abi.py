"""The C-ABI library loads without a GPU, exports every symbol include/rt_api.h declares, its pure
host helpers match the reference fixtures, and device entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from raytracinginonesemester_b200 import _abi as A, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rt_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "librt_b200.so does not export " + n
    assert set(names) == set(A.EXPORTS), set(names) ^ set(A.EXPORTS)
    assert lib.rt_api_version() == A.RT_API_VERSION == 5


def test_struct_layouts_match_the_reference():
    assert C.sizeof(A.rt_material) == 52      # Material, GPUandCPU/include/material.h (52 B)
    assert C.sizeof(A.rt_light) == 28         # Light, scene.h:21-25
    assert C.sizeof(A.rt_camera) == 48


def test_host_helpers_match_reference_vectors(golden):
    v = golden("ref_vectors.npz")
    for row in v["cameras"]:
        cam = api.camera_init(row[0:3], row[3:6], row[6:9], row[9], row[10], int(row[11]), int(row[12]))
        got = np.array(list(cam.center) + list(cam.pixel00_loc) + list(cam.pixel_delta_u) + list(cam.pixel_delta_v), np.float32)
        assert np.array_equal(got, row[13:].astype(np.float32))
    with pytest.raises(api.RtError):
        api.camera_init((0, 0, 0), (0, 1, 0), (0, 0, 1), 50, 24, 0, 1)
    with pytest.raises(api.RtError):
        api.camera_init((0, 0, 0), (0, 1, 0), (0, 0, 1), 50, 24, 1, -3)
    assert np.array_equal(api.jitter_table(16, 42, True), v["jitter16_seed42"])
    assert np.array_equal(api.jitter_table(4, 42, False), v["jitter_hw1_4_seed42"])
    assert np.array_equal(api.jitter_table(700, 12345, True), v["jitter700_seed12345"])


def test_camera_properties_of_the_reference_tests():
    """HW1/tests/test_camera.cpp:10-79 restated against the current API: the 1x1 pixel sits on the
    optical axis at the focal distance; the pixel grid is planar and perpendicular to forward."""
    cam = api.camera_init((0, 0, 0), (0, 0, -1), (0, 1, 0), 50.0, 24.0, 1, 1)
    assert np.allclose(list(cam.pixel00_loc), [0, 0, -0.05], atol=1e-9)
    cam = api.camera_init((1, 2, 3), (4, -1, 0.5), (0, 0, 1), 35.0, 24.0, 64, 48)
    c, p00, du, dv = (np.array(list(x), np.float64) for x in (cam.center, cam.pixel00_loc, cam.pixel_delta_u, cam.pixel_delta_v))
    fwd = np.array([3, -3, -2.5]); fwd /= np.linalg.norm(fwd)
    for (i, j) in ((0, 0), (63, 0), (0, 47), (63, 47), (31, 20)):
        p = p00 + i * du + j * dv
        assert abs(np.dot(p - c, fwd) - 0.035) < 1e-7
    assert abs(np.dot(du, fwd)) < 1e-9 and abs(np.dot(dv, fwd)) < 1e-9 and abs(np.dot(du, dv)) < 1e-12


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="only meaningful on a box without a GPU")
def test_device_entry_points_fail_loudly_without_a_gpu():
    with pytest.raises(api.RtError) as e:
        api.Renderer(0)
    assert e.value.code in (A.RT_ERR_CUDA, A.RT_ERR_ARG)
