"""Drop-in check of the C++ host driver: rt_render_cli (built here with g++ against the in-tree librt_b200.so) is
run on the very scene JSON + OBJ files the reference's whole bvh_viz program was run on (fixture
tests/golden/e2e_scene.npz, tools/make_golden_e2e.py), and its PPM is compared with the reference's 8-bit image.
Covers the JSON dialect, OBJ ingest (on the host, and on the device with --device-ingest), transforms (baked on the device by default, on the host with --host-transform), per-object materials, two lights, 2 spp jitter, mirror and
hash-RNG diffuse bounces (max_bounces 3) and the P6 writer."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_p6(path):
    data = open(path, "rb").read()
    assert data[:2] == b"P6"
    parts, pos = [], 2
    while len(parts) < 3:                      # width, height, maxval
        while data[pos:pos + 1].isspace():
            pos += 1
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        parts.append(int(data[pos:end]))
        pos = end
    pos += 1                                   # single whitespace after maxval
    w, h, maxval = parts
    assert maxval == 255
    return np.frombuffer(data[pos:pos + 3 * w * h], np.uint8).reshape(h, w, 3)


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    exe = str(d / "rt_render_cli")
    H = os.path.join(ROOT, "raytracinginonesemester_b200", "host")
    pkg = os.path.join(ROOT, "raytracinginonesemester_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(H, "rt_render_cli.cpp"), os.path.join(H, "mesh_ingest.cpp"),
           os.path.join(H, "scene_json.cpp"), "-I" + os.path.join(ROOT, "include"), "-L" + pkg, "-l:librt_b200.so", "-Wl,-rpath," + pkg, "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


@pytest.mark.parametrize("name,extra", [("mirror", []), ("diffuse", []), ("mirror", ["--host-transform"]), ("mirror", ["--device-ingest"]), ("diffuse", ["--device-ingest"])])
def test_cli_matches_reference_program(cli, golden, tmp_path, name, extra):
    g = golden("e2e_scene.npz")
    open(tmp_path / "ball.obj", "w").write(str(g["ball_obj"]))
    open(tmp_path / "ground.obj", "w").write(str(g["ground_obj"]))
    open(tmp_path / "scene.json", "w").write(str(g["json_" + name]))
    r = subprocess.run([cli, "scene.json", "-o", "out.ppm"] + extra, cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "GPU LBVH Build Time" in r.stdout and "GPU Render Time" in r.stdout      # the reference's own progress lines
    got = read_p6(str(tmp_path / "out.ppm"))
    ref = g["image_" + name]
    assert got.shape == ref.shape
    # reference PNG bytes are trunc(255*min(c,1)) (main.cu:428-430); the P6 writer rounds (ppm_p6.cpp:137-155): <= 1 LSB apart
    d = np.abs(got.astype(int) - ref.astype(int))
    assert d.max() <= 1, "%d px differ by more than 1 LSB (max %d)" % ((d.max(-1) > 1).sum(), d.max())
    assert (got.astype(int) - ref.astype(int)).min() >= 0                             # rounding never falls below truncation


def test_cli_error_behaviour(cli, tmp_path):
    r = subprocess.run([cli, "missing.json"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to load scene" in r.stderr
    open(tmp_path / "empty.obj", "w").write("# nothing\n")
    r = subprocess.run([cli, "empty.obj"], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 1 and "No valid geometry loaded" in r.stderr                # main.cu:192-195
