"""OBJ ingest on the device (rt_dmesh_parse_obj, csrc/rt_obj_device.cu) against the host loader (rt_mesh_load_obj,
host/mesh_ingest.cpp — itself pinned to the reference's LoadOBJ_ToMesh output by tests/test_host_ingest.py and
tests/golden/*_mesh.npz): the same bytes must give the same arrays bit for bit — positions, normals, indices, per-triangle
object ids, the next object id — and the same refusals.  The texts are generated here (the reference's .obj files do not
travel to the GPU box)."""
import os

import numpy as np
import pytest

from raytracinginonesemester_b200 import _abi as A, api, scenes
from raytracinginonesemester_b200.api import DeviceMesh, DeviceScene, RtError

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T


def both(renderer, text, tmp_path, first_id=0):
    """(host result | exception text, device result | exception text)"""
    data = text if isinstance(text, bytes) else text.encode()
    path = tmp_path / "m.obj"
    path.write_bytes(data)
    try:
        host = api.load_obj(str(path), first_id)
    except RtError as e:
        host = str(e)
    try:
        dm, nid = DeviceMesh.parse_obj(renderer, data, first_id)
        dev = dm.download() + (nid,)
        dev_stats = dm.stats()
        dm.close()
    except RtError as e:
        dev, dev_stats = str(e), None
    return host, dev, dev_stats


def assert_same(host, dev, what=""):
    assert not isinstance(host, str), what + ": host loader refused: " + str(host)
    assert not isinstance(dev, str), what + ": device parser refused: " + str(dev)
    hp, hn, hi, ho, hid = host
    dp, dn, di, do, did = dev
    assert hp.shape == dp.shape and np.array_equal(hp.view(np.uint32), dp.view(np.uint32)), what + ": positions"
    assert (hn is None) == (dn is None), what + ": normal stream present on one side only"
    if hn is not None:
        assert hn.shape == dn.shape and np.array_equal(hn.view(np.uint32), dn.view(np.uint32)), what + ": normals"
    assert np.array_equal(hi, di), what + ": indices"
    assert np.array_equal(ho, do), what + ": object ids"
    assert hid == did, what + ": next object id"


def obj_text(pos, idx, nrm=None, fmt="%.6f", style="vn"):
    lines = ["# generated", ""]
    lines += ["v " + " ".join(fmt % c for c in p) for p in pos]
    if nrm is not None:
        lines += ["vn " + " ".join(fmt % c for c in n) for n in nrm]
    for a, b, c in idx + 1:
        if nrm is not None and style == "vn":
            lines.append("f %d//%d %d//%d %d//%d" % (a, a, b, b, c, c))
        else:
            lines.append("f %d %d %d" % (a, b, c))
    return "\n".join(lines) + "\n"


def test_frog_sized_mesh_with_normals(renderer, tmp_path):
    g = np.load(os.path.join(GOLD, "frog_mesh.npz"))
    text = obj_text(g["positions"], g["indices"].astype(np.int64), g["normals"])
    host, dev, st = both(renderer, text, tmp_path)
    assert_same(host, dev, "frog")
    assert host[2].shape[0] == g["indices"].shape[0]
    assert st[1] > g["indices"].shape[0] and st[2] < 64           # lines parsed; almost nothing needs the host's strtof


def test_every_syntax_the_loader_accepts(renderer, tmp_path):
    text = "\r\n".join([
        "# comment", "", "   \t ", "mtllib x.mtl", "v 0 0 0", "v 1 0 0", "v\t1 1 0", "v 0 1 0", "v .5 5. -0.25", "v +1e-3 1E+2 -2.5e0",
        "v 0.1234567890123456789 16777217 33554434.0", "v 1e-45 3.4028235e38 1.17549435e-38", "v 0.3 0.7 1e23",
        "vt 0 0", "vt 1 0.5", "vn 0 0 1", "vn 0 1 0", "vn 1 0 0",
        "f 1 2 3", "f 1/1 2/2 3/1", "f 1//1 2//2 3//3", "f 1/1/1 2/2/2 3/1/3 4/2/1", "o second", "f -1 -2 -3", "f -1//-1 -2//-2 -3//-3",
        "g third", "f 5/1 6/1 7/", "f 1/2/3 2/1/1 3/2/", "f 1 2 3 4 5 6", "usemtl m", "s off", "f 6 7 8 9", "f 9//9 1//0 2//-7", "f 1 2 3 junk",
        "g", "f 3/1/2 2/1/2 1/1/2"]) + "\r\n"
    for data in (text, text.replace("\r\n", "\n"), text.replace("\r\n", "\n").rstrip("\n")):      # CRLF, LF, no final newline
        host, dev, st = both(renderer, data, tmp_path, first_id=3)
        assert_same(host, dev, "syntax")
        assert st[2] >= 4                                         # the long / boundary / sub-normal literals went to strtof
    assert host[4] == 3 + 3 + 1


def test_object_ids_follow_the_tag_rule(renderer, tmp_path):
    tri = "v 0 0 0\nv 1 0 0\nv 0 1 0\n"
    for body in ("f 1 2 3\no a\nf 1 2 3\no b\nf 3 2 1\n", "o a\nf 1 2 3\ng b\nf 3 2 1\n", "f 1 2 3\nf 3 2 1\n", "o a\no b\nf 1 2 3\n", "f 1 2 3\no a\n"):
        host, dev, _ = both(renderer, tri + body, tmp_path, first_id=5)
        assert_same(host, dev, body)


@pytest.mark.parametrize("seed", range(12))
def test_random_files(renderer, tmp_path, seed):
    rng = np.random.default_rng(seed)
    nv, nn, nt_ = int(rng.integers(3, 40)), int(rng.integers(0, 12)), int(rng.integers(0, 6))
    lines, have_v, have_n, have_t = [], 0, 0, 0
    pending_v = [("v %s %s %s" % tuple(repr(float(np.float32(x))) if rng.random() < 0.5 else "%.*g" % (int(rng.integers(1, 12)), x) for x in rng.normal(size=3) * 10.0 ** rng.integers(-3, 4))) for _ in range(nv)]
    pending_n = ["vn %.5f %.5f %.5f" % tuple(rng.normal(size=3)) for _ in range(nn)]
    pending_t = ["vt %.4f %.4f" % tuple(rng.random(2)) for _ in range(nt_)]
    # all vertex data first (a stream that starts after the first face is an error, tested separately), faces and tags mixed after
    for block, kind in ((pending_v, "v"), (pending_t, "t"), (pending_n, "n")):
        lines += block
    have_v, have_n, have_t = nv, nn, nt_
    for _ in range(int(rng.integers(1, 60))):
        r = rng.random()
        if r < 0.08:
            lines.append(rng.choice(["o part", "g grp", "# c", "", "usemtl a", "s 1"]))
            continue
        k = 4 if rng.random() < 0.3 else 3
        style = int(rng.integers(0, 4))
        cs = []
        for _ in range(k):
            v = int(rng.integers(1, have_v + 1))
            if rng.random() < 0.2:
                v = v - have_v - 1
            t = int(rng.integers(1, have_t + 1)) if have_t else 1
            n = int(rng.integers(1, have_n + 2)) if have_n else 1          # sometimes one past the end: zero normal
            cs.append(["%d" % v, "%d/%d" % (v, t), "%d//%d" % (v, n), "%d/%d/%d" % (v, t, n)][style])
        lines.append("f " + (" " if rng.random() < 0.5 else "\t").join(cs))
    text = "\n".join(lines) + "\n"
    host, dev, _ = both(renderer, text, tmp_path, first_id=int(rng.integers(0, 4)))
    if isinstance(host, str):
        assert isinstance(dev, str), "host refused (%s), device accepted" % host
    else:
        assert_same(host, dev, "seed %d" % seed)


@pytest.mark.parametrize("body,needle", [
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n", "fewer than 3"),
    ("v 0 0 0\nv 1 0\nv 0 1 0\nf 1 2 3\n", "bad 'v' line"),
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0\nf 1 2 3\n", "bad 'vn' line"),
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n", "missing vertex"),
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 -4\n", "missing vertex"),
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\nvn 0 0 1\nf 1//1 2//1 3//1\n", "normal stream misaligned"),
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\nvt 0 0\n", "uv stream misaligned"),
    ("v 0 0 0\nv 1 0 0\nv 0 1 0\n", "no geometry"),
    ("# nothing\n", "no geometry"),
])
def test_refusals_match_the_host_loader(renderer, tmp_path, body, needle):
    host, dev, _ = both(renderer, body, tmp_path)
    assert isinstance(host, str) and needle in host
    assert isinstance(dev, str) and needle in dev


def test_forms_only_the_host_loader_takes_are_refused_loudly(renderer, tmp_path):
    for body in ("v 0 0 inf\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", "v 0x1p3 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3 #" + "x" * 1100 + "\n"):
        with pytest.raises(RtError) as e:
            DeviceMesh.parse_obj(renderer, body.encode())
        assert e.value.code == A.RT_ERR_UNSUPPORTED


def test_scene_from_device_mesh_renders_the_same_image(renderer, tmp_path):
    """Parse on the device -> rt_scene of device pointers (+ a transform baked on the device) -> the same frame as the host path."""
    g = np.load(os.path.join(GOLD, "frog_mesh.npz"))
    text = obj_text(g["positions"], g["indices"].astype(np.int64), g["normals"])
    path = tmp_path / "frog.obj"
    path.write_text(text)
    pos, nrm, idx, obj, _ = api.load_obj(str(path))
    xf = [(0, pos.shape[0], (0.1, -0.2, 0.05), (10.0, 20.0, -5.0), (1.1, 0.9, 1.0))]
    mats = [api.make_material(**scenes.FROG_MATERIAL)]
    fr = scenes.frog_frame(320, 180, filling=True, outputs=ALL)
    renderer.upload_scene(api.Scene(pos, idx, normals=nrm, tri_obj_ids=obj, materials=mats, transforms=xf))
    renderer.render(fr)
    want = renderer.download()
    dm, _ = DeviceMesh.parse_obj(renderer, text.encode())
    info = renderer.upload_scene(DeviceScene(dm, materials=mats, transforms=xf))
    assert info.num_triangles == idx.shape[0]
    renderer.render(fr)
    got = renderer.download()
    for k in ("tri_id", "t", "rgb", "rgb8"):
        assert np.array_equal(want[k], got[k]), k
    dm.close()


def test_append_on_the_device(renderer, tmp_path):
    a = "v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n"
    b = "v 0 0 1\nv 1 0 1\nv 0 1 1\nv 1 1 1\nf 1 2 3 4\n"
    (tmp_path / "a.obj").write_text(a)
    (tmp_path / "b.obj").write_text(b)
    lib = renderer.lib
    import ctypes as C
    hs = []
    nid = C.c_int32(0)
    for name in ("a.obj", "b.obj"):
        h = C.c_void_p()
        assert lib.rt_mesh_load_obj(os.fsencode(str(tmp_path / name)), C.byref(nid), C.byref(h)) == A.RT_OK
        hs.append(h)
    dst = C.c_void_p()
    lib.rt_mesh_create(C.byref(dst))
    for h in hs:
        lib.rt_mesh_append(dst, h)
    nv, nn, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.rt_mesh_counts(dst, C.byref(nv), C.byref(nn), C.byref(nt))
    hp, hn = np.zeros((nv.value, 3), np.float32), np.zeros((nn.value, 3), np.float32)
    hi, ho = np.zeros((nt.value, 3), np.uint32), np.zeros(nt.value, np.int32)
    lib.rt_mesh_copy(dst, api._ptr(hp, A.f32p), api._ptr(hn, A.f32p), api._ptr(hi, A.u32p), api._ptr(ho, A.i32p))
    for h in hs + [dst]:
        lib.rt_mesh_free(h)
    total = DeviceMesh(renderer)
    did = 0
    for t in (a, b):
        m, did = DeviceMesh.parse_obj(renderer, t.encode(), did)
        total.append(m)
        m.close()
    dp, dn, di, do = total.download()
    assert did == nid.value
    assert np.array_equal(hp, dp) and np.array_equal(hn, dn) and np.array_equal(hi, di) and np.array_equal(ho, do)
    total.close()


def test_terrain_sized_file(renderer, tmp_path):
    """200 000 triangles of the C4 terrain as text (8 MB): arrays identical, and the device parse is timed."""
    pos, idx = scenes.terrain(500, 200)
    text = obj_text(pos, idx.astype(np.int64), fmt="%.9g")
    host, dev, st = both(renderer, text, tmp_path)
    assert_same(host, dev, "terrain")
    assert host[2].shape[0] == 200000 and st[0] > 0
