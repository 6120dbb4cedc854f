"""CPU check of the ALGORITHM behind the default frame kernel's traversal (csrc/rt_trace.cu, frustum_trace): on the BVH
the host emulation of the device build produces, the frustum-culled wide traversal (8-wide view, four bounding planes that
need no common apex, depth range, per-ray leaf cull) must find exactly the closest hits / occlusion flags of the per-lane
packet traversal, for camera packets and for their point-light shadow packets.  tools/packet_sim.py is the float64 model
both kernels were designed with; the device code itself is covered by the GPU tier (test_gpu_parity.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("width,height,levels,batch", [(320, 180, 3, 4), (96, 54, 3, 4), (640, 360, 2, 8)])
def test_frustum_traversal_is_sound_on_terrain(width, height, levels, batch):
    import packet_sim as ps
    from raytracinginonesemester_b200 import scenes
    b = ps.Bvh(*ps.load_bvh(60, 30))
    fr = scenes.terrain_frame(width, height)
    cam = fr.cam
    c3 = lambda v: np.array(list(v), np.float64)
    cen, p00, du, dv = c3(cam.center), c3(cam.pixel00_loc), c3(cam.pixel_delta_u), c3(cam.pixel_delta_v)
    L = np.array([-2.0, -1.0, 1.5])
    slot_of = np.zeros(b.ids.max() + 1, np.int64)
    slot_of[b.ids] = np.arange(len(b.ids))
    rng = np.random.default_rng(7)
    rounds = visits = 0
    for _ in range(25):
        tx, ty = rng.integers(0, width // 8), rng.integers(0, height // 4)
        xs, ys = np.meshgrid(tx * 8 + np.arange(8), ty * 4 + np.arange(4))
        d = p00 + (xs.ravel() + 0.1)[:, None] * du + (ys.ravel() - 0.2)[:, None] * dv - cen
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        o = np.broadcast_to(cen, d.shape).copy()
        live = np.ones(32, bool)
        live[rng.integers(0, 32)] = False                       # a dead lane (pixel outside the frame)
        s1, t1, id1, _ = ps.trace_current(b, o, d, live, False, np.full(32, np.inf))
        s2, t2, id2, _ = ps.trace_wide(b, o, d, live, False, np.full(32, np.inf), batch, "none", levels)
        assert np.array_equal(id1[live], id2[live]) and np.array_equal(t1[live], t2[live])
        rounds += s2["rounds"]; visits += s1["visits"]
        hit = np.isfinite(t1) & live
        if not hit.any():
            continue
        P = o + d * np.where(hit, t1, 0.0)[:, None]
        g = b.geom[slot_of[np.where(hit, id1, 0)]]
        n = np.cross(g[:, 4:7], g[:, 8:11]); n /= np.linalg.norm(n, axis=1, keepdims=True)
        n[np.einsum("ij,ij->i", n, d) > 0] *= -1
        toL = L - P
        dist = np.linalg.norm(toL, axis=1)
        sd = toL / dist[:, None]
        need = hit & (np.einsum("ij,ij->i", n, sd) > 0)
        if need.any():
            _, _, _, b1 = ps.trace_current(b, P + n * 1e-3, sd, need, True, dist)
            _, _, _, b2 = ps.trace_wide(b, P + n * 1e-3, sd, need, True, dist, batch, "none", levels)
            assert np.array_equal(b1[need], b2[need])
    assert rounds < visits          # and it gets there in fewer steps than one node visit at a time


@pytest.mark.parametrize("nx,ny,leaf_max", [(60, 30, 2), (17, 9, 1), (8, 4, 4), (1, 1, 2)])
def test_wide_view_of_the_bvh(nx, ny, leaf_max):
    """rt_wide_node (what k_build_wide runs per BVH2 node) against an independent expansion of the same nodes: entry k is
    the box three left/right steps down (path bits k2 k1 k0), a leaf met early sits where the remaining bits are zero,
    everything else is absent; and walking the wide view from the root reaches every triangle slot exactly once."""
    import ctypes as C
    import orclib
    import packet_sim as ps
    from raytracinginonesemester_b200 import scenes
    sc = scenes.terrain_scene(nx, ny)
    h = orclib.emul_build(sc, leaf_max)
    lib = orclib.emul()
    lib.emu_num_nodes.restype = C.c_uint32
    nn = lib.emu_num_nodes(C.c_void_p(h))
    nodes = np.zeros((nn, 16), np.uint32)
    lib.emu_export(C.c_void_p(h), nodes.ctypes.data_as(C.c_void_p), None)
    wide = np.zeros((nn, 8, 8), np.uint32)
    lib.emu_wide(C.c_void_p(h), wide.ctypes.data_as(C.c_void_p))
    q = nodes[:, :12].view(np.float32)
    refs = nodes[:, 12:14].view(np.int32)
    wf, wref = wide[..., :6].view(np.float32), wide[..., 6].view(np.int32)

    def child(n, k):
        return q[n, 6 * k:6 * k + 6], int(refs[n, k])
    for i in range(nn):
        for k in range(8):
            box, ref = child(i, (k >> 2) & 1)
            ok = True
            if ref >= 0:
                box, ref2 = child(ref, (k >> 1) & 1)
                if ref2 >= 0:
                    box, ref = child(ref2, k & 1)
                else:
                    ok, ref = (k & 1) == 0, ref2
            else:
                ok = (k & 3) == 0
            if ok and box[3] >= 0:
                assert np.array_equal(wf[i, k], box) and wref[i, k] == ref, (i, k)
            else:
                assert wf[i, k, 3] < 0 and wref[i, k] == -1, (i, k)
    seen = np.zeros(sc.indices.shape[0], int)
    stack = [0]
    while stack:
        n = stack.pop()
        for k in range(8):
            if wf[n, k, 3] < 0:
                continue
            r = int(wref[n, k])
            if r >= 0:
                stack.append(r)
            else:
                first, cnt = ps.leaf_range(r)
                seen[first:first + cnt] += 1
    assert (seen == 1).all()


@pytest.mark.parametrize("nx,ny,leaf_max", [(60, 30, 2), (17, 9, 1), (8, 4, 4), (1, 1, 2)])
def test_compact_wide_view_in_every_phase(nx, ny, leaf_max):
    """The view the product keeps (rt_build_wide): wide nodes only for the BVH2 nodes at every third depth, the root hopping
    `phase` levels.  Each node's depth-mod-3 tag is its real depth; in every phase the walk from wide node 0 stays inside the
    compact array, reaches every triangle slot exactly once and finds the same leaf boxes as the BVH2; the automatic choice
    is the smallest of the three."""
    import ctypes as C
    import orclib
    import packet_sim as ps
    from raytracinginonesemester_b200 import scenes
    sc = scenes.terrain_scene(nx, ny)
    h = orclib.emul_build(sc, leaf_max)
    lib = orclib.emul()
    lib.emu_num_nodes.restype = C.c_uint32
    lib.emu_wide_compact.restype = C.c_uint32
    nn = lib.emu_num_nodes(C.c_void_p(h))
    nodes = np.zeros((nn, 16), np.uint32)
    lib.emu_export(C.c_void_p(h), nodes.ctypes.data_as(C.c_void_p), None)
    q = nodes[:, :12].view(np.float32)
    refs = nodes[:, 12:14].view(np.int32)
    tag = (nodes[:, 14] >> 28) & 3
    depth = np.full(nn, -1)
    depth[0] = 0
    leaves = {}
    stack = [0]
    while stack:
        n = stack.pop()
        for k in range(2):
            if q[n, 6 * k + 3] < 0:
                continue
            r = int(refs[n, k])
            if r >= 0:
                depth[r] = depth[n] + 1
                stack.append(r)
            else:
                leaves[r] = q[n, 6 * k:6 * k + 6].copy()
    assert (depth >= 0).all() and np.array_equal(tag, depth % 3)
    sizes = []
    for phase in (0, 1, 2, -1):
        wide = np.zeros((nn, 8, 8), np.uint32)
        ph = C.c_int(phase)
        cnt = lib.emu_wide_compact(C.c_void_p(h), C.byref(ph), wide.ctypes.data_as(C.c_void_p))
        assert cnt != 0xFFFFFFFF and cnt <= nn
        if phase >= 0:
            sizes.append(cnt)
            assert cnt == int((depth % 3 == phase).sum()) + (phase != 0)
        else:
            assert cnt == min(sizes) and sizes[ph.value] == cnt
        wf, wref = wide[..., :6].view(np.float32), wide[..., 6].view(np.int32)
        seen = np.zeros(sc.indices.shape[0], int)
        visited = np.zeros(cnt, int)
        stack = [0]
        while stack:
            n = stack.pop()
            visited[n] += 1
            for k in range(8):
                if wf[n, k, 3] < 0:
                    assert wref[n, k] == -1
                    continue
                r = int(wref[n, k])
                if r >= 0:
                    assert r < cnt
                    stack.append(r)
                else:
                    first, c = ps.leaf_range(r)
                    seen[first:first + c] += 1
                    assert np.array_equal(wf[n, k], leaves[r])
        assert (seen == 1).all() and (visited == 1).all()
