#!/usr/bin/env python
"""packet_sim.py — CPU-side study of packet traversal strategies on the PRODUCT's BVH (design tool, not a test).

Builds the C4-style terrain BVH with the host emulation of the device build (tests/emul), then walks sampled 8x4-pixel
packets (primary rays + their point-light shadow rays) with
  (a) the shipped algorithm: every lane slab-tests both children of a node, the warp visits a child when any lane hits it
      (k_render_packet / packet_trace) — counts warp node visits and triangle blocks;
  (b) a frustum-culled wide traversal: the BVH2 is collapsed three levels at a time into 8-wide nodes, a round pops up to
      `batch` wide nodes and tests their <= 32 child boxes — one box per LANE — against the packet's bounding frustum
      (apex = camera centre or the light, four side planes + a depth range); leaves are tested per ray exactly as today.
It reports rounds / boxes / triangle blocks per packet trace so the instruction budget of (b) can be estimated before any
CUDA is written, and checks that (b) finds the same closest hits / occlusion as (a) (soundness of the frustum test).

  python tools/packet_sim.py [--nx 1000 --ny 500 --width 3840 --height 2160 --packets 300]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_bvh(nx, ny, leaf_max=2):
    import orclib
    from raytracinginonesemester_b200 import scenes
    sc = scenes.terrain_scene(nx, ny)
    h = orclib.emul_build(sc, leaf_max)
    lib = orclib.emul()
    lib.emu_num_nodes.restype = C.c_uint32
    nn = lib.emu_num_nodes(C.c_void_p(h))
    nodes = np.zeros((nn, 16), np.uint32)
    geom = np.zeros((sc.indices.shape[0], 12), np.float32)
    lib.emu_export(C.c_void_p(h), nodes.ctypes.data_as(C.c_void_p), geom.ctypes.data_as(C.c_void_p))
    q = nodes[:, :12].copy().view(np.float32).astype(np.float64)
    refs = nodes[:, 12:14].copy().view(np.int32)
    axis = nodes[:, 15] >> 29
    return q, refs, axis, geom.astype(np.float64), geom[:, 3].copy().view(np.int32)


class Bvh:
    def __init__(self, q, refs, axis, geom, ids):
        self.q, self.refs, self.axis, self.geom, self.ids = q, refs, axis, geom, ids
        self.wide = {}

    def child_box(self, n, k):
        q = self.q[n]
        return (q[0:3], q[3:6]) if k == 0 else (q[6:9], q[9:12])

    def wide_node(self, n, levels=3):
        """<= 2^levels entries (c, h, ref) reached by expanding node n `levels` levels."""
        w = self.wide.get(n)
        if w is None:
            ent = [(self.child_box(n, k), int(self.refs[n, k])) for k in (0, 1)]
            for _ in range(levels - 1):
                nxt = []
                for (box, ref) in ent:
                    if ref >= 0:
                        nxt += [(self.child_box(ref, k), int(self.refs[ref, k])) for k in (0, 1)]
                    else:
                        nxt.append((box, ref))
                ent = nxt
            ent = [e for e in ent if e[0][1][0] >= 0]          # absent children have negative half extents
            w = (np.array([e[0][0] for e in ent]), np.array([e[0][1] for e in ent]), np.array([e[1] for e in ent]))
            self.wide[n] = w
        return w


def leaf_range(ref):
    v = (~ref) & 0xFFFFFFFF
    return v >> 3, (v & 7) + 1


def mt(o, d, g, tmin, tmax):
    """Möller–Trumbore for 32 rays x 1 triangle (float64 — statistics only)."""
    v0, e1, e2 = g[0:3], g[4:7], g[8:11]
    p = np.cross(d, e2)
    det = p @ e1
    ok = np.abs(det) >= 1e-8
    inv = 1.0 / np.where(ok, det, 1.0)
    tv = o - v0
    u = np.einsum("ij,ij->i", tv, p) * inv
    qv = np.cross(tv, e1)
    v = np.einsum("ij,ij->i", d, qv) * inv
    t = (qv @ e2) * inv
    return ok & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t >= tmin) & (t <= tmax), t


def leaf_tests(b, ref, o, d, live, any_hit, tlim, best_id, blocked, st):
    first, cnt = leaf_range(ref)
    for s in range(first, first + cnt):
        st["blocks"] += 1
        hit, t = mt(o, d, b.geom[s], 1e-4, np.where(any_hit, np.inf, tlim))
        hit &= live
        if any_hit:
            nb = hit & (t < tlim)
            blocked |= nb
            live &= ~nb
        else:
            upd = hit & ((t < tlim) | (b.ids[s] < best_id))
            tlim[upd] = t[upd]
            best_id[upd] = b.ids[s]


def slab(o, inv, c, h, tmin, tmax):
    m = (c - o) * inv
    a = np.abs(inv) * h
    tn = np.maximum((m - a).max(axis=1), tmin)
    tf = np.minimum((m + a).min(axis=1), tmax)
    return tn <= tf


def trace_current(b, o, d, live, any_hit, tlim):
    """packet_trace of rt_trace.cu: returns stats, per-lane tlim/best id/blocked."""
    st = dict(visits=0, blocks=0)
    live = live.copy(); tlim = tlim.copy()
    best_id = np.full(32, 2**31 - 1); blocked = np.zeros(32, bool)
    with np.errstate(divide="ignore"):
        inv = 1.0 / d
    half = live.sum()
    dirneg = [2 * (live & (d[:, k] < 0)).sum() > half for k in (2, 1, 0)]       # bit 0 = z, 1 = y, 2 = x
    stack, cur = [], 0
    while True:
        if cur >= 0:
            st["visits"] += 1
            q = b.q[cur]
            a0 = (slab(o, inv, q[0:3], q[3:6], 1e-4, tlim) & live).any()
            a1 = (slab(o, inv, q[6:9], q[9:12], 1e-4, tlim) & live).any()
            r0, r1 = int(b.refs[cur, 0]), int(b.refs[cur, 1])
            if a0 and a1:
                ax = int(b.axis[cur])
                c1first = any(dirneg[k] for k in range(3) if ax >> k & 1)
                stack.append(r0 if c1first else r1)
                cur = r1 if c1first else r0
                continue
            if a0: cur = r0; continue
            if a1: cur = r1; continue
        else:
            leaf_tests(b, cur, o, d, live, any_hit, tlim, best_id, blocked, st)
            if any_hit and not live.any():
                break
        if not stack:
            break
        cur = stack.pop()
    return st, tlim, best_id, blocked


def frustum(o, d, live):
    """Four bounding planes + a depth axis for rays o_i + t d_i, t >= 0 (no common apex needed): with w the mean direction
    and slopes sa_i = (d_i.u)/(d_i.w), every ray satisfies x.n >= min_i(o_i.n) for n = u - min(sa) w (the coefficient of t
    is >= 0), and likewise for the other three sides."""
    w = d[live].mean(axis=0); w /= np.linalg.norm(w)
    t = np.array([1.0, 0, 0]) if abs(w[0]) < 0.7 else np.array([0, 1.0, 0])
    u = np.cross(w, t); u /= np.linalg.norm(u)
    v = np.cross(w, u)
    dw = d @ w
    if (dw[live] <= 1e-3).any():
        return None
    sa, sb = (d @ u) / dw, (d @ v) / dw
    m = 1e-6
    planes = np.array([u - (sa[live].min() - m) * w, (sa[live].max() + m) * w - u, v - (sb[live].min() - m) * w, (sb[live].max() + m) * w - v])
    off = (o[live] @ planes.T).min(axis=0) - 1e-7
    return dict(w=w, planes=planes, off=off, dw=dw, ow=o @ w)


def frustum_test(fr, c, h, near, far):
    s = c @ fr["planes"].T + h @ np.abs(fr["planes"]).T          # (n, 4): max of each plane function over the box
    dep = c @ fr["w"]
    dh = h @ np.abs(fr["w"])
    return (s >= fr["off"]).all(axis=1) & (dep + dh >= near) & (dep - dh <= far), dep - dh


def trace_wide(b, o, d, live, any_hit, tlim, batch=4, order="depth", levels=3):
    """Frustum-culled wide traversal; lanes = child boxes.  Returns stats + results."""
    st = dict(rounds=0, boxes=0, blocks=0, leaves=0, fallback=0)
    live = live.copy(); tlim = tlim.copy()
    best_id = np.full(32, 2**31 - 1); blocked = np.zeros(32, bool)
    fr = frustum(o, d, live)
    if fr is None:
        st["fallback"] = 1
        s2, tlim, best_id, blocked = trace_current(b, o, d, live, any_hit, tlim)
        st["rounds"] = s2["visits"]; st["blocks"] = s2["blocks"]
        return st, tlim, best_id, blocked

    def far_depth():
        if not live.any():
            return -1e30
        t = np.where(np.isfinite(tlim), tlim, 1e30)
        return (fr["ow"] + t * fr["dw"])[live].max() * (1 + 1e-6) + 1e-6
    near0 = fr["ow"][live].min() - 1e-6
    stack = [0]
    while stack:
        far = far_depth()
        take = [stack.pop() for _ in range(min(batch, len(stack)))]
        cs, hs, rs = zip(*[b.wide_node(n, levels) for n in take])
        c, h, r = np.concatenate(cs), np.concatenate(hs), np.concatenate(rs)
        st["rounds"] += 1; st["boxes"] += len(r)
        hit, near = frustum_test(fr, c, h, near0, far)
        idx = np.nonzero(hit)[0]
        idx = idx[np.argsort(near[idx])] if order == "depth" else idx
        for i in idx:                                   # leaves first, near to far
            if r[i] < 0:
                if near[i] > far:
                    continue
                st["leaves"] += 1
                leaf_tests(b, int(r[i]), o, d, live, any_hit, tlim, best_id, blocked, st)
                if any_hit and not live.any():
                    return st, tlim, best_id, blocked
                far = far_depth()
        for i in idx[::-1]:                             # far to near, so the nearest ends up on top
            if r[i] >= 0 and near[i] <= far:
                stack.append(int(r[i]))
    return st, tlim, best_id, blocked


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1000); ap.add_argument("--ny", type=int, default=500)
    ap.add_argument("--width", type=int, default=3840); ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--packets", type=int, default=200); ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--levels", type=int, default=3); ap.add_argument("--order", default="depth")
    ap.add_argument("--tile", default="8x4", help="pixel footprint of a packet, WxH with W*H = 32")
    args = ap.parse_args()
    from raytracinginonesemester_b200 import scenes
    b = Bvh(*load_bvh(args.nx, args.ny))
    fr = scenes.terrain_frame(args.width, args.height)
    cam = fr.cam
    c3 = lambda v: np.array(list(v), np.float64)
    cen, p00, du, dv = c3(cam.center), c3(cam.pixel00_loc), c3(cam.pixel_delta_u), c3(cam.pixel_delta_v)
    jit = np.asarray(fr.jitter, np.float64).reshape(-1, 2)[0]
    L = np.array([-2.0, -1.0, 1.5])
    rng = np.random.default_rng(1)
    tot = {k: np.zeros(2) for k in ("cur_visits", "cur_blocks", "w_rounds", "w_boxes", "w_blocks", "w_leaves", "fallback", "n")}
    mism = 0
    for _ in range(args.packets):
        tw, th = [int(v) for v in args.tile.split("x")]
        tx, ty = rng.integers(0, args.width // tw), rng.integers(0, args.height // th)
        xs, ys = np.meshgrid(tx * tw + np.arange(tw), ty * th + np.arange(th))
        px, py = xs.ravel() + jit[0], ys.ravel() + jit[1]
        d = p00 + px[:, None] * du + py[:, None] * dv - cen
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        o = np.broadcast_to(cen, d.shape).copy()
        live = np.ones(32, bool)
        s1, t1, id1, _ = trace_current(b, o, d, live, False, np.full(32, np.inf))
        s2, t2, id2, _ = trace_wide(b, o, d, live, False, np.full(32, np.inf), args.batch, args.order, args.levels)
        mism += int((id1 != id2).sum())
        for k, v in (("cur_visits", s1["visits"]), ("cur_blocks", s1["blocks"]), ("w_rounds", s2["rounds"]), ("w_boxes", s2["boxes"]),
                     ("w_blocks", s2["blocks"]), ("w_leaves", s2["leaves"]), ("fallback", s2["fallback"]), ("n", 1)):
            tot[k][0] += v
        # shadow rays of the hit lanes (IsInShadow: origin P + N*1e-3, direction to the light, blocked iff t < dist)
        hit = np.isfinite(t1)
        if not hit.any():
            continue
        P = o + d * np.where(hit, t1, 0.0)[:, None]
        N = np.zeros_like(P)
        for i in np.nonzero(hit)[0]:
            s = int(np.nonzero(b.ids == id1[i])[0][0]) if False else None
        # geometric normal from the hit triangle: find slot by id via a lookup table
        slot_of = getattr(b, "slot_of", None)
        if slot_of is None:
            slot_of = np.zeros(b.ids.max() + 1, np.int64); slot_of[b.ids] = np.arange(len(b.ids)); b.slot_of = slot_of
        g = b.geom[slot_of[np.where(hit, id1, 0)]]
        n = np.cross(g[:, 4:7], g[:, 8:11]); n /= np.linalg.norm(n, axis=1, keepdims=True)
        n[np.einsum("ij,ij->i", n, d) > 0] *= -1
        toL = L - P
        dist = np.linalg.norm(toL, axis=1)
        sd = toL / dist[:, None]
        need = hit & (np.einsum("ij,ij->i", n, sd) > 0)
        if not need.any():
            continue
        so = P + n * 1e-3
        s1, _, _, b1 = trace_current(b, so, sd, need, True, dist)
        s2, _, _, b2 = trace_wide(b, so, sd, need, True, dist, args.batch, args.order, args.levels)
        mism += int((b1 != b2).sum())
        for k, v in (("cur_visits", s1["visits"]), ("cur_blocks", s1["blocks"]), ("w_rounds", s2["rounds"]), ("w_boxes", s2["boxes"]),
                     ("w_blocks", s2["blocks"]), ("w_leaves", s2["leaves"]), ("fallback", s2["fallback"]), ("n", 1)):
            tot[k][1] += v
    for j, name in enumerate(("primary", "shadow")):
        n = max(tot["n"][j], 1)
        print("%-8s packets %4d | shipped: %.1f node visits, %.1f tri blocks | wide(frustum, %d levels, batch %d, %s): %.1f rounds, %.1f boxes (%.1f / round), %.1f leaves, %.1f tri blocks, fallback %.0f"
              % (name, n, tot["cur_visits"][j] / n, tot["cur_blocks"][j] / n, args.levels, args.batch, args.order,
                 tot["w_rounds"][j] / n, tot["w_boxes"][j] / n, tot["w_boxes"][j] / max(tot["w_rounds"][j], 1), tot["w_leaves"][j] / n,
                 tot["w_blocks"][j] / n, tot["fallback"][j]))
    print("lanes whose result differs between the two traversals:", mism)


if __name__ == "__main__":
    main()
