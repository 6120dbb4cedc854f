// rt_build_core.h — per-element steps of the BVH build, host/device compilable: 63-bit
// Morton keys, Karras-2012 hierarchy emission, leaf padding, subtree collapse, node packing.
#pragma once

#include "rt_params.h"

struct Bounds { float lo[3]; float hi[3]; };
// Karras numbering: internal k -> k, leaf k -> (n-1)+k.  [first,last] = sorted-order range.
struct Topo { uint32_t left, right, first, last; };   // last: bits 0..29 = index, bits 30..31 = split axis (0 z, 1 y, 2 x, 3 none)
#define RT_TOPO_LAST(tp) ((tp).last & 0x3fffffffu)
#define RT_TOPO_AXIS(tp) ((tp).last >> 30)
// Node word slot_count: bits 0..28 = triangle slots of the subtree, bits 29..31 = one-hot split axis (29 z, 30 y, 31 x;
// none set = no spatial split), so the packet kernel's descent order is one AND with its direction-sign mask.
#define RT_NODE_AXIS_BITS(axis) ((axis) < 3u ? (1u << (29u + (axis))) : 0u)
// Node word first_slot: bits 0..27 = first triangle slot of the subtree (slots < 2^28, rt_leaf_ref), bits 28..29 = the node's
// depth modulo 3, which the compact 8-wide view is derived from (rt_build_wide).
#define RT_NODE_DEPTH3_BITS(depth) ((uint32_t)((depth) % 3u) << 28)
#define RT_NODE_DEPTH3(first_slot) (((first_slot) >> 28) & 3u)

RT_HD void rt_tri_verts(const BuildParams& bp, uint32_t i, f3& a, f3& b, f3& c, uint32_t& ia, uint32_t& ib, uint32_t& ic) {
    ia = RT_LDG(bp.indices + 3 * (size_t)i); ib = RT_LDG(bp.indices + 3 * (size_t)i + 1); ic = RT_LDG(bp.indices + 3 * (size_t)i + 2);
    a = ld3(bp.positions + 3 * (size_t)ia); b = ld3(bp.positions + 3 * (size_t)ib); c = ld3(bp.positions + 3 * (size_t)ic);
}
RT_HD void rt_tri_box(f3 a, f3 b, f3 c, float lo[3], float hi[3]) {
    lo[0] = fminf(a.x, fminf(b.x, c.x)); lo[1] = fminf(a.y, fminf(b.y, c.y)); lo[2] = fminf(a.z, fminf(b.z, c.z));
    hi[0] = fmaxf(a.x, fmaxf(b.x, c.x)); hi[1] = fmaxf(a.y, fmaxf(b.y, c.y)); hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
}

RT_HD uint64_t rt_expand21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

// Morton key of the box centroid, 21 bits per axis (the reference uses 10: bvh.h:142-151).
RT_HD uint64_t rt_morton63(f3 a, f3 b, f3 c, const Bounds& scene) {
    float lo[3], hi[3];
    rt_tri_box(a, b, c, lo, hi);
    const float R = 2097152.0f;
    // One scale for all three axes (the scene's longest extent): a thin axis (a terrain's height)
    // then keeps its top bits constant instead of slicing the mesh into interleaved layers, which is
    // what per-axis normalisation (the reference's, bvh.cu:44-47) does.
    const float e = fmaxf(scene.hi[0] - scene.lo[0], fmaxf(scene.hi[1] - scene.lo[1], scene.hi[2] - scene.lo[2]));
    uint64_t q[3];
    for (int k = 0; k < 3; ++k) {
        const float ce = 0.5f * (lo[k] + hi[k]);
        const float nrm = e > 0.f ? (ce - scene.lo[k]) / e : 0.f;
        q[k] = (uint64_t)fminf(fmaxf(nrm * R, 0.f), R - 1.f);
    }
    return (rt_expand21(q[0]) << 2) | (rt_expand21(q[1]) << 1) | rt_expand21(q[2]);
}

// Common-prefix length with index tie-break (Karras 2012, section 4).
RT_HD int rt_delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + RT_CLZ32((uint32_t)(i ^ j));
    return RT_CLZ64(a ^ b);
}

// Internal node i of the radix tree over n sorted keys (replaces determine_range + find_split,
// GPUandCPU/include/bvh.h:163-257).
RT_HD Topo rt_karras_node(const uint64_t* __restrict__ keys, int n, int i) {
    int d = (rt_delta(keys, n, i, i + 1) - rt_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    if (i == 0) d = 1;
    const int dmin = rt_delta(keys, n, i, i - d);
    int lmax = 2;
    while (rt_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (rt_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = rt_delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (rt_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int first = i < j ? i : j, last = i < j ? j : i;
    Topo tp;
    tp.left = (first == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
    tp.right = (last == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    // Split axis: the highest key bit in which the two children's ranges differ (key bit p belongs to
    // axis p % 3: 2 = x, 1 = y, 0 = z).  The right child holds the larger coordinates along it, so a
    // ray travelling in the negative direction of that axis meets the right child first.
    const uint64_t ka = keys[gamma], kb = keys[gamma + 1];
    const uint32_t axis = (ka == kb) ? 3u : (uint32_t)((63 - RT_CLZ64(ka ^ kb)) % 3);
    tp.first = (uint32_t)first; tp.last = (uint32_t)last | (axis << 30);
    return tp;
}

// Leaf box, padded outward by 2^-17 of the local/scene scale so the fp32 slab test stays a
// superset of what the fp32 Möller–Trumbore test accepts (rt_core.h, rt_slab).
RT_HD void rt_padded_leaf_box(f3 a, f3 b, f3 c, const Bounds& scene, float lo[3], float hi[3]) {
    rt_tri_box(a, b, c, lo, hi);
    const float ext = fmaxf(scene.hi[0] - scene.lo[0], fmaxf(scene.hi[1] - scene.lo[1], scene.hi[2] - scene.lo[2]));
    for (int q = 0; q < 3; ++q) {
        const float pad = fmaxf(ext, fmaxf(fabsf(lo[q]), fabsf(hi[q]))) * 7.6293945e-06f + 1e-30f;   // 2^-17
        lo[q] -= pad; hi[q] += pad;
    }
}

RT_HD int32_t rt_child_ref(uint32_t c, int n, const Topo* __restrict__ topo, const uint32_t* __restrict__ keep,
                           const uint32_t* __restrict__ newidx) {
    if (c >= (uint32_t)(n - 1)) return rt_leaf_ref(c - (uint32_t)(n - 1), 1u);     // original leaf
    if (keep[c]) return (int32_t)newidx[c];
    return rt_leaf_ref(topo[c].first, RT_TOPO_LAST(topo[c]) - topo[c].first + 1u);   // collapsed subtree
}

// lo/hi -> centre + half-extent, half-extent rounded up so that [c-h, c+h] contains [lo, hi].
RT_HD void rt_box_to_ch(float lo, float hi, float& c, float& h) {
    if (!(lo <= hi)) { c = 0.f; h = -1.f; return; }           // absent child
    c = 0.5f * lo + 0.5f * hi;
    const float e = fmaxf(hi - c, c - lo);
    h = e * 1.000001f + 1e-30f + fabsf(c) * 1.2e-7f;
}
// Depth of Karras internal node i (root = 0) by its parent links; a kept node's ancestors are all kept, so this is also its
// depth in the emitted tree.
RT_HD uint32_t rt_node_depth(const uint32_t* __restrict__ parent, uint32_t i) {
    uint32_t d = 0;
    for (uint32_t p = parent[i]; p != 0xFFFFFFFFu; p = parent[p]) ++d;
    return d;
}

RT_HD BvhNode rt_make_node(float4 l0, float4 h0, float4 l1, float4 h1, int32_t r0, int32_t r1, uint32_t first, uint32_t count) {
    BvhNode nd;
    rt_box_to_ch(l0.x, h0.x, nd.q[0], nd.q[3]); rt_box_to_ch(l0.y, h0.y, nd.q[1], nd.q[4]); rt_box_to_ch(l0.z, h0.z, nd.q[2], nd.q[5]);
    rt_box_to_ch(l1.x, h1.x, nd.q[6], nd.q[9]); rt_box_to_ch(l1.y, h1.y, nd.q[7], nd.q[10]); rt_box_to_ch(l1.z, h1.z, nd.q[8], nd.q[11]);
    nd.ref0 = r0; nd.ref1 = r1; nd.first_slot = first; nd.slot_count = count;   // count: bits 29..31 carry the one-hot split axis (RT_NODE_AXIS_BITS)
    return nd;
}

// 8-wide view of BVH2 node i (WideNode, rt_core.h): entry k = the box reached by the left/right steps k2 k1 k0; a leaf met
// early sits in the entry whose remaining path bits are zero; everything else below it, and absent children, are "absent"
// entries (negative half extents, ref -1).  `levels` < 3 stops the expansion early (the root of a phased view, rt_build.cu):
// an inner box reached at that level is placed like an early leaf and keeps its (non-negative) node reference.
RT_HD WideNode rt_wide_node(const BvhNode* __restrict__ nodes, uint32_t i, int levels = 3) {
    WideNode out;
    for (int k = 0; k < 8; ++k) {
        bool valid = true;
        uint32_t par = i;
        int pb = (k >> 2) & 1;
        int ref = pb ? nodes[par].ref1 : nodes[par].ref0;
        if (ref < 0 || levels == 1) valid = (k & 3) == 0;
        else {
            par = (uint32_t)ref; pb = (k >> 1) & 1; ref = pb ? nodes[par].ref1 : nodes[par].ref0;
            if (ref < 0 || levels == 2) valid = (k & 1) == 0;
            else { par = (uint32_t)ref; pb = k & 1; ref = pb ? nodes[par].ref1 : nodes[par].ref0; }
        }
        const float* q = nodes[par].q + 6 * pb;
        WideEntry e;
        e.cx = q[0]; e.cy = q[1]; e.cz = q[2]; e.hx = q[3]; e.hy = q[4]; e.hz = q[5]; e.ref = ref; e.pad = 0;
        if (!valid || !(e.hx >= 0.f)) { e.cx = e.cy = e.cz = 0.f; e.hx = e.hy = e.hz = -1.f; e.ref = -1; }
        out.e[k] = e;
    }
    return out;
}

// Triangle blocks of slot k (triangle `tri`).  e1/e2 are the same rounded differences the
// reference forms inside every test (query.h:80-81, HW1 ray.h:72-73).  The geometry block also carries the object id
// (g[1].w); the normals block exists only for meshes with per-vertex normals (shade_k == nullptr otherwise).
RT_HD void rt_pack_tri(const BuildParams& bp, uint32_t tri, TriBlock* geom_k, TriBlock* shade_k) {
    f3 a, b, c; uint32_t ia, ib, ic;
    rt_tri_verts(bp, tri, a, b, c, ia, ib, ic);
    const f3 e1 = xsub3(b, a), e2 = xsub3(c, a);
    const int obj = bp.obj_ids ? bp.obj_ids[tri] : -1;
    float4* g = reinterpret_cast<float4*>(geom_k);
    g[0] = make_float4(a.x, a.y, a.z, RT_I2F((int)tri));
    g[1] = make_float4(e1.x, e1.y, e1.z, RT_I2F(obj));
    g[2] = make_float4(e2.x, e2.y, e2.z, 0.f);
    if (shade_k == nullptr) return;
    f3 n0 = mk3(0.f, 0.f, 0.f), n1 = n0, n2 = n0;
    if (bp.normals) {
        n0 = ld3(bp.normals + 3 * (size_t)ia); n1 = ld3(bp.normals + 3 * (size_t)ib); n2 = ld3(bp.normals + 3 * (size_t)ic);
    }
    float4* s = reinterpret_cast<float4*>(shade_k);
    s[0] = make_float4(n0.x, n0.y, n0.z, RT_I2F(obj));
    s[1] = make_float4(n1.x, n1.y, n1.z, 0.f);
    s[2] = make_float4(n2.x, n2.y, n2.z, 0.f);
}

// -------------------------------------------------------- per-object transform bake ----
// applyObjectTransform / rotateXYZ, GPUandCPU/src/main.cu:52-96.  The trigonometric values come from the host.
struct BakeXform { float sx, sy, sz, cx_, sx_, cy_, sy_, cz_, sz_, tx, ty, tz; };   // scale; cos/sin about X, Y, Z; translation
RT_HD f3 rt_rotate_xyz(f3 v, const BakeXform& T) {
    const float y1 = XSUB(XMUL(T.cx_, v.y), XMUL(T.sx_, v.z)), z1 = XADD(XMUL(T.sx_, v.y), XMUL(T.cx_, v.z));      // about X
    const float x2 = XADD(XMUL(T.cy_, v.x), XMUL(T.sy_, z1)), z2 = XADD(XMUL(-T.sy_, v.x), XMUL(T.cy_, z1));       // about Y
    const float x3 = XSUB(XMUL(T.cz_, x2), XMUL(T.sz_, y1)), y3 = XADD(XMUL(T.sz_, x2), XMUL(T.cz_, y1));          // about Z
    return mk3(x3, y3, z2);
}
RT_HD f3 rt_bake_point(f3 p, const BakeXform& T) {
    const f3 r = rt_rotate_xyz(mk3(XMUL(p.x, T.sx), XMUL(p.y, T.sy), XMUL(p.z, T.sz)), T);
    return mk3(XADD(r.x, T.tx), XADD(r.y, T.ty), XADD(r.z, T.tz));
}
RT_HD f3 rt_bake_normal(f3 n, const BakeXform& T) {
    if (fabsf(T.sx) > 1e-8f) n.x = XDIV(n.x, T.sx);
    if (fabsf(T.sy) > 1e-8f) n.y = XDIV(n.y, T.sy);
    if (fabsf(T.sz) > 1e-8f) n.z = XDIV(n.z, T.sz);
    const f3 r = rt_rotate_xyz(n, T);
    const float len2 = xdot(r, r);
    if (len2 > 1e-12f) return xmuls(r, XDIV(1.0f, XSQRT(len2)));
    return mk3(0.0f, 0.0f, 1.0f);
}
