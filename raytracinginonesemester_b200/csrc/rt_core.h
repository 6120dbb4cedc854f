// rt_core.h — data layout in HBM and the per-ray device functions (ray generation,
// Möller–Trumbore under the three reference contracts, slab test, shading, quantisers).
// Host-compilable (RT_HD) so the same code can be exercised by tests without a GPU;
// kernels live in rt_trace.cu / rt_build.cu.
#pragma once

#include <float.h>
#include "rt_math.h"
#include "../../include/rt_api.h"

// ----------------------------------------------------------------- layout ----
// One BVH2 node = one 64-byte line: both children's boxes + both child references, so a
// node visit is two 32-byte loads from a single aligned line (ld.global.nc.v8 = LDG.256 on sm_100).
// Boxes are stored as centre + half-extent: the slab distances are then m -+ h*|1/d| with
// m = (c - o)/d, i.e. nine FMAs per box and no per-axis min/max (those run on the half-rate
// ALU pipe, which ncu showed to be the busiest pipe of the first packet kernel).
//   q0 = (c0.x, c0.y, c0.z, h0.x)   q1 = (h0.y, h0.z, c1.x, c1.y)
//   q2 = (c1.z, h1.x, h1.y, h1.z)   q3 = (ref0, ref1, first_slot, slot_count) as ints
// Child reference: >= 0 -> index of another node; < 0 -> leaf, ~ref = (first_slot << 3) | (count-1).
// An absent child has negative half-extents (never hit; if NaNs from a zero direction component
// make it pass, its reference is a valid one-triangle leaf, and testing an extra triangle cannot
// change a closest-hit result).
struct alignas(64) BvhNode {
    float q[12];
    int32_t ref0, ref1;
    uint32_t first_slot, slot_count;   // triangle slots spanned by this subtree (debug / stats)
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be one 64-byte line");

// 8-wide view of the same tree for the frustum traversal (rt_trace.cu, frustum_trace): a wide node holds the boxes and
// references reached from one BVH2 node by three left/right steps (entry k = path bits k2 k1 k0; a leaf met early sits in
// the entry whose remaining path bits are zero, the other entries of that subtree are absent).  32 bytes per entry:
// (c.x, c.y, c.z, h.x) (h.y, h.z, bits(ref), 0); absent: h = -1.  Leaf references are those of BvhNode; an inner reference
// is the index of the next WIDE node: only the BVH2 nodes at every third depth get one, in a compact array
// (rt_build_wide, rt_build.cu; rt_wide_node below leaves BVH2 indices, the build translates them).
struct alignas(32) WideEntry { float cx, cy, cz, hx, hy, hz; int32_t ref; int32_t pad; };
struct alignas(256) WideNode { WideEntry e[8]; };
static_assert(sizeof(WideNode) == 256, "WideNode must be 256 bytes");

#define RT_LEAF_MAX_LOG2 3
RT_HD int32_t rt_leaf_ref(uint32_t first, uint32_t count) { return ~(int32_t)((first << RT_LEAF_MAX_LOG2) | (count - 1u)); }
RT_HD uint32_t rt_leaf_first(int32_t ref) { return ((uint32_t)~ref) >> RT_LEAF_MAX_LOG2; }
RT_HD uint32_t rt_leaf_count(int32_t ref) { return (((uint32_t)~ref) & ((1u << RT_LEAF_MAX_LOG2) - 1u)) + 1u; }

// Triangle "geometry block", 48 bytes, stored in BVH leaf order (slot index):
//   g0 = (v0.x, v0.y, v0.z, bits(original triangle id))
//   g1 = (e1.x, e1.y, e1.z, 0)   e1 = fl(v1 - v0)   (the reference recomputes the same
//   g2 = (e2.x, e2.y, e2.z, 0)   e2 = fl(v2 - v0)    rounded differences per test)
// Triangle "shading block", 48 bytes, same slot: (n0, bits(object id)), (n1, 0), (n2, 0).
struct alignas(16) TriBlock { float g[12]; };
static_assert(sizeof(TriBlock) == 48, "TriBlock must be 48 bytes");

struct Ray { f3 o, d; };

struct Hit {
    float t, u, v;
    int32_t slot;     // leaf-order slot of the closest triangle, -1 = miss
    int32_t id;       // original triangle id
};

// -------------------------------------------------------- ray generation ----
// HW2: Camera::get_ray(float,float), GPUandCPU/include/camera.h:49-53.
// HW1: get_pixel_position(int,int) + Ray ctor, HW1/include/camera.h:33-35, ray.h:25 — the
//      float pixel coordinate is truncated to int by overload resolution (quirk Q1).
RT_HD Ray rt_make_ray(const rt_camera& cam, int mode, int x, int y, float jx, float jy) {
    float px = XADD((float)x, jx), py = XADD((float)y, jy);
    f3 c = ld3(cam.center), p00 = ld3(cam.pixel00_loc), du = ld3(cam.pixel_delta_u), dv = ld3(cam.pixel_delta_v);
    Ray r; r.o = c;
    if (mode == RT_MODE_HW2_CPU) {
        // CPUOnly/src/render.cpp:124-135: u = double(i) + du, narrowed by get_pixel_position(double,double)
        // (CPUOnly/include/camera.h:41-43); the Ray constructor normalises (CPUOnly/include/ray.h:13-14).
        px = (float)((double)x + (double)jx); py = (float)((double)y + (double)jy);
        f3 pp = xadd3(xadd3(p00, xmuls(du, px)), xmuls(dv, py));
        r.d = xunit_c(xsub3(pp, c));
    } else if (mode == RT_MODE_HW1) {
        px = (float)(int)px; py = (float)(int)py;
        f3 pp = xadd3(xadd3(p00, xmuls(du, px)), xmuls(dv, py));
        r.d = xunit(xsub3(pp, c));
    } else {
        f3 pp = xadd3(xadd3(p00, xmuls(du, px)), xmuls(dv, py));
        r.d = xunit_cam(xsub3(pp, c));
    }
    return r;
}

// ------------------------------------------------------- Möller–Trumbore ----
// One routine, three contracts (SURVEY §8a "mode parameters"):
//   HW1      HW1/include/ray.h:67-117           |det| < FLT_EPSILON rejects, t >= 0
//   HW2_BVH  GPUandCPU/include/query.h:72-132    |det| < 1e-8 rejects, tmin <= t <= tmax
//   HW2_CPU  CPUOnly/include/ray.h:48-97         as HW1 (1.0f/det)
// Operation order and rounding are the reference's; e1/e2 come pre-rounded from the block.
// Returns true when the triangle test accepts; the caller applies the closest-hit rule.
RT_HD bool rt_moller_trumbore(const Ray& r, f3 v0, f3 e1, f3 e2, float det_eps, float tmin, float tmax,
                              float& t_out, float& u_out, float& v_out) {
    f3 pvec = xcross(r.d, e2);
    float det = xdot(e1, pvec);
    if (fabsf(det) < det_eps) return false;
    float invDet = XRCP(det);
    f3 tvec = xsub3(r.o, v0);
    float u = XMUL(xdot(tvec, pvec), invDet);
    if (u < 0.0f || u > 1.0f) return false;
    f3 qvec = xcross(tvec, e1);
    float v = XMUL(xdot(r.d, qvec), invDet);
    if (v < 0.0f || XADD(u, v) > 1.0f) return false;
    float t = XMUL(xdot(e2, qvec), invDet);
    if (t < tmin || t > tmax) return false;
    t_out = t; u_out = u; v_out = v;
    return true;
}

// The same test with a division-free front end (what the packet kernels run): most warp-level tests of a packet
// accept no lane, and the reference's formulation pays an IEEE divide before its first reject.  Here the three
// barycentric numerators are first judged with an APPROXIMATE reciprocal `ra` of det (rcp.approx on the device,
// relative error <= 2^-22); a lane leaves early only when the exactly-rounded test below could not accept either:
//   ua = fl(U*ra) < -2^-100        =>  U/det < -2^-101                 =>  u = fl(U*fl(1/det)) < 0
//   ua > 1 + 2^-20                 =>  U/det > 1 + 2^-21               =>  u > 1
//   va < -2^-100                   =>  v < 0                           (same argument)
//   fl(ua+va) > 1 + 2^-19          =>  u + v > 1 + 2^-21 (u, v >= -2^-100 here, |u-ua| <= 2^-20 |ua|)  =>  fl(u+v) > 1
//   ta < 0.999996 tmin - 2^-100    =>  t < tmin          ta > 1.000004 tmax + 2^-100  =>  t > tmax
// (the 2^-100 guard keeps products that underflow to -0 — which the reference accepts — on the exact path; NaNs
// fail every comparison and stay on it too).  Survivors run the reference's operations in the reference's order, so
// an accepted hit carries bit-identical (t, u, v).  tests/test_device_code_on_host.py sweeps the reference's
// 57-point barycentric grid and random near-edge probes with adversarially perturbed `ra`.
#define RT_LAZY_TINY 7.8886091e-31f          /* 2^-100 */
RT_HD bool rt_moller_trumbore_lazy_r(const Ray& r, f3 v0, f3 e1, f3 e2, float det_eps, float tmin, float tmax, float ra_scale,
                                     float& t_out, float& u_out, float& v_out) {
    f3 pvec = xcross(r.d, e2);
    float det = xdot(e1, pvec);
    if (fabsf(det) < det_eps) return false;
#if defined(__CUDA_ARCH__)
    float ra;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(det));
    (void)ra_scale;
#else
    const float ra = (1.0f / det) * ra_scale;       // host model: tests pass 1 +- 2^-22 to emulate the worst approximation
#endif
    f3 tvec = xsub3(r.o, v0);
    const float U = xdot(tvec, pvec);
    const float ua = U * ra;
    if (ua < -RT_LAZY_TINY || ua > 1.00000095367431640625f) return false;
    f3 qvec = xcross(tvec, e1);
    const float V = xdot(r.d, qvec);
    const float va = V * ra;
    if (va < -RT_LAZY_TINY || ua + va > 1.0000019073486328125f) return false;
    const float T = xdot(e2, qvec);
    const float ta = T * ra;
    if (ta < 0.999996f * tmin - RT_LAZY_TINY || ta > 1.000004f * tmax + RT_LAZY_TINY) return false;
    const float invDet = XRCP(det);
    const float u = XMUL(U, invDet);
    if (u < 0.0f || u > 1.0f) return false;
    const float v = XMUL(V, invDet);
    if (v < 0.0f || XADD(u, v) > 1.0f) return false;
    const float t = XMUL(T, invDet);
    if (t < tmin || t > tmax) return false;
    t_out = t; u_out = u; v_out = v;
    return true;
}
RT_HD bool rt_moller_trumbore_lazy(const Ray& r, f3 v0, f3 e1, f3 e2, float det_eps, float tmin, float tmax,
                                   float& t_out, float& u_out, float& v_out) {
    return rt_moller_trumbore_lazy_r(r, v0, e1, e2, det_eps, tmin, tmax, 1.0f, t_out, u_out, v_out);
}

RT_HD float rt_det_eps(int mode) { return mode == RT_MODE_HW2_BVH ? 1e-8f : FLT_EPSILON; }
RT_HD float rt_tmin(int mode) { return mode == RT_MODE_HW1 ? 0.0f : 1e-4f; }

// ------------------------------------------------------------- slab test ----
// Conservative fp32 slab test.  The reference tests boxes in fp64 (GPUandCPU/include/bvh.h:81-129);
// ours only has to accept a superset: boxes are padded outward at build time
// (2^-17 of the scene scale, half-extents rounded up) and the far bound is widened by a few ulps
// here, so every triangle the fp32 Möller–Trumbore test would accept is reached.  A zero
// direction component gives +-inf / NaN plane distances; fminf/fmaxf drop NaNs, which makes that
// axis accept — a superset of the reference's origin-in-slab branch (bvh.h:90-91).
struct RayInv { f3 o, inv, ainv; };
RT_HD RayInv rt_ray_inv(const Ray& r) {
    RayInv k; k.o = r.o;
    k.inv = mk3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    k.ainv = mk3(fabsf(k.inv.x), fabsf(k.inv.y), fabsf(k.inv.z));
    return k;
}
RT_HD bool rt_slab_combine(float mx, float my, float mz, const f3& ainv, float hx, float hy, float hz,
                           float tmin, float tmax, float& tnear) {
    const float nx = fmaf(-hx, ainv.x, mx), fx = fmaf(hx, ainv.x, mx);
    const float ny = fmaf(-hy, ainv.y, my), fy = fmaf(hy, ainv.y, my);
    const float nz = fmaf(-hz, ainv.z, mz), fz = fmaf(hz, ainv.z, mz);
    const float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    const float tf = fminf(fminf(fx, fy), fminf(fz, tmax));
    tnear = tn;
    return tn <= tf * 1.0000005f + 1e-30f;
}
// exact-difference form: m = (c - o) * inv (no cancellation error however far the origin is)
RT_HD bool rt_slab(const RayInv& k, float cx, float cy, float cz, float hx, float hy, float hz,
                   float tmin, float tmax, float& tnear) {
    return rt_slab_combine((cx - k.o.x) * k.inv.x, (cy - k.o.y) * k.inv.y, (cz - k.o.z) * k.inv.z, k.ainv, hx, hy, hz, tmin, tmax, tnear);
}
// Fused form: m = fma(c, inv, -(o*inv)).  Its extra rounding error is a few ulps of |o| + |c| in
// space, which the build-time padding absorbs as long as the ray origin is within ~8 scene
// extents of the scene (the host checks this per frame and otherwise selects rt_slab; shadow-ray
// origins lie on the surface and always qualify).
struct RayFma { f3 inv, ainv, oi; };
RT_HD RayFma rt_ray_fma(const Ray& r) {
    RayFma k;
    k.inv = mk3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    k.ainv = mk3(fabsf(k.inv.x), fabsf(k.inv.y), fabsf(k.inv.z));
    k.oi = mk3(r.o.x * k.inv.x, r.o.y * k.inv.y, r.o.z * k.inv.z);
    return k;
}
RT_HD bool rt_slab_fma(const RayFma& k, float cx, float cy, float cz, float hx, float hy, float hz,
                       float tmin, float tmax, float& tnear) {
    return rt_slab_combine(fmaf(cx, k.inv.x, -k.oi.x), fmaf(cy, k.inv.y, -k.oi.y), fmaf(cz, k.inv.z, -k.oi.z), k.ainv, hx, hy, hz, tmin, tmax, tnear);
}
// Same without the final widening of the far bound: with the origin within 8 scene extents the
// accumulated rounding error of the nine FMAs is below scale*2^-18 in space, half of the 2^-17
// padding every box carries, so the plain comparison is already conservative.
RT_HD bool rt_slab_fma_tight(const RayFma& k, float cx, float cy, float cz, float hx, float hy, float hz, float tmin, float tmax) {
    const float mx = fmaf(cx, k.inv.x, -k.oi.x), my = fmaf(cy, k.inv.y, -k.oi.y), mz = fmaf(cz, k.inv.z, -k.oi.z);
    const float tn = fmaxf(fmaxf(fmaf(-hx, k.ainv.x, mx), fmaf(-hy, k.ainv.y, my)), fmaxf(fmaf(-hz, k.ainv.z, mz), tmin));
    const float tf = fminf(fminf(fmaf(hx, k.ainv.x, mx), fmaf(hy, k.ainv.y, my)), fminf(fmaf(hz, k.ainv.z, mz), tmax));
    return tn <= tf;
}

// ---------------------------------------------------------------- shading ----
// shade(), HW1/include/raytracer.h:21-48, constant METAL material of HW1/include/ray.h:111-114.
RT_HD f3 rt_shade_hw1_miss(const Ray& r) {
    f3 ud = xunit(r.d);
    float t = XMUL(0.5f, XADD(ud.z, 1.0f));
    return xadd3(xmuls(mk3(1.0f, 1.0f, 1.0f), XSUB(1.0f, t)), xmuls(mk3(0.5f, 0.7f, 1.0f), t));
}
RT_HD f3 rt_shade_hw1_hit(const Ray& r, f3 p, f3 normal, const rt_light& light) {
    const f3 albedo = mk3(0.8f, 0.2f, 0.2f);
    f3 lpos = ld3(light.position), lcol = ld3(light.color);
    f3 ambient = xmuls(albedo, 0.1f);
    f3 lightDir = xunit(xsub3(lpos, p));
    float diff = fmaxf(xdot(normal, lightDir), 0.0f);
    f3 diffuse = xmuls(xmulv(albedo, lcol), diff);
    f3 viewDir = xunit(xsub3(r.o, p));
    f3 halfDir = xunit(xadd3(lightDir, viewDir));
    float spec = XPOW(fmaxf(xdot(normal, halfDir), 0.0f), 64.0f);
    f3 specular = xmuls(lcol, spec);
    f3 c = xadd3(xadd3(ambient, diffuse), specular);
    if (c.x > 1.0f) c.x = 1.0f;
    if (c.y > 1.0f) c.y = 1.0f;
    if (c.z > 1.0f) c.z = 1.0f;
    return c;
}

// Hit attributes of intersectTriangle, GPUandCPU/include/query.h:110-129.
RT_HD void rt_hit_frame_hw2(const Ray& r, f3 e1, f3 e2, f3 n0, f3 n1, f3 n2, float t, float u, float v,
                            f3& p, f3& normal) {
    p = xadd3(r.o, xmuls(r.d, t));
    f3 geomN = xunit(xcross(e1, e2));
    bool front = xdot(r.d, geomN) < 0.0f;
    if (!front) geomN = xneg3(geomN);
    f3 sn = xadd3(xadd3(xmuls(n0, XSUB(XSUB(1.0f, u), v)), xmuls(n1, u)), xmuls(n2, v));
    if (xdot(sn, sn) < 1e-12f) {
        sn = geomN;
    } else {
        sn = xunit(sn);
        if (xdot(sn, geomN) < 0.0f) sn = xneg3(sn);
    }
    normal = sn;
}

// ray_intersection's hit attributes in the CPUOnly renderer (CPUOnly/include/ray.h:74-94): face normal from the
// winding, front/back by the ray, interpolated vertex normals normalised and negated on back faces.
// has_normals == false: the triangle carries its face normal in all three slots (CPUOnly/src/render.cpp:88-96).
RT_HD void rt_hit_frame_cpu(const Ray& r, f3 e1, f3 e2, f3 n0, f3 n1, f3 n2, bool has_normals, float t, float u, float v,
                            f3& p, f3& normal) {
    p = xadd3(r.o, xmuls(r.d, t));
    const f3 faceN = xunit_c(xcross(e1, e2));
    const bool front = xdot(r.d, faceN) < 0.0f;
    if (!has_normals) { n0 = faceN; n1 = faceN; n2 = faceN; }
    f3 sn = xadd3(xadd3(xmuls(n0, XSUB(XSUB(1.0f, u), v)), xmuls(n1, u)), xmuls(n2, v));
    sn = xunit_c(sn);
    if (!front) sn = xneg3(sn);
    normal = sn;
}

// EvaluateBRDF of the CPUOnly renderer (CPUOnly/include/brdf.h:12-37): fs = specularColor * (ks * specLobe).
RT_HD f3 rt_brdf_cpu(const rt_material& m, f3 N, f3 V, f3 L) {
    float NdotL = fmaxf(xdot(N, L), 0.0f);
    float NdotV = fmaxf(xdot(N, V), 0.0f);
    if (NdotL <= 0.f || NdotV <= 0.f) return mk3(0.f, 0.f, 0.f);
    const float invPi = 0.31830988618f;
    f3 fd = xmuls(ld3(m.albedo), XMUL(m.kd, invPi));
    f3 H = xunit_c(xadd3(L, V));
    float NdotH = fmaxf(xdot(N, H), 0.0f);
    const float inv2Pi = 0.15915494309f;
    float specNorm = XMUL(XADD(m.shininess, 2.0f), inv2Pi);
    float specLobe = XMUL(specNorm, XPOW(NdotH, m.shininess));
    f3 fs = xmuls(ld3(m.specular_color), XMUL(m.ks, specLobe));
    return xadd3(fd, fs);
}

// EvaluateBRDF, GPUandCPU/include/brdf.h:12-39.
RT_HD f3 rt_brdf_hw2(const rt_material& m, f3 N, f3 V, f3 L) {
    float NdotL = fmaxf(xdot(N, L), 0.0f);
    float NdotV = fmaxf(xdot(N, V), 0.0f);
    if (NdotL <= 0.f || NdotV <= 0.f) return mk3(0.f, 0.f, 0.f);
    const float invPi = 0.31830988618f;
    f3 fd = xmuls(ld3(m.albedo), XMUL(m.kd, invPi));
    f3 H = xunit(xadd3(L, V));
    float NdotH = fmaxf(xdot(N, H), 0.0f);
    const float inv2Pi = 0.15915494309f;
    float specNorm = XMUL(XADD(m.shininess, 2.0f), inv2Pi);
    float specLobe = XMUL(specNorm, XPOW(NdotH, m.shininess));
    f3 fs = xmuls(xmuls(ld3(m.specular_color), m.ks), specLobe);
    return xadd3(fd, fs);
}

RT_HD f3 rt_clamp01(f3 c) {   // clamp(), GPUandCPU/include/shader.h:24-32
    if (c.x > 1.0f) c.x = 1.0f; if (c.y > 1.0f) c.y = 1.0f; if (c.z > 1.0f) c.z = 1.0f;
    if (c.x < 0.0f) c.x = 0.0f; if (c.y < 0.0f) c.y = 0.0f; if (c.z < 0.0f) c.z = 0.0f;
    return c;
}

RT_HD rt_material rt_default_material() {   // Material(), GPUandCPU/include/material.h:6-20
    rt_material m;
    m.albedo[0] = m.albedo[1] = m.albedo[2] = 0.8f; m.kd = 1.0f;
    m.specular_color[0] = m.specular_color[1] = m.specular_color[2] = 0.04f; m.ks = 0.0f;
    m.shininess = 32.0f; m.kr = 0.0f; m.emission[0] = m.emission[1] = m.emission[2] = 0.0f;
    return m;
}

// -------------------------------------------------------------- quantiser ----
RT_UNIT_FN uint8_t rt_quantise(float c, int q) {      // (one copy in device code, see rt_math.h RT_UNIT_FN)
    if (q == RT_QUANT_HW1_TRUNC) return (uint8_t)(XMUL(255.99f, c));                 // HW1/src/render.cpp:121-123
    if (q == RT_QUANT_HW2_TRUNC) return (uint8_t)(XMUL(255.0f, (c < 1.0f ? c : 1.0f))); // GPUandCPU/src/main.cu:428-430
    if (q == RT_QUANT_CPU_TRUNC) { float x = c > 1.0f ? 1.0f : c; if (x < 0.0f) x = 0.0f; return (uint8_t)(XMUL(255.99f, x)); }   // CPUOnly/src/render.cpp:157-163
    double x = (double)c;                                                            // ppm_p6.cpp:137-155
    if (q == RT_QUANT_PPM_GAMMA2) { if (x < 0.0) x = 0.0; x = sqrt(x); }
    if (x < 0.0) x = 0.0;
    if (x > 1.0) x = 1.0;
    long long r = llround(x * 255.0);   // std::lround: half away from zero
    if (r < 0) r = 0;
    if (r > 255) r = 255;
    return (uint8_t)r;
}
