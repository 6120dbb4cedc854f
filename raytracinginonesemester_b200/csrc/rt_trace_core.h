// rt_trace_core.h — per-ray traversal and shading glue, host/device compilable.
// The kernels in rt_trace.cu call these with a shared-memory stack column per thread
// (stride RT_BLOCK_THREADS); the host-side emulation in tests/ calls them with stride 1.
#pragma once

#include "rt_params.h"

struct Tri { f3 v0, e1, e2; int id; };

RT_HD Tri rt_load_tri(const TriBlock* __restrict__ geom, uint32_t slot) {
    const float4* p = reinterpret_cast<const float4*>(geom + slot);
    const float4 a = RT_LDG(p), b = RT_LDG(p + 1), c = RT_LDG(p + 2);
    Tri t;
    t.v0 = mk3(a.x, a.y, a.z); t.id = RT_F2I(a.w);
    t.e1 = mk3(b.x, b.y, b.z);
    t.e2 = mk3(c.x, c.y, c.z);
    return t;
}

// Canonical closest-hit rule: min t, then min original triangle id.
// rt_moller_trumbore is called with tmax = best.t, so an accepted t is <= best.t.
RT_HD void rt_consider(const Ray& ray, const Tri& tr, uint32_t slot, float det_eps, float tmin, Hit& best) {
    float t, u, v;
    if (rt_moller_trumbore(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, best.t, t, u, v)) {
        if (t < best.t || tr.id < best.id) { best.t = t; best.u = u; best.v = v; best.slot = (int)slot; best.id = tr.id; }
    }
}

RT_HD void rt_hit_reset(Hit& h) { h.t = FLT_MAX; h.u = 0.f; h.v = 0.f; h.slot = -1; h.id = 0x7fffffff; }

struct NodeQ { float4 q0, q1, q2; int4 q3; };
RT_HD NodeQ rt_load_node(const BvhNode* __restrict__ nodes, int idx) {
    const float4* n = reinterpret_cast<const float4*>(nodes + idx);
    NodeQ q;
#ifdef __CUDA_ARCH__
    // sm_100: the 64-byte node line as two 256-bit read-only loads (LDG.E.ENL2.256.CONSTANT) instead of four
    // 128-bit ones: half the load instructions and half the address set-up per node visit.
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(q.q0.x), "=f"(q.q0.y), "=f"(q.q0.z), "=f"(q.q0.w), "=f"(q.q1.x), "=f"(q.q1.y), "=f"(q.q1.z), "=f"(q.q1.w) : "l"(n));
    float i0, i1, i2, i3;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(q.q2.x), "=f"(q.q2.y), "=f"(q.q2.z), "=f"(q.q2.w), "=f"(i0), "=f"(i1), "=f"(i2), "=f"(i3) : "l"(n + 2));
    q.q3 = make_int4(__float_as_int(i0), __float_as_int(i1), __float_as_int(i2), __float_as_int(i3));
#else
    q.q0 = RT_LDG(n); q.q1 = RT_LDG(n + 1); q.q2 = RT_LDG(n + 2);
    q.q3 = RT_LDG(reinterpret_cast<const int4*>(n + 3));
#endif
    return q;
}

struct TraceStats { uint32_t nodes, tris, max_sp, wnodes, wtris; };

// Stack-based BVH2 closest hit (replaces SearchBVH, GPUandCPU/include/query.h:224-311):
// one 64-byte node visit tests both children, descends into the nearer one and pushes the
// farther; leaves are contiguous runs of triangle blocks.  STRIDE is the distance between
// consecutive stack entries of one thread (shared-memory column layout on the device).
template <int MODE, int STRIDE, bool STATS>
RT_HD void rt_trace_closest(const FrameParams& P, const Ray& ray, uint32_t* stk, Hit& best, TraceStats* st) {
    const float det_eps = rt_det_eps(MODE), tmin = rt_tmin(MODE);
    rt_hit_reset(best);
    const RayInv k = rt_ray_inv(ray);
    int sp = 0, cur = 0;
    bool overflow = false;
    while (true) {
        if (cur >= 0) {
            const NodeQ q = rt_load_node(P.nodes, cur);
            if (STATS) st->nodes++;
            float tn0, tn1;
            const bool h0 = rt_slab(k, q.q0.x, q.q0.y, q.q0.z, q.q0.w, q.q1.x, q.q1.y, tmin, best.t, tn0);
            const bool h1 = rt_slab(k, q.q1.z, q.q1.w, q.q2.x, q.q2.y, q.q2.z, q.q2.w, tmin, best.t, tn1);
            if (h0 & h1) {
                int nearc = q.q3.x, farc = q.q3.y;
                if (tn1 < tn0) { nearc = q.q3.y; farc = q.q3.x; }
                if (sp < RT_STACK_DEPTH) { stk[sp * STRIDE] = (uint32_t)farc; ++sp; } else overflow = true;
                if (STATS && (uint32_t)sp > st->max_sp) st->max_sp = (uint32_t)sp;
                cur = nearc;
                continue;
            }
            if (h0) { cur = q.q3.x; continue; }
            if (h1) { cur = q.q3.y; continue; }
        } else {
            const uint32_t first = rt_leaf_first(cur), cnt = rt_leaf_count(cur);
            for (uint32_t s = first; s < first + cnt; ++s) {
                const Tri tr = rt_load_tri(P.geom, s);
                if (STATS) st->tris++;
                rt_consider(ray, tr, s, det_eps, tmin, best);
            }
        }
        if (sp == 0) break;
        --sp;
        cur = (int)stk[sp * STRIDE];
    }
    if (overflow) {   // same safety net as the reference (query.h:297-308): finish by brute force
        for (uint32_t s = 0; s < P.num_tris; ++s) {
            const Tri tr = rt_load_tri(P.geom, s);
            rt_consider(ray, tr, s, det_eps, tmin, best);
        }
    }
}

// Any-hit query for IsInShadow (GPUandCPU/include/shader.h:44-62): blocked iff some triangle is
// accepted by intersectTriangle(tmin, FLT_MAX) with t < dist — the same boolean as the
// reference's "closest hit exists and its t < dist".
template <int MODE, int STRIDE, bool STATS>
RT_HD bool rt_trace_any(const FrameParams& P, const Ray& ray, float tmax_excl, uint32_t* stk, TraceStats* st) {
    const float det_eps = rt_det_eps(MODE), tmin = rt_tmin(MODE);
    const RayInv k = rt_ray_inv(ray);
    int sp = 0, cur = 0;
    bool overflow = false;
    while (true) {
        if (cur >= 0) {
            const NodeQ q = rt_load_node(P.nodes, cur);
            if (STATS) st->nodes++;
            float tn0, tn1;
            const bool h0 = rt_slab(k, q.q0.x, q.q0.y, q.q0.z, q.q0.w, q.q1.x, q.q1.y, tmin, tmax_excl, tn0);
            const bool h1 = rt_slab(k, q.q1.z, q.q1.w, q.q2.x, q.q2.y, q.q2.z, q.q2.w, tmin, tmax_excl, tn1);
            if (h0 & h1) {
                int nearc = q.q3.x, farc = q.q3.y;
                if (tn1 < tn0) { nearc = q.q3.y; farc = q.q3.x; }
                if (sp < RT_STACK_DEPTH) { stk[sp * STRIDE] = (uint32_t)farc; ++sp; } else overflow = true;
                cur = nearc;
                continue;
            }
            if (h0) { cur = q.q3.x; continue; }
            if (h1) { cur = q.q3.y; continue; }
        } else {
            const uint32_t first = rt_leaf_first(cur), cnt = rt_leaf_count(cur);
            for (uint32_t s = first; s < first + cnt; ++s) {
                const Tri tr = rt_load_tri(P.geom, s);
                if (STATS) st->tris++;
                float t, u, v;
                if (rt_moller_trumbore(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, FLT_MAX, t, u, v) && t < tmax_excl) return true;
            }
        }
        if (sp == 0) break;
        --sp;
        cur = (int)stk[sp * STRIDE];
    }
    if (overflow) {
        for (uint32_t s = 0; s < P.num_tris; ++s) {
            const Tri tr = rt_load_tri(P.geom, s);
            float t, u, v;
            if (rt_moller_trumbore(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, FLT_MAX, t, u, v) && t < tmax_excl) return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------ shading glue ----
struct Surface {       // hit-point state shared by the per-light steps
    f3 p, normal, N, V, Lo;
    rt_material mat;
};

// Per-vertex normals + object id of a slot.  A mesh without normals has no normals blocks (P.shade == nullptr, uniform over
// the launch): the normals read as zero, which is what the reference's loaders leave there, and the id comes from the
// geometry block.
RT_HD void rt_load_normals(const FrameParams& P, int slot, f3& n0, f3& n1, f3& n2, int& obj) {
    if (P.shade == nullptr) {
        n0 = n1 = n2 = mk3(0.f, 0.f, 0.f);
        obj = RT_F2I(RT_LDG(reinterpret_cast<const float4*>(P.geom + slot) + 1).w);
        return;
    }
    const float4* p = reinterpret_cast<const float4*>(P.shade + slot);
    const float4 a = RT_LDG(p), b = RT_LDG(p + 1), c = RT_LDG(p + 2);
    n0 = mk3(a.x, a.y, a.z); obj = RT_F2I(a.w);
    n1 = mk3(b.x, b.y, b.z);
    n2 = mk3(c.x, c.y, c.z);
}

// ShadeDirect prologue, GPUandCPU/include/shader.h:75-85 (+ assignMaterialToHit, query.h:134-153)
RT_HD void rt_surface_hw2(const FrameParams& P, const Ray& ray, const Hit& h, Surface& s) {
    const Tri tr = rt_load_tri(P.geom, (uint32_t)h.slot);
    f3 n0, n1, n2; int obj;
    rt_load_normals(P, h.slot, n0, n1, n2, obj);
    rt_hit_frame_hw2(ray, tr.e1, tr.e2, n0, n1, n2, h.t, h.u, h.v, s.p, s.normal);
    s.mat = rt_default_material();
    if (P.materials != nullptr && obj >= 0 && obj < P.num_materials) s.mat = P.materials[obj];
    s.N = xunit(s.normal);
    s.V = xunit(xsub3(ray.o, s.p));
    s.Lo = mk3(0.f, 0.f, 0.f);
    s.Lo = xadd3(s.Lo, xmuls(ld3(s.mat.albedo), 0.05f));
    s.Lo = xadd3(s.Lo, ld3(s.mat.emission));
}

// Per-light setup: false when the light contributes nothing (N.L <= 0).  need_shadow is set and
// (sray, dist) describe the IsInShadow query (shader.h:44-58, RT_EPS = 1e-3, shader.h:22).
RT_HD bool rt_light_setup_hw2(const Surface& s, const rt_light& light, f3& L, float& NdotL,
                              bool& need_shadow, Ray& sray, float& dist) {
    const f3 lpos = ld3(light.position);
    L = xunit(xsub3(lpos, s.p));
    NdotL = fmaxf(xdot(s.N, L), 0.0f);
    need_shadow = false;
    if (NdotL <= 0.0f) return false;
    const f3 toL = xsub3(lpos, s.p);
    dist = xlen3(toL);
    if (!(dist <= 0.0f)) {
        need_shadow = true;
        sray.d = xdivs(toL, dist);
        sray.o = xadd3(s.p, xmuls(s.N, 1e-3f));
    }
    return true;
}
RT_HD f3 rt_light_direct_hw2(const Surface& s, const rt_light& light, f3 L, float NdotL) {
    const f3 f = rt_brdf_hw2(s.mat, s.normal, s.V, L);
    const f3 radiance = xmuls(ld3(light.color), (float)light.intensity);
    return xmuls(xmulv(radiance, f), NdotL);
}
RT_HD void rt_light_finish_hw2(Surface& s, const rt_light& light, f3 L, float NdotL) {
    s.Lo = xadd3(s.Lo, rt_light_direct_hw2(s, light, L, NdotL));
}
// TraceRayIterative epilogue at depth 1 (query.h:186-191, 219)
RT_HD f3 rt_radiance_hw2(f3 direct) {
    f3 radiance = mk3(0.f, 0.f, 0.f);
    const f3 throughput = mk3(1.f, 1.f, 1.f);
    radiance = xadd3(radiance, xmulv(throughput, direct));
    return rt_clamp01(radiance);
}

RT_HD f3 rt_shade_hw1(const FrameParams& P, const Ray& ray, const Hit& h) {
    if (h.slot < 0) return rt_shade_hw1_miss(ray);
    f3 n0, n1, n2; int obj;
    rt_load_normals(P, h.slot, n0, n1, n2, obj);
    const f3 p = xadd3(ray.o, xmuls(ray.d, h.t));
    const f3 normal = xadd3(xadd3(xmuls(n0, XSUB(XSUB(1.0f, h.u), h.v)), xmuls(n1, h.u)), xmuls(n2, h.v));
    return rt_shade_hw1_hit(ray, p, normal, P.lights[0]);
}

// ---------------------------------------------------------------- bounce loop ----
// rng_next / make_rng_seed / random_unit_vector / random_on_hemisphere, GPUandCPU/include/query.h:32-70.
RT_HD float rt_rng_next(uint32_t& state) {
    state = state * 1664525u + 1013904223u;
    uint32_t h = state;
    h = (h ^ 61u) ^ (h >> 16u);
    h *= 9u;
    h ^= h >> 4u;
    h *= 0x27d4eb2du;
    h ^= h >> 15u;
    return XDIV(RT_U2F(h), 4294967296.0f);          // float(h) / float(0xFFFFFFFFu); the divisor rounds to 2^32
}
RT_HD uint32_t rt_rng_seed(int x, int y, int sample) {
    return (uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)sample * 83492791u;
}
RT_HD f3 rt_random_on_hemisphere(f3 normal, uint32_t& state) {
    f3 u;
    for (;;) {
        const float x = XSUB(XMUL(2.0f, rt_rng_next(state)), 1.0f);
        const float y = XSUB(XMUL(2.0f, rt_rng_next(state)), 1.0f);
        const float z = XSUB(XMUL(2.0f, rt_rng_next(state)), 1.0f);
        const float lensq = XADD(XADD(XMUL(x, x), XMUL(y, y)), XMUL(z, z));
        if (lensq > 1e-10f && lensq <= 1.0f) {
            const float inv = XDIV(1.0f, XSQRT(lensq));
            u = mk3(XMUL(x, inv), XMUL(y, inv), XMUL(z, inv));
            break;
        }
    }
    if (xdot(u, normal) > 0.0f) return u;
    return xneg3(u);
}
// The ray that leaves a hit, query.h:193-216.  Returns false when the path ends (no reflecting material or
// throughput below 1e-4 in every channel).
RT_HD bool rt_bounce_hw2(const FrameParams& P, const Surface& sf, Ray& ray, f3& throughput, uint32_t& rng) {
    const float kd = sf.mat.kd, kr = sf.mat.kr, total = XADD(kd, kr);
    if (total <= 0.0f) return false;
    const f3 N = sf.N;                                   // normalize(hitRecord.normal): the same three divides as unit_vector
    const float xi = rt_rng_next(rng);
    const f3 o = xadd3(sf.p, xmuls(N, 1e-3f));           // RT_EPS, shader.h:22
    if (P.diffuse_bounce && xi < XDIV(kd, total)) {
        const f3 d = rt_random_on_hemisphere(N, rng);
        const float NdotL = fmaxf(xdot(N, d), 0.0f);
        throughput = xmulv(throughput, xmuls(ld3(sf.mat.albedo), XMUL(2.0f, NdotL)));
        ray.o = o; ray.d = d;
    } else {
        const f3 I = xunit(ray.d);
        const f3 refl = xsub3(I, xmuls(N, XMUL(2.0f, xdot(I, N))));   // reflect_dir, shader.h:38-42
        throughput = xmulv(throughput, xmuls(ld3(sf.mat.specular_color), kr));
        ray.o = o; ray.d = refl;
    }
    return !(throughput.x < 1e-4f && throughput.y < 1e-4f && throughput.z < 1e-4f);
}

// ------------------------------------------------------ CPUOnly renderer (N1) ----
// TraceRay of HW2/HW2/CPUOnly/include/raytracer.h:215-260 with diffuse_bounce == false and point lights: direct light
// (ShadeDirect :171-211, one shadow ray per lit hit, ShadowVisibility :121-168 with S == 1) plus perfect-mirror
// recursion.  The recursion `Lo + kr * (tint * TraceRay(...))` is evaluated innermost-first like the reference: the
// walk down records (Lo, kr, tint) per level, the fold runs back up.  Sky gradient on a miss (:224-230).
// random_in_unit_disk, CPUOnly/include/raytracer.h:76-85, drawing from the hash RNG (see rt_frame.rng_seed)
RT_HD void rt_random_in_unit_disk(uint32_t& state, float& dx, float& dy) {
    for (;;) {
        const float x = XSUB(XMUL(2.0f, rt_rng_next(state)), 1.0f);
        const float y = XSUB(XMUL(2.0f, rt_rng_next(state)), 1.0f);
        const float r2 = XADD(XMUL(x, x), XMUL(y, y));
        if (r2 > 1e-10f && r2 <= 1.0f) { dx = x; dy = y; return; }
    }
}

template <int STRIDE, bool STATS>
RT_HD f3 rt_sample_cpuonly(const FrameParams& P, const Ray& primary, uint32_t rng, uint32_t* stk, Hit& first,
                           unsigned& nprim, unsigned& nshadow, TraceStats* st) {
    const float EPS = 1e-4f;                                           // RT_EPS, raytracer.h:49
    f3 lvl_Lo[RT_CPU_MAX_DEPTH], lvl_tint[RT_CPU_MAX_DEPTH];
    float lvl_kr[RT_CPU_MAX_DEPTH];
    int depth = P.max_depth > 1 ? P.max_depth : 1;                     // std::max(1, max_bounces), render.cpp:113
    if (depth > RT_CPU_MAX_DEPTH) depth = RT_CPU_MAX_DEPTH;
    Ray ray = primary;
    rt_hit_reset(first);
    int n = 0;
    f3 tail = mk3(0.f, 0.f, 0.f);                                      // TraceRay(depth 0) = black
    for (;;) {
        Hit h;
        rt_trace_closest<RT_MODE_HW2_CPU, STRIDE, STATS>(P, ray, stk, h, st);
        ++nprim;
        if (n == 0) first = h;
        if (h.slot < 0) {
            const f3 ud = xunit_c(ray.d);
            const float t = XMUL(0.5f, XADD(ud.z, 1.0f));
            tail = xadd3(xmuls(mk3(1.0f, 1.0f, 1.0f), XSUB(1.0f, t)), xmuls(mk3(0.5f, 0.7f, 1.0f), t));
            break;
        }
        const Tri tr = rt_load_tri(P.geom, (uint32_t)h.slot);
        f3 n0, n1, n2, p, normal; int obj;
        rt_load_normals(P, h.slot, n0, n1, n2, obj);
        rt_hit_frame_cpu(ray, tr.e1, tr.e2, n0, n1, n2, P.has_normals != 0, h.t, h.u, h.v, p, normal);
        rt_material mat = rt_default_material();
        if (P.materials != nullptr && obj >= 0 && obj < P.num_materials) mat = P.materials[obj];
        const f3 N = xunit_c(normal);
        const f3 V = xunit_c(xsub3(ray.o, p));
        f3 Lo = mk3(0.f, 0.f, 0.f);
        Lo = xadd3(Lo, xmuls(ld3(mat.albedo), 0.05f));
        Lo = xadd3(Lo, ld3(mat.emission));
        for (int l = 0; l < P.num_lights; ++l) {
            const rt_light light = P.lights[l];
            const f3 toL = xsub3(ld3(light.position), p);
            const float dist = xlen3(toL);
            if (dist <= 0.0f) continue;
            const f3 L = xdivs(toL, dist);
            const float NdotL = fmaxf(xdot(N, L), 0.0f);
            if (NdotL <= 0.0f) continue;
            float vis = 1.0f;
            if (P.shadows) {
                // ShadowVisibility, raytracer.h:121-168: S samples of a disk of `radius` at the light, facing the shaded point
                // (radius 0: the light itself, once).  distC == dist > 0 here.
                const float radius = P.light_radius ? P.light_radius[l] : 0.0f;
                int S = (radius > 0.0f && P.light_samples) ? P.light_samples[l] : 1;
                if (S < 1) S = 1;
                const f3 lpos = ld3(light.position);
                f3 T = mk3(0.f, 0.f, 0.f), B = T;
                if (radius > 0.0f) {                                   // make_basis(W, T, B), :88-93
                    const f3 Wd = xdivs(xsub3(p, lpos), dist);
                    const f3 a = fabsf(Wd.x) > 0.9f ? mk3(0.f, 1.f, 0.f) : mk3(1.f, 0.f, 0.f);
                    T = xunit_c(xcross(a, Wd));
                    B = xcross(Wd, T);
                }
                float unoccluded = 0.0f;
                for (int i = 0; i < S; ++i) {
                    f3 lightPos = lpos;
                    if (radius > 0.0f) {
                        float dx, dy;
                        rt_random_in_unit_disk(rng, dx, dy);
                        lightPos = xadd3(xadd3(lpos, xmuls(T, XMUL(dx, radius))), xmuls(B, XMUL(dy, radius)));
                    }
                    const f3 toS = xsub3(lightPos, p);
                    const float distToL = xlen3(toS);
                    if (distToL <= 0.0f) { unoccluded = XADD(unoccluded, 1.0f); continue; }
                    Ray sray;
                    sray.o = xadd3(p, xmuls(N, EPS));
                    sray.d = xunit_c(xdivs(toS, distToL));             // Ldir, and the Ray constructor normalises again
                    // blocked iff some triangle has 1e-4 <= t and double(t) < double(dist) - double(1e-4f): as a float
                    // threshold, t < the smallest float >= that double
                    const float thr = RT_D2F_UP((double)distToL - (double)EPS);
                    ++nshadow;
                    if (!rt_trace_any<RT_MODE_HW2_CPU, STRIDE, STATS>(P, sray, thr, stk, st)) unoccluded = XADD(unoccluded, 1.0f);
                }
                vis = XDIV(unoccluded, (float)S);
            }
            if (vis <= 0.0f) continue;
            const f3 f = rt_brdf_cpu(mat, N, V, L);
            const f3 radiance = xmuls(ld3(light.color), light.intensity_f);
            Lo = xadd3(Lo, xmuls(xmulv(radiance, f), XMUL(NdotL, vis)));
        }
        lvl_Lo[n] = Lo; lvl_kr[n] = 0.0f; lvl_tint[n] = mk3(0.f, 0.f, 0.f);
        const bool bounce = XADD(mat.kd, mat.kr) > 0.0f && mat.kr > 0.0f;
        if (!bounce) { ++n; tail = mk3(0.f, 0.f, 0.f); lvl_kr[n - 1] = -1.0f; break; }   // kr slot < 0: level returns Lo as is
        lvl_kr[n] = mat.kr; lvl_tint[n] = ld3(mat.specular_color);
        ++n;
        if (n >= depth) { tail = mk3(0.f, 0.f, 0.f); break; }          // the next call has depth 0
        const f3 I = xunit_c(ray.d);
        const f3 refl = xsub3(I, xmuls(N, XMUL(2.0f, xdot(I, N))));    // reflect_dir, raytracer.h:70-74
        ray.o = xadd3(p, xmuls(N, EPS));
        ray.d = xunit_c(refl);
    }
    f3 v = tail;
    for (int k = n - 1; k >= 0; --k) {
        if (lvl_kr[k] < 0.0f) v = lvl_Lo[k];
        else v = xadd3(lvl_Lo[k], xmuls(xmulv(lvl_tint[k], v), lvl_kr[k]));
    }
    return v;
}

// The loop of TraceRayIterative (query.h:156-216) from the point where segment `depth` has found `cur`: miss colour or
// surface + direct light (+ shadow rays), then the bounce and the next closest hit.  rt_sample_bvh enters it at depth 0; the
// packet kernels trace depth 0 as packets and enter it at depth 1 with the first bounce ray's hit.
template <int MODE, int STRIDE, bool STATS>
RT_HD void rt_path_continue(const FrameParams& P, Ray& ray, Hit cur, int depth, f3& radiance, f3& throughput, uint32_t& rng,
                            uint32_t* stk, unsigned& nprim, unsigned& nshadow, TraceStats* st) {
    for (;;) {
        if (cur.slot < 0) { radiance = xadd3(radiance, xmulv(throughput, ld3(P.miss))); break; }
        Surface sf;
        rt_surface_hw2(P, ray, cur, sf);
        for (int l = 0; l < P.num_lights; ++l) {
            const rt_light light = P.lights[l];
            f3 L; float NdotL, dist; bool need; Ray sray;
            if (!rt_light_setup_hw2(sf, light, L, NdotL, need, sray, dist)) continue;
            if (need && P.shadows) {
                ++nshadow;
                if (rt_trace_any<MODE, STRIDE, STATS>(P, sray, dist, stk, st)) continue;
            }
            rt_light_finish_hw2(sf, light, L, NdotL);
        }
        radiance = xadd3(radiance, xmulv(throughput, sf.Lo));
        if (++depth >= P.max_depth) break;
        if (!rt_bounce_hw2(P, sf, ray, throughput, rng)) break;
        rt_trace_closest<MODE, STRIDE, STATS>(P, ray, stk, cur, st);
        ++nprim;
    }
}

// One sample of one pixel through the BVH path: TraceRayIterative (query.h:156-220) — closest hit, shading,
// shadow rays, then mirror / diffuse bounces up to P.max_depth.  h = the depth-0 hit.
template <int MODE, int STRIDE, bool STATS>
RT_HD f3 rt_sample_bvh(const FrameParams& P, int x, int y, int s, uint32_t* stk, Hit& h,
                       unsigned& nprim, unsigned& nshadow, TraceStats* st) {
    const float jx = P.jitter ? RT_LDG(P.jitter + 2 * s) : 0.0f;
    const float jy = P.jitter ? RT_LDG(P.jitter + 2 * s + 1) : 0.0f;
    Ray ray = rt_make_ray(P.cam, MODE, x, y, jx, jy);
    if (MODE == RT_MODE_HW2_BVH && P.max_depth <= 0) {   // TraceRayIterative: maxDepth <= 0 -> black
        rt_hit_reset(h);
        return mk3(0.f, 0.f, 0.f);
    }
    if (MODE == RT_MODE_HW2_CPU) return rt_sample_cpuonly<STRIDE, STATS>(P, ray, rt_rng_seed(x, y, s) ^ P.rng_seed, stk, h, nprim, nshadow, st);
    rt_trace_closest<MODE, STRIDE, STATS>(P, ray, stk, h, st);
    ++nprim;
    if (MODE == RT_MODE_HW1) return rt_shade_hw1(P, ray, h);
    f3 radiance = mk3(0.f, 0.f, 0.f), throughput = mk3(1.f, 1.f, 1.f);
    uint32_t rng = rt_rng_seed(x, y, s);
    rt_path_continue<MODE, STRIDE, STATS>(P, ray, h, 0, radiance, throughput, rng, stk, nprim, nshadow, st);
    return rt_clamp01(radiance);
}

// Resolve: col / float(spp) (query.cu:163, render.cpp:110) + requested planes.
RT_HD void rt_write_pixel(const FrameParams& P, size_t out, f3 accum, const Hit& first) {
    const f3 fin = xdivs(accum, (float)P.spp);
    if (P.rgb) { P.rgb[3 * out] = fin.x; P.rgb[3 * out + 1] = fin.y; P.rgb[3 * out + 2] = fin.z; }
    if (P.rgb8) {
        P.rgb8[3 * out] = rt_quantise(fin.x, P.quantiser);
        P.rgb8[3 * out + 1] = rt_quantise(fin.y, P.quantiser);
        P.rgb8[3 * out + 2] = rt_quantise(fin.z, P.quantiser);
    }
    if (P.tri_id) P.tri_id[out] = first.slot >= 0 ? first.id : -1;
    if (P.t) P.t[out] = first.slot >= 0 ? first.t : -1.0f;
}
