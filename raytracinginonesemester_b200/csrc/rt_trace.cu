// rt_trace.cu — ray generation + closest hit + shading + resolve, one fused kernel per frame.
//
// Replaces renderBatchCUDA / normalizeCUDA / render() of the reference
// (HW2/HW2/GPUandCPU/include/query.cu:12-167) and the HW1 pixel loop (HW1/src/render.cpp:72-124).
// Not a port: the reference walks a 16-byte-node + separate-AABB LBVH with a 512-entry local
// stack and fp64 slabs per thread.  Here a thread block owns a 16x8 pixel tile and a warp is a
// 32-ray packet (an 8x4 patch); inner nodes are culled against the packet's bounding frustum
// with one LANE per BOX of an 8-wide view of the tree (frustum_trace, the default), leaves are
// culled per ray and their triangles — 48-byte pre-differenced blocks in leaf order — tested by
// every lane; shadow rays are any-hit queries in the same kernel.  A per-lane packet traversal
// (packet_trace) and a per-ray kernel with a shared-memory stack column (k_render_bvh) are kept
// as cross-checks and for incoherent rays.  The Möller–Trumbore and shading arithmetic is
// exactly rounded (rt_math.h) so hit ids, t and colours equal the reference CPU build.
#include "rt_kernels.h"
#include "rt_trace_core.h"

#include <cuda_runtime.h>

namespace {

__device__ __forceinline__ Pixel map_pixel(const FrameParams& P) { return rt_map_pixel(P, (int)blockIdx.x + P.tile_offset, (int)threadIdx.x); }

__device__ __forceinline__ void flush_counters(const FrameParams& P, unsigned nprim, unsigned nshadow) {
    for (int o = 16; o > 0; o >>= 1) {
        nprim += __shfl_xor_sync(0xffffffffu, nprim, o);
        nshadow += __shfl_xor_sync(0xffffffffu, nshadow, o);
    }
    if ((threadIdx.x & 31) == 0 && P.counters) {
        if (nprim) atomicAdd(&P.counters[0], (unsigned long long)nprim);
        if (nshadow) atomicAdd(&P.counters[1], (unsigned long long)nshadow);
    }
}

// ------------------------------------------------------------------- BVH kernel ----
template <int MODE, bool STATS>
__global__ void __launch_bounds__(RT_BLOCK_THREADS)
k_render_bvh(const __grid_constant__ FrameParams P) {
    // Per-thread traversal stack as a shared-memory column: entry i of thread t at [i*128 + t],
    // so a warp's pushes/pops hit 32 distinct banks.
    __shared__ uint32_t s_stack[RT_STACK_DEPTH * RT_BLOCK_THREADS];
    uint32_t* stk = s_stack + threadIdx.x;
    const Pixel px = map_pixel(P);
    unsigned nprim = 0, nshadow = 0;
    TraceStats st{0, 0, 0, 0, 0};
    if (px.inside) {
        f3 accum = mk3(0.f, 0.f, 0.f);
        Hit first; rt_hit_reset(first);
        for (int s = 0; s < P.spp; ++s) {
            Hit h;
            const f3 color = rt_sample_bvh<MODE, RT_BLOCK_THREADS, STATS>(P, px.x, px.y, s, stk, h, nprim, nshadow, &st);
            if (s == 0) first = h;
            accum = xadd3(accum, color);
        }
        rt_write_pixel(P, px.out, accum, first);
    }
    flush_counters(P, nprim, nshadow);
    if (STATS) {   // RT_VARIANT_STATS: node visits / triangle tests for the roofline's bytes-per-ray figure
        unsigned nn = st.nodes, nt = st.tris;
        for (int o = 16; o > 0; o >>= 1) { nn += __shfl_xor_sync(0xffffffffu, nn, o); nt += __shfl_xor_sync(0xffffffffu, nt, o); }
        if ((threadIdx.x & 31) == 0 && P.counters) {
            atomicAdd(&P.counters[2], (unsigned long long)nn);
            atomicAdd(&P.counters[3], (unsigned long long)nt);
            atomicAdd(&P.counters[4], (unsigned long long)nn);          // per-ray kernel: every lane fetches its own lines
            atomicAdd(&P.counters[5], (unsigned long long)nt);
        }
    }
}

// ------------------------------------------------------------ brute-force kernel ----
// HW1 contract (HW1/src/render.cpp:89-107): every ray tests every triangle.  The block streams
// the triangle blocks through a double-buffered shared-memory ring (coalesced 16-byte loads,
// broadcast reads), so HBM/L2 traffic is P*48 bytes per 128 rays instead of per ray.
#define RT_BRUTE_CHUNK RT_BLOCK_THREADS

template <bool ANY>
__device__ void brute_pass(const FrameParams& P, float4* s_tri, bool active, const Ray& ray, float det_eps,
                           float tmin, float tmax_excl, Hit& best, bool& blocked) {
    const uint32_t n = P.num_tris;
    const uint32_t nchunks = (n + RT_BRUTE_CHUNK - 1) / RT_BRUTE_CHUNK;
    const float4* __restrict__ g = reinterpret_cast<const float4*>(P.geom);
    auto stage = [&](uint32_t c) {
        float4* dst = s_tri + (c & 1u) * (RT_BRUTE_CHUNK * 3);
        const uint32_t base = c * RT_BRUTE_CHUNK * 3;
        const uint32_t lim = n * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            uint32_t i = base + j * RT_BLOCK_THREADS + threadIdx.x;
            if (i < lim) dst[j * RT_BLOCK_THREADS + threadIdx.x] = __ldg(g + i);
        }
    };
    __syncthreads();           // previous users of s_tri are done
    if (nchunks) stage(0);
    __syncthreads();
    for (uint32_t c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) stage(c + 1);
        if (active) {
            const float4* src = s_tri + (c & 1u) * (RT_BRUTE_CHUNK * 3);
            const uint32_t cnt = min((uint32_t)RT_BRUTE_CHUNK, n - c * RT_BRUTE_CHUNK);
            for (uint32_t j = 0; j < cnt; ++j) {
                const float4 a = src[3 * j], b = src[3 * j + 1], d = src[3 * j + 2];
                Tri tr;
                tr.v0 = mk3(a.x, a.y, a.z); tr.id = RT_F2I(a.w);
                tr.e1 = mk3(b.x, b.y, b.z); tr.e2 = mk3(d.x, d.y, d.z);
                if (ANY) {
                    float t, u, v;
                    if (rt_moller_trumbore(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, FLT_MAX, t, u, v) && t < tmax_excl) {
                        blocked = true; active = false; break;
                    }
                } else {
                    rt_consider(ray, tr, c * RT_BRUTE_CHUNK + j, det_eps, tmin, best);
                }
            }
        }
        __syncthreads();
    }
}

template <int MODE>
__global__ void __launch_bounds__(RT_BLOCK_THREADS)
k_render_brute(const __grid_constant__ FrameParams P) {
    __shared__ float4 s_tri[2 * RT_BRUTE_CHUNK * 3];
    const Pixel px = map_pixel(P);
    const float det_eps = rt_det_eps(MODE), tmin = rt_tmin(MODE);
    unsigned nprim = 0, nshadow = 0;
    f3 accum = mk3(0.f, 0.f, 0.f);
    Hit first; rt_hit_reset(first);
    for (int s = 0; s < P.spp; ++s) {
        const float jx = P.jitter ? __ldg(P.jitter + 2 * s) : 0.0f;
        const float jy = P.jitter ? __ldg(P.jitter + 2 * s + 1) : 0.0f;
        Ray ray = rt_make_ray(P.cam, MODE, px.inside ? px.x : 0, px.inside ? px.y : 0, jx, jy);
        const bool live = px.inside && (MODE == RT_MODE_HW1 || P.max_depth > 0);
        Hit h; rt_hit_reset(h);
        bool dummy = false;
        brute_pass<false>(P, s_tri, live, ray, det_eps, tmin, 0.f, h, dummy);
        if (live) ++nprim;
        f3 color = mk3(0.f, 0.f, 0.f);
        if (MODE == RT_MODE_HW1) {
            if (live) color = rt_shade_hw1(P, ray, h);
        } else {
            Surface sf;
            const bool hit = live && h.slot >= 0;
            if (hit) rt_surface_hw2(P, ray, h, sf);
            for (int l = 0; l < P.num_lights; ++l) {       // block-uniform loop: every thread joins every pass
                const rt_light light = P.lights[l];
                f3 L = mk3(0.f, 0.f, 0.f); float NdotL = 0.f, dist = 0.f; bool need = false; Ray sray = ray;
                bool lit = hit && rt_light_setup_hw2(sf, light, L, NdotL, need, sray, dist);
                need = need && lit && P.shadows;
                bool blocked = false;
                if (__syncthreads_or(need ? 1 : 0)) {
                    Hit unused = h;
                    brute_pass<true>(P, s_tri, need, sray, det_eps, tmin, dist, unused, blocked);
                    if (need) ++nshadow;
                }
                if (lit && !blocked) rt_light_finish_hw2(sf, light, L, NdotL);
            }
            if (live) color = hit ? rt_radiance_hw2(sf.Lo) : rt_radiance_hw2(ld3(P.miss));
        }
        if (s == 0) first = h;
        accum = xadd3(accum, color);
    }
    if (px.inside) rt_write_pixel(P, px.out, accum, first);
    flush_counters(P, nprim, nshadow);
}

// ------------------------------------------------------------ warp-packet kernel ----
// One warp = one 8x4 pixel tile = one 32-ray packet walking the BVH together.  Control flow is
// warp-uniform: a node is visited when __ballot_sync says any lane's slab test passed, the
// descent order is a majority vote, node and triangle loads are single broadcast requests, and
// the traversal stack lives in two registers per lane spread across the warp (entry i sits in
// lane i%32, popped with one __shfl_sync) — no per-lane stack, no divergent branches, no
// reconvergence stalls.  Every lane still culls with its own best t and keeps its own closest
// hit under the canonical rule, so results are identical to the per-ray kernel; a lane merely
// tests a superset of the triangles its own traversal would reach.  Primary rays of a tile and
// their point-light shadow rays are coherent, so the union of visited nodes stays close to a
// single ray's while SIMD efficiency goes from ~44 % (per-ray, ncu r1_v1) to ~100 %.
#define FULLMASK 0xffffffffu
#define RT_PACKET_STACK 64
#ifndef RT_FSTACK
#define RT_FSTACK 128                             // frontier stack of the frustum traversal, entries per warp
#endif
#define RT_SLOT_DEAD (-2)                         // Hit.slot of a lane that traced nothing (outside the frame, depth 0)

struct TraceResult { Hit hit; bool blocked; };

struct WarpStack {
    uint32_t s0, s1; int sp;
    __device__ __forceinline__ void reset() { sp = 0; s0 = s1 = 0u; }
    // caller guarantees sp < RT_PACKET_STACK; entry sp lives in lane sp % 32 of s0 (sp < 32) or s1
    __device__ __forceinline__ void push(uint32_t v, int lane) {
        const bool mine = lane == (sp & 31);
        if (mine && sp < 32) s0 = v;
        if (mine && sp >= 32) s1 = v;
        ++sp;
    }
    __device__ __forceinline__ uint32_t pop() {
        --sp;
        return sp < 32 ? __shfl_sync(FULLMASK, s0, sp) : __shfl_sync(FULLMASK, s1, sp - 32);
    }
};

// any == false: closest hit into `best` for lanes with live == true.
// any == true : live lanes become blocked when a triangle is accepted with t < tlimit (IsInShadow).
// `any` is a run-time, warp-uniform flag so that primary and shadow queries share one copy of the
// loop (the first packet kernel spent 15 % of its issue slots waiting on instruction fetch).
// FAST selects the fused slab test (rt_slab_fma).
// TAG makes the instance private to one kernel: ptxas 12.9 crashes (-lineinfo) on a non-inlined function with
// 256-bit loads that is shared by two entry points.
template <int MODE, bool STATS, bool FAST, bool PF = false, int TAG = 0>
__device__ __noinline__ TraceResult packet_trace(const BvhNode* __restrict__ nodes, const TriBlock* __restrict__ geom, const uint32_t num_tris,
                                                 const Ray ray, bool live, const bool any, float tlimit, TraceStats* st) {
    Hit best; rt_hit_reset(best);
    if (!live) best.slot = RT_SLOT_DEAD;          // lets the caller tell "traced nothing" from "missed" without keeping `live` across the call
    bool blocked = false;
    const float det_eps = rt_det_eps(MODE), tmin = rt_tmin(MODE);
    const int lane = threadIdx.x & 31;
    RayInv k; RayFma kf;
    if (FAST) kf = rt_ray_fma(ray); else k = rt_ray_inv(ray);
    WarpStack stk; stk.reset();
    int cur = 0;
    bool overflow = false;
    const unsigned mlive = __ballot_sync(FULLMASK, live);
    if (!mlive) return TraceResult{best, blocked};
    // Descent order without a per-node vote: LBVH children are split along a known axis with the
    // right child on the high side, so the packet visits the right child first iff most of its rays
    // travel in the negative direction of that axis (bit a of dirneg; axis 3 = no spatial split).
    const int half = __popc(mlive);
    const unsigned dirneg = (2 * __popc(__ballot_sync(FULLMASK, live && ray.d.z < 0.f)) > half ? 1u : 0u) |
                            (2 * __popc(__ballot_sync(FULLMASK, live && ray.d.y < 0.f)) > half ? 2u : 0u) |
                            (2 * __popc(__ballot_sync(FULLMASK, live && ray.d.x < 0.f)) > half ? 4u : 0u);
    if (!any) tlimit = FLT_MAX;                 // closest: the far bound is the lane's best t so far
    while (true) {
        if (cur >= 0) {
            const NodeQ q = rt_load_node(nodes, cur);
            if (STATS) { if (live) st->nodes++; if (lane == 0) st->wnodes++; }
            bool h0, h1;
            if (FAST) {
                h0 = rt_slab_fma_tight(kf, q.q0.x, q.q0.y, q.q0.z, q.q0.w, q.q1.x, q.q1.y, tmin, tlimit);
                h1 = rt_slab_fma_tight(kf, q.q1.z, q.q1.w, q.q2.x, q.q2.y, q.q2.z, q.q2.w, tmin, tlimit);
            } else {
                float tn0, tn1;
                h0 = rt_slab(k, q.q0.x, q.q0.y, q.q0.z, q.q0.w, q.q1.x, q.q1.y, tmin, tlimit, tn0);
                h1 = rt_slab(k, q.q1.z, q.q1.w, q.q2.x, q.q2.y, q.q2.z, q.q2.w, tmin, tlimit, tn1);
            }
            const bool a0 = __any_sync(FULLMASK, h0 && live), a1 = __any_sync(FULLMASK, h1 && live);
            if (a0 && a1) {
                const bool c1first = (((unsigned)q.q3.w >> 29) & dirneg) != 0u;
                const int far = c1first ? q.q3.x : q.q3.y;
                if (stk.sp >= RT_PACKET_STACK) { overflow = true; break; }     // deeper than 64 levels (never on an LBVH of < 2^64 keys; kept as the safety net)
                stk.push((uint32_t)far, lane);
                if (PF) {      // the deferred child will be popped later: start pulling its line into L1 now
                    const void* a = far >= 0 ? (const void*)(nodes + far) : (const void*)(geom + rt_leaf_first(far));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(a));
                }
                cur = c1first ? q.q3.y : q.q3.x;
                continue;
            }
            if (a0) { cur = q.q3.x; continue; }
            if (a1) { cur = q.q3.y; continue; }
        } else {
            const uint32_t first = rt_leaf_first(cur), cnt = rt_leaf_count(cur);
            for (uint32_t s = first; s < first + cnt; ++s) {
                const Tri tr = rt_load_tri(geom, s);
                if (STATS && lane == 0) st->wtris++;
                if (live) {
                    if (STATS) st->tris++;
                    float t, u, v;
                    // closest: accepted t <= best.t (== tlimit); canonical rule min t, then min id
                    if (rt_moller_trumbore(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, any ? FLT_MAX : tlimit, t, u, v)) {
                        if (any) { if (t < tlimit) { blocked = true; live = false; } }
                        else if (t < tlimit || tr.id < best.id) { best.t = t; best.u = u; best.v = v; best.slot = (int)s; best.id = tr.id; tlimit = t; }
                    }
                }
            }
            if (any && !__any_sync(FULLMASK, live)) return TraceResult{best, blocked};
        }
        if (stk.sp == 0) break;
        cur = (int)stk.pop();
    }
    if (overflow) {            // finish by brute force, like the reference (query.h:297-308)
        for (uint32_t s = 0; s < num_tris; ++s) {
            const Tri tr = rt_load_tri(geom, s);
            if (!live) continue;
            float t, u, v;
            if (rt_moller_trumbore(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, any ? FLT_MAX : tlimit, t, u, v)) {
                if (any) { if (t < tlimit) { blocked = true; live = false; } }
                else if (t < tlimit || tr.id < best.id) { best.t = t; best.u = u; best.v = v; best.slot = (int)s; best.id = tr.id; tlimit = t; }
            }
        }
    }
    return TraceResult{best, blocked};
}

// ------------------------------------------------------- frustum-culled wide traversal ----
// packet_trace spends half of the frame's instructions on node visits in which all 32 lanes slab-test the SAME
// two boxes.  Here the roles are swapped for the inner nodes: one LANE = one BOX.  The packet's rays are bounded
// by four side planes and a depth range (a frustum that needs no common apex, see below); a round pops up to four
// nodes from a warp-wide frontier stack in shared memory, every group of eight lanes reads that node's 8-wide view
// (WideNode, rt_core.h: the boxes three left/right steps below it, one 32-byte entry per lane), and each lane tests
// its box against the frustum — 32 boxes per round instead of 2 per visit.  Inner
// boxes that pass are pushed with a ballot-ranked compaction; leaf boxes that pass are handed to the whole warp,
// which culls them per ray with the ordinary slab test (own t limit) and then runs the same exactly-rounded
// triangle tests as packet_trace.  A frustum test can only let a lane reach MORE triangles than its own
// traversal would, so ids, t and colours are unchanged (canonical rule: min t, then min id).
//
// The frustum: for rays o_i + t d_i (t >= 0) pick any w with d_i.w > 0 and let sa_i = (d_i.u)/(d_i.w) for a
// second vector u.  With n = u - s w and s <= min_i sa_i, f(x) = x.n has a non-negative coefficient of t along
// every ray, so each ray stays in the half space x.n >= min_i o_i.n.  The other three sides follow with
// (max sa, -u), (min sb, v), (max sb, -v).  Depth: x.w grows along every ray, so a box lies behind every
// origin or beyond every ray's t limit when its depth interval misses [min o_i.w, max (o_i + tlimit_i d_i).w].
// same_origin: the caller guarantees one origin for all live lanes (camera rays), which saves the five reductions.
// Rounding: the slope bounds are widened by 4e-6 (1 + |s|) (directions are within 60 degrees of w, so |s| < 1.8),
// and every plane offset by feps = 1.6e-5 x the largest coordinate in play (host: scene bounds, camera) — an
// order of magnitude more than the error of the 6-term FMA sums.
// 9 resident blocks (36 warps, 56 registers) measured best for the frustum kernel on C4: 6 -> 2.21 ms, 8 -> 2.07, 9 -> 2.01, 10 -> 2.05, 12 -> 2.04
#ifndef RT_FRUSTUM_MINB
#define RT_FRUSTUM_MINB 9               // persistent kernel
#endif
#ifndef RT_TILE_MINB
#define RT_TILE_MINB 10                 // block-per-tile launch of the frustum traversal: r2, with the shared-memory stash, 10 blocks (48 registers): 9 -> 1.883 ms, 10 -> 1.848 ms
#endif

// sm_100a: float min/max reductions are one instruction (CREDUX.MIN/MAX.F32, result in a uniform register)
__device__ __forceinline__ float warp_fmin(float v) { float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float warp_fmax(float v) { float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float warp_bcast(float v, int src, int lane) { return __int_as_float((int)__reduce_or_sync(FULLMASK, lane == src ? (unsigned)__float_as_int(v) : 0u)); }

template <int MODE, bool STATS, bool FAST, int TAG, bool LAZY = false>
__device__ __noinline__ TraceResult frustum_trace(const BvhNode* __restrict__ nodes, const WideNode* __restrict__ wide, const TriBlock* __restrict__ geom, const uint32_t num_tris,
                                                  const Ray ray, bool live, const bool any, const bool same_origin, float tlimit, TraceStats* st,
                                                  int* __restrict__ wstack, float4* __restrict__ wfr, const float feps, const f3 sc, const float sr2) {
    Hit best; rt_hit_reset(best);
    if (!live) best.slot = RT_SLOT_DEAD;          // lets the caller tell "traced nothing" from "missed" without keeping `live` across the call
    bool blocked = false;
    const float det_eps = rt_det_eps(MODE), tmin = rt_tmin(MODE);
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));       // volatile: kept in a register instead of an S2R + mask per use in the round loop
    if (!any) {   // (shadow rays start on the scene: nothing to gain)
        // packets that miss the scene's bounding sphere altogether (sky, mostly-empty frames) leave before any set-up.
        // Division-free, a dozen instructions.  The host pads the squared radius by 3 % and passes +inf (test always
        // true) unless the origins are close enough for the cancellation error of the expression to stay far below that.
        const f3 oc = mk3(sc.x - ray.o.x, sc.y - ray.o.y, sc.z - ray.o.z);
        const float b = oc.x * ray.d.x + oc.y * ray.d.y + oc.z * ray.d.z, c2 = oc.x * oc.x + oc.y * oc.y + oc.z * oc.z;
        const float dd = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
        live = live && (b * b >= dd * (c2 - sr2)) && (b >= 0.f || c2 <= sr2);
    }
    const unsigned mlive = __ballot_sync(FULLMASK, live);
    if (!mlive) return TraceResult{best, blocked};
    // frame: w = bisector of the first and last live rays' directions (any w with d.w > 0 is valid; this one is central)
    const int l0 = __ffs(mlive) - 1, l1 = 31 - __clz(mlive);
    f3 w = mk3(warp_bcast(ray.d.x, l0, lane) + warp_bcast(ray.d.x, l1, lane), warp_bcast(ray.d.y, l0, lane) + warp_bcast(ray.d.y, l1, lane),
               warp_bcast(ray.d.z, l0, lane) + warp_bcast(ray.d.z, l1, lane));
    {
        const float r = rsqrtf(w.x * w.x + w.y * w.y + w.z * w.z);
        w = mk3(w.x * r, w.y * r, w.z * r);
    }
    f3 u = fabsf(w.x) < 0.7f ? mk3(0.f, w.z, -w.y) : mk3(-w.z, 0.f, w.x);        // w x e_x  or  w x e_y
    {
        const float r = rsqrtf(u.x * u.x + u.y * u.y + u.z * u.z);
        u = mk3(u.x * r, u.y * r, u.z * r);
    }
    const f3 v = mk3(w.y * u.z - w.z * u.y, w.z * u.x - w.x * u.z, w.x * u.y - w.y * u.x);
    const float dw = ray.d.x * w.x + ray.d.y * w.y + ray.d.z * w.z;
    const float dd = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
    // every live direction within 60 degrees of w (NaN / zero directions fail the comparison), else the per-lane traversal
    if (!__all_sync(FULLMASK, !live || (dw > 0.f && dw * dw >= 0.25f * dd)))
        return packet_trace<MODE, STATS, FAST, false, TAG>(nodes, geom, num_tris, ray, live, any, tlimit, st);
    const float BIG = 3.0e38f;
    float samin, samax, sbmin, sbmax;
    {
        const float idw = __fdividef(1.0f, dw);
        const float sa = (ray.d.x * u.x + ray.d.y * u.y + ray.d.z * u.z) * idw, sb = (ray.d.x * v.x + ray.d.y * v.y + ray.d.z * v.z) * idw;
        samin = warp_fmin(live ? sa : BIG); samax = warp_fmax(live ? sa : -BIG);
        sbmin = warp_fmin(live ? sb : BIG); sbmax = warp_fmax(live ? sb : -BIG);
        samin -= 4e-6f * (1.0f + fabsf(samin)); samax += 4e-6f * (1.0f + fabsf(samax));
        sbmin -= 4e-6f * (1.0f + fabsf(sbmin)); sbmax += 4e-6f * (1.0f + fabsf(sbmax));
    }
    float farD = BIG;
    {
        const f3 n0 = mk3(u.x - samin * w.x, u.y - samin * w.y, u.z - samin * w.z), n1 = mk3(samax * w.x - u.x, samax * w.y - u.y, samax * w.z - u.z);
        const f3 n2 = mk3(v.x - sbmin * w.x, v.y - sbmin * w.y, v.z - sbmin * w.z), n3 = mk3(sbmax * w.x - v.x, sbmax * w.y - v.y, sbmax * w.z - v.z);
        const f3 o = ray.o;
        float off0 = o.x * n0.x + o.y * n0.y + o.z * n0.z, off1 = o.x * n1.x + o.y * n1.y + o.z * n1.z;
        float off2 = o.x * n2.x + o.y * n2.y + o.z * n2.z, off3 = o.x * n3.x + o.y * n3.y + o.z * n3.z;
        const float ow = o.x * w.x + o.y * w.y + o.z * w.z;
        float nearD = ow;
        if (!same_origin) {                               // (lane 0 writes the frustum; with one origin its values are everyone's)
            off0 = warp_fmin(live ? off0 : BIG); off1 = warp_fmin(live ? off1 : BIG);
            off2 = warp_fmin(live ? off2 : BIG); off3 = warp_fmin(live ? off3 : BIG);
            nearD = warp_fmin(live ? ow : BIG);
        } else {
            const int l0b = __ffs(mlive) - 1;
            off0 = __shfl_sync(FULLMASK, off0, l0b); off1 = __shfl_sync(FULLMASK, off1, l0b);
            off2 = __shfl_sync(FULLMASK, off2, l0b); off3 = __shfl_sync(FULLMASK, off3, l0b); nearD = __shfl_sync(FULLMASK, nearD, l0b);
        }
        off0 -= feps; off1 -= feps; off2 -= feps; off3 -= feps; nearD -= feps;
        if (any) { farD = warp_fmax(live ? fmaf(tlimit, dw, ow) : -BIG); farD += fabsf(farD) * 2e-6f + feps; }
        else tlimit = FLT_MAX;
        // the frustum lives in shared memory (five broadcast 128-bit loads per round) instead of 20 registers per lane
        if (lane == 0) {
            wfr[0] = make_float4(n0.x, n0.y, n0.z, off0); wfr[1] = make_float4(n1.x, n1.y, n1.z, off1);
            wfr[2] = make_float4(n2.x, n2.y, n2.z, off2); wfr[3] = make_float4(n3.x, n3.y, n3.z, off3);
            wfr[4] = make_float4(w.x, w.y, w.z, nearD);
            wstack[0] = 0;
        }
    }
    RayInv k; RayFma kf;
    if (FAST) kf = rt_ray_fma(ray); else k = rt_ray_inv(ray);
    bool overflow = false;
    __syncwarp();
    int sp = 1;
    while (sp > 0) {
        const int npop = sp < 4 ? sp : 4;
        bool valid = (lane >> 3) < npop;
        int par = valid ? wstack[sp - 1 - (lane >> 3)] : 0;
        sp -= npop;
        __syncwarp();
        // lane (g, k) reads entry k of wide node g: one 256-bit load, no dependent steps.  Idle lanes read node 0's entry
        // (harmless, already cached) and are masked out of the result: cheaper than eight predicated defaults per round.
        float cx, cy, cz, hx, hy, hz;
        int ref;
        const int lines = valid ? 1 : 0;
        {
            const float* ep = reinterpret_cast<const float*>(wide + par) + 8 * (lane & 7);
            float fr, fp;
#ifndef RT_X_NODE_HINT
#define RT_X_NODE_HINT ""
#endif
            asm volatile("ld.global.nc" RT_X_NODE_HINT ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(cx), "=f"(cy), "=f"(cz), "=f"(hx), "=f"(hy), "=f"(hz), "=f"(fr), "=f"(fp) : "l"(ep));
            ref = __float_as_int(fr);
        }
        if (STATS) {
            int tot = lines;
            for (int q = 16; q > 0; q >>= 1) tot += __shfl_xor_sync(FULLMASK, tot, q);
            if (lane == 0) { st->wnodes += (uint32_t)tot; st->nodes += (uint32_t)tot; }     // 32-byte entries requested
        }
        bool hit = valid && hx >= 0.f;                       // absent entries carry negative half extents
        {
            const float4 p0 = wfr[0], p1 = wfr[1], p2 = wfr[2], p3 = wfr[3], pw = wfr[4];
            const float s0 = fmaf(hx, fabsf(p0.x), fmaf(hy, fabsf(p0.y), fmaf(hz, fabsf(p0.z), fmaf(cx, p0.x, fmaf(cy, p0.y, cz * p0.z)))));
            const float s1 = fmaf(hx, fabsf(p1.x), fmaf(hy, fabsf(p1.y), fmaf(hz, fabsf(p1.z), fmaf(cx, p1.x, fmaf(cy, p1.y, cz * p1.z)))));
            const float s2 = fmaf(hx, fabsf(p2.x), fmaf(hy, fabsf(p2.y), fmaf(hz, fabsf(p2.z), fmaf(cx, p2.x, fmaf(cy, p2.y, cz * p2.z)))));
            const float s3 = fmaf(hx, fabsf(p3.x), fmaf(hy, fabsf(p3.y), fmaf(hz, fabsf(p3.z), fmaf(cx, p3.x, fmaf(cy, p3.y, cz * p3.z)))));
            const float dep = fmaf(cx, pw.x, fmaf(cy, pw.y, cz * pw.z)), dh = fmaf(hx, fabsf(pw.x), fmaf(hy, fabsf(pw.y), hz * fabsf(pw.z)));
            hit = hit && s0 >= p0.w && s1 >= p1.w && s2 >= p2.w && s3 >= p3.w && dep + dh >= pw.w && dep - dh <= farD;
        }
        const unsigned minner = __ballot_sync(FULLMASK, hit && ref >= 0);
        unsigned mleaf = __ballot_sync(FULLMASK, hit && ref < 0);
        const int ninner = __popc(minner);
        if (sp + ninner > RT_FSTACK) { overflow = true; break; }
        // (prefetching the pushed wide nodes / the candidates' triangle blocks into L1 was measured: 2.05 vs 2.00 ms —
        // the kernel is issue-bound, the extra instructions cost more than the latency they hide)
        if (hit && ref >= 0) wstack[sp + __popc(minner & ((1u << lane) - 1u))] = ref;
        sp += ninner;
        __syncwarp();
        bool improved = false;
        while (mleaf) {
            const int src = __ffs(mleaf) - 1;
            mleaf &= mleaf - 1u;
            // the candidate's entry again, this time by every lane (uniform address, the line is in L1): one shuffle
            // and one 256-bit load instead of seven shuffles
            const int eidx = __shfl_sync(FULLMASK, par * 8 + (lane & 7), src);
            float lcx, lcy, lcz, lhx, lhy, lhz, lfr, lfp;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(lcx), "=f"(lcy), "=f"(lcz), "=f"(lhx), "=f"(lhy), "=f"(lhz), "=f"(lfr), "=f"(lfp)
                         : "l"(reinterpret_cast<const float*>(wide) + 8 * (size_t)eidx));
            const int lref = __float_as_int(lfr);
            if (STATS && lane == 0) { st->wnodes++; st->nodes++; }
            bool lh; float tn;
            if (FAST) lh = rt_slab_fma_tight(kf, lcx, lcy, lcz, lhx, lhy, lhz, tmin, tlimit);
            else lh = rt_slab(k, lcx, lcy, lcz, lhx, lhy, lhz, tmin, tlimit, tn);
            if (!__any_sync(FULLMASK, lh && live)) continue;
            const uint32_t first = rt_leaf_first(lref), cnt = rt_leaf_count(lref);
#pragma unroll 1
            for (uint32_t s = first; s < first + cnt; ++s) {      // one copy of the test: the kernel is instruction-cache sensitive
                // all three 16-byte loads of the block are issued together, up front (volatile asm: ptxas otherwise sinks the
                // v0 load below the determinant test, which nearly every test passes — a second full memory latency per test)
#ifndef RT_X_EARLY_LOAD
#define RT_X_EARLY_LOAD 1
#endif
                Tri tr;
                if (!RT_X_EARLY_LOAD) {
                    uint32_t sv = s;
                    asm volatile("" : "+r"(sv));
                    tr = rt_load_tri(geom, sv);
                } else {
                    const float4* tp = reinterpret_cast<const float4*>(geom + s);
                    float idf;
#ifndef RT_X_TRI_HINT
#define RT_X_TRI_HINT ""
#endif
                    asm volatile("ld.global.nc" RT_X_TRI_HINT ".v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(tr.v0.x), "=f"(tr.v0.y), "=f"(tr.v0.z), "=f"(idf) : "l"(tp));
                    float p0, p1;
                    asm volatile("ld.global.nc" RT_X_TRI_HINT ".v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(tr.e1.x), "=f"(tr.e1.y), "=f"(tr.e1.z), "=f"(p0) : "l"(tp + 1));
                    asm volatile("ld.global.nc" RT_X_TRI_HINT ".v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(tr.e2.x), "=f"(tr.e2.y), "=f"(tr.e2.z), "=f"(p1) : "l"(tp + 2));
                    tr.id = __float_as_int(idf);
                }
                if (STATS && lane == 0) st->wtris++;
                if (live) {
                    if (STATS) st->tris++;
                    float t, uu, vv;
                    // LAZY: division-free front end, exact test for the survivors (rt_core.h, rt_moller_trumbore_lazy)
                    if (LAZY ? rt_moller_trumbore_lazy(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, any ? FLT_MAX : tlimit, t, uu, vv)
                             : rt_moller_trumbore(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, any ? FLT_MAX : tlimit, t, uu, vv)) {
                        if (any) { if (t < tlimit) { blocked = true; live = false; } }
                        else if (t < tlimit || tr.id < best.id) { best.t = t; best.u = uu; best.v = vv; best.slot = (int)s; best.id = tr.id; tlimit = t; improved = true; }
                    }
                }
            }
            if (any && !__any_sync(FULLMASK, live)) return TraceResult{best, blocked};
        }
        if (!any && __any_sync(FULLMASK, improved)) {        // closest: pull the far plane in to the farthest current hit
            const float4 pw = wfr[4];
            const float hx_ = fmaf(tlimit, ray.d.x, ray.o.x), hy_ = fmaf(tlimit, ray.d.y, ray.o.y), hz_ = fmaf(tlimit, ray.d.z, ray.o.z);
            farD = warp_fmax(live ? fmaf(hx_, pw.x, fmaf(hy_, pw.y, hz_ * pw.z)) : -BIG);
            farD += fabsf(farD) * 4e-6f + feps;
        }
    }
    if (overflow) {            // frontier wider than the shared-memory stack: finish by brute force (as packet_trace does)
        for (uint32_t s = 0; s < num_tris; ++s) {
            const Tri tr = rt_load_tri(geom, s);
            if (!live) continue;
            float t, uu, vv;
            if (rt_moller_trumbore(ray, tr.v0, tr.e1, tr.e2, det_eps, tmin, any ? FLT_MAX : tlimit, t, uu, vv)) {
                if (any) { if (t < tlimit) { blocked = true; live = false; } }
                else if (t < tlimit || tr.id < best.id) { best.t = t; best.u = uu; best.v = vv; best.slot = (int)s; best.id = tr.id; tlimit = t; }
            }
        }
    }
    return TraceResult{best, blocked};
}

// One packet = the 32 pixels of one 8x4 patch (lane q owns pixel q): ray generation, closest-hit trace, one any-hit
// shadow trace per light, shading, resolve into the requested planes.  Shared by the block-per-tile kernel
// (k_render_packet) and the persistent kernel (k_render_persist).
//
// NOTHING stays in registers across a trace call.  ptxas allocates the registers of the (non-inlined) traversal
// function together with its callers': every value that is live across a call site is one register less for the
// traversal loop, and with the 56-register budget of 9 resident blocks the loop then spills (ncu r2: eight LDL/STL per
// round, +13 % instructions, 3x the local-memory traffic, after a version of this function that merely let the compiler
// hoist the surface set-up above the shadow trace).  So the few values that must survive a trace — ray direction, hit,
// the light's contribution, the pixel sum, the packet's tile — live in a shared-memory STASH (one column of
// RT_STASH_WORDS words per thread, stride RT_BLOCK_THREADS, conflict-free; volatile, so that nothing is cached in
// registers or hoisted), and everything else is recomputed from them.  Side effect: the ray is generated once per
// sample instead of three times, and the surface code exists once — the kernel is instruction-cache bound outside the
// traversal loop, so code that is not there is the cheapest code.
enum { ST_DX = 0, ST_DY, ST_DZ, ST_HT, ST_HU, ST_HV, ST_HSLOT, ST_LOX, ST_LOY, ST_LOZ, ST_DIX, ST_DIY, ST_DIZ, ST_ACX, ST_ACY, ST_ACZ, RT_STASH_WORDS };
#define STASH(k) stash[(k) * RT_BLOCK_THREADS]
struct WarpSlot { unsigned tile, sub, nprim, nshadow; };       // per warp, shared memory: the packet in flight + ray counters

template <bool GROUPED>
__device__ __forceinline__ Pixel packet_pixel(const FrameParams& P, const volatile WarpSlot* ws, int lane) {
    return rt_map_pixel(P, (int)ws->tile, (int)(ws->sub * 32u) + lane);
}

template <int MODE, bool STATS, bool FAST, bool PF, bool GROUPED, bool FRUSTUM, bool LAZY, int TAG>
__device__ __forceinline__ void render_packet(const FrameParams& P, volatile WarpSlot* const ws, volatile float* const stash,
                                              int* const wstack, float4* const wfr, TraceStats* const st) {
    const int lane = (int)(threadIdx.x & 31);
    STASH(ST_ACX) = 0.f; STASH(ST_ACY) = 0.f; STASH(ST_ACZ) = 0.f;    // pixel sum
    // Sample-major packets: with G = P.sample_group samples of one pixel side by side in the warp, a pass traces
    // 32/G neighbouring pixels x G samples — a footprint of 32/G pixels instead of 8x4, so the rays of a packet stay
    // together much deeper into the tree when spp > 1 (G == 1: one sample of each of the warp's 32 pixels, as before).
    // Lane q owns pixel q of the warp's 8x4 patch and sums its samples in sample order, like the reference's loop.
    // GROUPED == false is the G == 1 instance without the shuffles (host picks the instance from P.sample_group).
    for (int pass = 0; pass < P.spp; ++pass) {
        const int G = GROUPED ? P.sample_group : 1, gsh = GROUPED ? __ffs(G) - 1 : 0, PP = 32 >> gsh, chunks = P.spp >> gsh;
        const int pb = GROUPED ? pass / chunks : 0, pc = pass - pb * chunks;
        const int qi = GROUPED ? pb * PP + (lane >> gsh) : lane;                 // the pixel (lane of the patch) this lane samples
        const int s = GROUPED ? (pc << gsh) + (lane & (G - 1)) : pass;           // ... and which of its samples
        {   // ---- camera ray + closest hit
            const Pixel px = packet_pixel<GROUPED>(P, ws, lane);
            const bool inside = GROUPED ? __shfl_sync(FULLMASK, (int)px.inside, qi) != 0 : px.inside;
            int x = GROUPED ? __shfl_sync(FULLMASK, px.x, qi) : px.x, y = GROUPED ? __shfl_sync(FULLMASK, px.y, qi) : px.y;
            if (!inside) { x = 0; y = 0; }
            const bool live = inside && (MODE == RT_MODE_HW1 || P.max_depth > 0);
            const float jx = P.jitter ? __ldg(P.jitter + 2 * s) : 0.0f;
            const float jy = P.jitter ? __ldg(P.jitter + 2 * s + 1) : 0.0f;
            const Ray ray = rt_make_ray(P.cam, MODE, x, y, jx, jy);
            STASH(ST_DX) = ray.d.x; STASH(ST_DY) = ray.d.y; STASH(ST_DZ) = ray.d.z;
            const unsigned nlive = (unsigned)__popc(__ballot_sync(FULLMASK, live));
            if (lane == 0) ws->nprim += nlive;
            Hit h;
            if (FRUSTUM) h = frustum_trace<MODE, STATS, FAST, TAG, LAZY>(P.nodes, P.wide, P.geom, P.num_tris, ray, live, false, true, 0.f, st, wstack, wfr, P.frustum_eps, ld3(P.scene_c), P.scene_r2).hit;
            else h = packet_trace<MODE, STATS, FAST, PF, TAG>(P.nodes, P.geom, P.num_tris, ray, live, false, 0.f, st).hit;
            // (a lane that traced nothing comes back with slot == RT_SLOT_DEAD)
            STASH(ST_HT) = h.t; STASH(ST_HU) = h.u; STASH(ST_HV) = h.v; STASH(ST_HSLOT) = __int_as_float(h.slot);
            if (P.tri_id || P.t) {                         // id / t planes describe sample 0
                const Pixel p2 = packet_pixel<GROUPED>(P, ws, lane);
                const bool in2 = GROUPED ? __shfl_sync(FULLMASK, (int)p2.inside, qi) != 0 : p2.inside;
                const unsigned long long out_q = GROUPED ? __shfl_sync(FULLMASK, (unsigned long long)p2.out, qi) : (unsigned long long)p2.out;
                if (s == 0 && in2) {
                    if (P.tri_id) P.tri_id[out_q] = h.slot >= 0 ? h.id : -1;
                    if (P.t) P.t[out_q] = h.slot >= 0 ? h.t : -1.0f;
                }
            }
        }
        f3 color = mk3(0.f, 0.f, 0.f);
        if (MODE == RT_MODE_HW1) {
            Hit h; h.t = STASH(ST_HT); h.u = STASH(ST_HU); h.v = STASH(ST_HV); h.slot = __float_as_int(STASH(ST_HSLOT)); h.id = 0;
            Ray ray; ray.o = ld3(P.cam.center); ray.d = mk3(STASH(ST_DX), STASH(ST_DY), STASH(ST_DZ));
            if (h.slot != RT_SLOT_DEAD) color = rt_shade_hw1(P, ray, h);
        } else {
            // one iteration per light; a frame without lights still takes one (for the ambient + emission term): a single
            // copy of the surface code in the binary.  The surface is rebuilt from the stash in every iteration.
            STASH(ST_LOX) = 0.f; STASH(ST_LOY) = 0.f; STASH(ST_LOZ) = 0.f;
            for (int l = 0; l < (P.num_lights > 0 ? P.num_lights : 1); ++l) {     // warp-uniform loop
                bool need = false;
                Ray sray; sray.o = mk3(0.f, 0.f, 0.f); sray.d = mk3(0.f, 0.f, 1.f);
                float dist = 0.f;
                STASH(ST_DIX) = 0.f; STASH(ST_DIY) = 0.f; STASH(ST_DIZ) = 0.f;      // this light's contribution, evaluated BEFORE its shadow trace
                const int slot = __float_as_int(STASH(ST_HSLOT));
                if (slot >= 0) {
                    Hit h; h.t = STASH(ST_HT); h.u = STASH(ST_HU); h.v = STASH(ST_HV); h.slot = slot; h.id = 0;
                    Ray ray; ray.o = ld3(P.cam.center); ray.d = mk3(STASH(ST_DX), STASH(ST_DY), STASH(ST_DZ));
                    Surface sf; f3 L; float NdotL;
                    rt_surface_hw2(P, ray, h, sf);
                    if (l == 0) { STASH(ST_LOX) = sf.Lo.x; STASH(ST_LOY) = sf.Lo.y; STASH(ST_LOZ) = sf.Lo.z; }     // ambient + emission
                    if (l < P.num_lights) {
                        bool lit = rt_light_setup_hw2(sf, P.lights[l], L, NdotL, need, sray, dist);
                        need = need && lit && P.shadows;
                        if (lit) {
                            sf.Lo = mk3(0.f, 0.f, 0.f);
                            const f3 direct = rt_light_direct_hw2(sf, P.lights[l], L, NdotL);
                            STASH(ST_DIX) = direct.x; STASH(ST_DIY) = direct.y; STASH(ST_DIZ) = direct.z;
                        }
                    }
                }
                if (P.num_lights > 0) {
                    const unsigned nneed = (unsigned)__popc(__ballot_sync(FULLMASK, need));
                    if (lane == 0) ws->nshadow += nneed;
                    bool blocked;
                    if (FRUSTUM) blocked = frustum_trace<MODE, STATS, FAST, TAG, LAZY>(P.nodes, P.wide, P.geom, P.num_tris, sray, need, true, false, dist, st, wstack, wfr, P.frustum_eps, ld3(P.scene_c), P.scene_r2).blocked;
                    else blocked = packet_trace<MODE, STATS, FAST, PF, TAG>(P.nodes, P.geom, P.num_tris, sray, need, true, dist, st).blocked;
                    // an unlit light left a zero contribution: Lo + 0 == Lo bit for bit (no -0 can arise: Lo >= +0)
                    if (!blocked) {
                        const f3 Lo = xadd3(mk3(STASH(ST_LOX), STASH(ST_LOY), STASH(ST_LOZ)), mk3(STASH(ST_DIX), STASH(ST_DIY), STASH(ST_DIZ)));
                        STASH(ST_LOX) = Lo.x; STASH(ST_LOY) = Lo.y; STASH(ST_LOZ) = Lo.z;
                    }
                }
            }
            const int slot = __float_as_int(STASH(ST_HSLOT));
            if (slot != RT_SLOT_DEAD) color = slot >= 0 ? rt_radiance_hw2(mk3(STASH(ST_LOX), STASH(ST_LOY), STASH(ST_LOZ))) : rt_radiance_hw2(ld3(P.miss));
        }
        f3 acc = mk3(STASH(ST_ACX), STASH(ST_ACY), STASH(ST_ACZ));
        if (!GROUPED) {
            acc = xadd3(acc, color);
        } else {                                         // owner lane q collects its pixel's G samples of this pass in order
            const bool owner = lane >= pb * PP && lane < (pb + 1) * PP;
            for (int k = 0; k < G; ++k) {
                const int src = (((lane - pb * PP) << gsh) + k) & 31;
                const f3 v = mk3(__shfl_sync(FULLMASK, color.x, src), __shfl_sync(FULLMASK, color.y, src), __shfl_sync(FULLMASK, color.z, src));
                if (owner) acc = xadd3(acc, v);
            }
        }
        STASH(ST_ACX) = acc.x; STASH(ST_ACY) = acc.y; STASH(ST_ACZ) = acc.z;
    }
    // col / float(spp): query.cu:163, render.cpp:110
    const Pixel px = packet_pixel<GROUPED>(P, ws, lane);
    const f3 fin = xdivs(mk3(STASH(ST_ACX), STASH(ST_ACY), STASH(ST_ACZ)), (float)P.spp);
    if (px.inside && P.rgb) { P.rgb[3 * px.out] = fin.x; P.rgb[3 * px.out + 1] = fin.y; P.rgb[3 * px.out + 2] = fin.z; }
    if (P.rgb8) {
        // 8-bit plane: a row of the warp's 8x4 patch is 24 contiguous bytes.  Six lanes per row assemble one 32-bit word
        // each from two neighbours' packed pixels and store it — 4 partial-sector writes per warp instead of ~12 with
        // byte stores, which is what the NVLink ingress of rank 0 has to absorb from 7 peers in the fused gather.
        uint32_t pk = 0u;
        if (px.inside) {
#pragma unroll 1
            for (int ch = 0; ch < 3; ++ch)
                pk |= (uint32_t)rt_quantise(ch == 0 ? fin.x : ch == 1 ? fin.y : fin.z, P.quantiser) << (8 * ch);
        }
        const int row0 = lane & ~7, kx = lane & 7;
        const unsigned rowmask = 0xffu << row0;
        const bool row_ok = (__ballot_sync(FULLMASK, px.inside) & rowmask) == rowmask;
        const unsigned long long out0 = __shfl_sync(FULLMASK, (unsigned long long)px.out, row0);
        const int p = (4 * kx) / 3, o = (4 * kx) - 3 * p;                  // first pixel / first channel of word kx (kx < 6)
        const uint32_t a = __shfl_sync(FULLMASK, pk, row0 + (p & 7)), b = __shfl_sync(FULLMASK, pk, row0 + ((p + 1) & 7));
        if (row_ok && (out0 & 3ull) == 0ull) {
            if (kx < 6) {
                const uint32_t word = __funnelshift_r(a | (b << 24), b >> 8, 8 * o);
                *reinterpret_cast<uint32_t*>(P.rgb8 + 3ull * out0 + 4 * kx) = word;
            }
        } else if (px.inside) {
            P.rgb8[3 * px.out] = (uint8_t)pk; P.rgb8[3 * px.out + 1] = (uint8_t)(pk >> 8); P.rgb8[3 * px.out + 2] = (uint8_t)(pk >> 16);
        }
    }
}

__device__ __forceinline__ void flush_stats(const FrameParams& P, const TraceStats& st, bool frustum) {
    unsigned nn = st.nodes, nt = st.tris;
    for (int o = 16; o > 0; o >>= 1) { nn += __shfl_xor_sync(FULLMASK, nn, o); nt += __shfl_xor_sync(FULLMASK, nt, o); }
    if ((threadIdx.x & 31) == 0 && P.counters) {
        atomicAdd(&P.counters[2], (unsigned long long)nn);
        atomicAdd(&P.counters[3], (unsigned long long)nt);
        atomicAdd(&P.counters[4], (unsigned long long)(frustum ? (st.wnodes + 1u) / 2u : st.wnodes));   // 64-byte units: one node line per warp visit (per-lane traversal), two 32-byte wide entries (frustum traversal)
        atomicAdd(&P.counters[5], (unsigned long long)st.wtris);    // one 48-byte triangle block per warp test
    }
}
__device__ __forceinline__ void flush_warp_counters(const FrameParams& P, const volatile WarpSlot* ws) {
    if ((threadIdx.x & 31) == 0 && P.counters) {
        if (ws->nprim) atomicAdd(&P.counters[0], (unsigned long long)ws->nprim);
        if (ws->nshadow) atomicAdd(&P.counters[1], (unsigned long long)ws->nshadow);
    }
}

#ifndef RT_X_TILE_LAZY
#define RT_X_TILE_LAZY 1
#endif
// Block-per-tile launch (one 128-thread block = one 16x8 tile, grid = the rank's tile slots): kept as the comparison
// point of the persistent kernel below and for the experimental traversal variants.
template <int MODE, bool STATS, bool FAST, int MINB, bool PF = false, bool GROUPED = false, bool FRUSTUM = false>
__global__ void __launch_bounds__(RT_BLOCK_THREADS, MINB)
k_render_packet(const __grid_constant__ FrameParams P) {
    __shared__ int s_wstack[FRUSTUM ? (RT_BLOCK_THREADS / 32) * RT_FSTACK : 1];
    int* const wstack = s_wstack + (FRUSTUM ? (threadIdx.x >> 5) * RT_FSTACK : 0);
    __shared__ float4 s_wfr[FRUSTUM ? (RT_BLOCK_THREADS / 32) * 5 : 1];
    float4* const wfr = s_wfr + (FRUSTUM ? (threadIdx.x >> 5) * 5 : 0);
    __shared__ float s_stash[RT_STASH_WORDS * RT_BLOCK_THREADS];
    __shared__ WarpSlot s_ws[RT_BLOCK_THREADS / 32];
    constexpr int TAG = MINB * 4 + (GROUPED ? 1 : 0) + (FRUSTUM ? 2 : 0);
    volatile WarpSlot* const ws = s_ws + (threadIdx.x >> 5);
    if ((threadIdx.x & 31) == 0) { ws->tile = blockIdx.x + (unsigned)P.tile_offset; ws->sub = threadIdx.x >> 5; ws->nprim = 0u; ws->nshadow = 0u; }
    __syncwarp();
    TraceStats st{0, 0, 0, 0, 0};
    render_packet<MODE, STATS, FAST, PF, GROUPED, FRUSTUM, RT_X_TILE_LAZY != 0, TAG>(P, ws, s_stash + threadIdx.x, wstack, wfr, STATS ? &st : nullptr);
    __syncwarp();
    flush_warp_counters(P, ws);
    if (STATS) flush_stats(P, st, FRUSTUM);
}

// --------------------------------------------------------------- persistent frame kernel ----
// The default frame kernel.  The grid is sized to the machine (SMs x resident blocks), not to the frame: every block
// pulls 16x8 tiles from the rank's tile queue (one atomicAdd per tile, issued by the first warp that finishes the current one) until it is empty; its
// four warps take the tile's four 8x4 packets and meet at one barrier per tile.
// Completion is published by the kernel itself — no flag kernels, no launch gaps: the rank's tile slots are cut into
// P.num_chunks bands (P.band_end); a block counts the tiles it finished per band and, when it moves on to
// another band (or runs dry), releases them: a block barrier, then one release-ordered atomicAdd on the band's counter
// by thread 0; the block whose add completes a band writes the band's flag word P.flags[band] = P.seq — in this GPU's
// memory or, in the fused multi-GPU gather, in rank 0's memory over NVLink, where the pixels went as well.
// Stream-ordered waits (cuStreamWaitValue32) on those words start the device->host copy of finished rows while later
// bands render.  (Release-only on purpose: atom.release / st.release compile to MEMBAR.ALL.SYS + the access, whereas
// __threadfence_system() and the acquire forms also emit CCTL.IVALL, which throws away the SM's whole L1 — the cache
// this kernel lives on — every time any block of the SM changes band.)
// Multi-GPU handshake, same kernel: rank 0 publishes `ready = seq` at its kernel start (its stream has finished
// reading the previous image by then); the other ranks' blocks wait for it before their first store; rank 0's last
// block waits for every other rank's band flags before the kernel ends (bounded by P.peer_timeout_ns), so the end of
// rank 0's kernel is the end of the gathered frame.
//
// Why the warps of a block stay in step (measured, B200, C4; block-per-tile launch = 1.91 ms):
//   * warp-granular pulling (each warp claims the next packet of the queue the moment it is free: no warp ever idles,
//     tail of one packet): 2.30-2.45 ms;
//   * block-granular tiles with a split barrier of depth two (a warp may run one tile ahead of its siblings): 2.29 ms;
//   * block-granular tiles, one barrier per tile (this kernel): 1.96 ms.
// The four packets of a tile share most of their nodes and triangles; started together they hit each other's lines
// in L1 (hit rate 77 %), drifting apart they do not (60 %), and 36 desynchronised warps per SM stream the ~30 KB of
// straight-line shading code through the instruction cache independently (instruction-fetch stalls 19 % of all stall
// samples against 5 %).  Locality beats the few per cent of idle warp slots.
struct PersistCtl {               // P.queue, zeroed by the host before every launch
    unsigned next_tile;           // global tile-slot counter
    unsigned blocks_done;         // blocks that left the loop
    unsigned pad[2];
    unsigned chunk_done[RT_PEER_MAX_CHUNKS];   // tiles finished per band
};
static_assert(sizeof(PersistCtl) <= RT_PERSIST_CTL_BYTES, "PersistCtl must fit the host's allocation");

__device__ __forceinline__ unsigned long long rt_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned rt_atom_add_release_sys(unsigned* p, unsigned v) { unsigned o; asm volatile("atom.add.release.sys.global.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o; }
__device__ __forceinline__ unsigned rt_atom_add_release_gpu(unsigned* p, unsigned v) { unsigned o; asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o; }
__device__ __forceinline__ void rt_store_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

template <int MODE, bool STATS, bool FAST, int MINB, bool GROUPED, bool LAZY>
__global__ void __launch_bounds__(RT_BLOCK_THREADS, MINB)
k_render_persist(const __grid_constant__ FrameParams P) {
    __shared__ int s_wstack[(RT_BLOCK_THREADS / 32) * RT_FSTACK];
    __shared__ float4 s_wfr[(RT_BLOCK_THREADS / 32) * 5];
    __shared__ float s_stash[RT_STASH_WORDS * RT_BLOCK_THREADS];
    __shared__ WarpSlot s_ws[RT_BLOCK_THREADS / 32];
    __shared__ unsigned s_tile[2];                    // double-buffered: one barrier per tile
    __shared__ unsigned s_claim;                      // iteration whose successor tile has been pulled from the queue
    __shared__ int s_go, s_band;
    __shared__ unsigned s_band_cnt, s_last;
    int* const wstack = s_wstack + (threadIdx.x >> 5) * RT_FSTACK;
    float4* const wfr = s_wfr + (threadIdx.x >> 5) * 5;
    volatile WarpSlot* const ws = s_ws + (threadIdx.x >> 5);
    constexpr int TAG = 1000 + MINB * 8 + (GROUPED ? 1 : 0) + (LAZY ? 2 : 0) + (STATS ? 4 : 0);
    PersistCtl* const ctl = reinterpret_cast<PersistCtl*>(P.queue);
    if ((threadIdx.x & 31) == 0) { ws->sub = threadIdx.x >> 5; ws->nprim = 0u; ws->nshadow = 0u; }
    if (threadIdx.x == 0) {
        s_tile[0] = atomicAdd(&ctl->next_tile, 1u);
        s_claim = 0u;
        s_band = -1; s_band_cnt = 0u;
        int go = 1;
        if (P.ready_out && blockIdx.x == 0) { *(volatile unsigned*)P.ready_out = P.seq; }
        if (P.ready_in) {                             // ranks != 0 of the fused gather: rank 0 must be done with its previous image
            const unsigned long long t0 = rt_globaltimer();
            while ((int)(*(volatile const unsigned*)P.ready_in - P.seq) < 0) {
                if (rt_globaltimer() - t0 > P.peer_timeout_ns) { go = 0; if (P.peer_err) atomicExch(P.peer_err, 1u); break; }
                __nanosleep(100);
            }
        }
        s_go = go;                                    // 0: timed out — trace nothing, store nothing, still publish (rank 0 must not hang on us)
    }
    TraceStats st{0, 0, 0, 0, 0};
    // Releases the tiles this block finished in its current band.  Called by the whole block (block-uniform condition).
    auto release_band = [&]() {
        __syncthreads();                              // every thread's pixel stores happen-before thread 0's release below
        if (threadIdx.x == 0 && s_band_cnt) {
            const unsigned band = (unsigned)s_band, cnt = s_band_cnt;
            const unsigned n = (unsigned)(P.band_end[band] - (band ? P.band_end[band - 1] : 0));
            // pixels in this GPU's own memory: the per-block release only has to reach L2 (the completing block's st.release.sys below
            // is cumulative); pixels stored into another GPU's memory over NVLink: every block releases at system scope
            const unsigned old = P.peer_stores ? rt_atom_add_release_sys(&ctl->chunk_done[band], cnt) : rt_atom_add_release_gpu(&ctl->chunk_done[band], cnt);
            if (old + cnt == n) rt_store_release_sys(P.flags + (size_t)band * RT_PEER_FLAG_STRIDE, P.seq);
            s_band_cnt = 0u;
        }
    };
    for (unsigned it = 0;; ++it) {
        __syncthreads();                              // s_tile[it & 1] is published; everybody is done with the previous tile
        // broadcast with a warp reduction: the result is warp-uniform FOR THE COMPILER (uniform datapath for the tile arithmetic)
        const unsigned tile = __reduce_max_sync(FULLMASK, *(volatile unsigned*)&s_tile[it & 1u]);
        if (tile >= (unsigned)P.local_tiles) break;
        if (P.num_chunks > 0) {
            int b = 0;
            while (b + 1 < P.num_chunks && tile >= (unsigned)P.band_end[b]) ++b;      // <= 16 uniform compares
            if (b != *(volatile int*)&s_band) {       // block-uniform
                if (*(volatile int*)&s_band >= 0) release_band();
                __syncthreads();
                if (threadIdx.x == 0) s_band = b;
            }
            if (threadIdx.x == 0) s_band_cnt = s_band_cnt + 1u;
        }
        if ((threadIdx.x & 31) == 0) ws->tile = tile;
        __syncwarp();
        if (*(volatile int*)&s_go)
            render_packet<MODE, STATS, FAST, false, GROUPED, true, LAZY, TAG>(P, ws, s_stash + threadIdx.x, wstack, wfr, STATS ? &st : nullptr);
        // The FIRST warp to finish its packet pulls the block's next tile: the atomic's round trip (~1 us) overlaps its wait for the
        // siblings at the barrier, and the tile is claimed as late as possible (tiles claimed early and held while another block idles
        // lengthen the tail).  (Pulled by thread 0 at the top of the iteration, the store of the result stalled warp 0 for the round
        // trip before its packet: +2.5 % on the whole frame.)
        if ((threadIdx.x & 31) == 0 && atomicCAS(&s_claim, it, it + 1u) == it) s_tile[(it + 1u) & 1u] = atomicAdd(&ctl->next_tile, 1u);
    }
    if (P.num_chunks > 0 && *(volatile int*)&s_band >= 0) release_band();
    __syncwarp();
    flush_warp_counters(P, ws);
    if (STATS) flush_stats(P, st, true);
    if (P.wait_ranks > 0 || P.host_counters) {
        __syncthreads();                              // this block's counter atomics happen-before thread 0's acq_rel add
        if (threadIdx.x == 0) {
            unsigned old;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(&ctl->blocks_done) : "memory");
            s_last = old + 1u == gridDim.x ? 1u : 0u;
            if (s_last && P.host_counters && P.counters) {      // every block's rays are in: hand the totals to the host
                const unsigned long long a = *(volatile unsigned long long*)&P.counters[0], b = *(volatile unsigned long long*)&P.counters[1];
                P.host_counters[0] = a; P.host_counters[1] = b;
                asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(&P.host_counters[2]), "l"((unsigned long long)P.seq) : "memory");
            }
        }
        __syncthreads();
        // rank 0 of the fused gather: the last block out waits for the other ranks' band flags (they live in this GPU's memory)
        if (s_last && P.wait_ranks > 0) {
            const int nflags = P.wait_ranks * P.num_chunks;
            const unsigned long long t0 = rt_globaltimer();
            for (int i = (int)threadIdx.x; i < nflags; i += RT_BLOCK_THREADS) {
                const int r = 1 + i / P.num_chunks, j = i - (r - 1) * P.num_chunks;
                const volatile unsigned* f = P.flags_base + RT_PEER_CHUNK_FLAG(r, j);
                while ((int)(*f - P.seq) < 0) {
                    if (rt_globaltimer() - t0 > P.peer_timeout_ns) { if (P.peer_err) atomicExch(P.peer_err, 1u + (unsigned)r); break; }
                    __nanosleep(100);
                }
            }
            __threadfence_system();
        }
    }
}

// -------------------------------------------------------------------- tile unpack ----
__global__ void k_unpack(FrameParams P, int src_rank, const float* rgb, const uint8_t* rgb8, const int32_t* tri_id,
                         const float* t, float* o_rgb, uint8_t* o_rgb8, int32_t* o_tri_id, float* o_t, int src_tiles) {
    const int ltile = blockIdx.x;
    if (ltile >= src_tiles) return;
    const long long di = rt_unpack_index(P, src_rank, ltile, (int)threadIdx.x);
    if (di < 0) return;
    const size_t src = (size_t)ltile * RT_BLOCK_THREADS + threadIdx.x;
    const size_t dst = (size_t)di;
    if (o_rgb && rgb) { o_rgb[3 * dst] = rgb[3 * src]; o_rgb[3 * dst + 1] = rgb[3 * src + 1]; o_rgb[3 * dst + 2] = rgb[3 * src + 2]; }
    if (o_rgb8 && rgb8) { o_rgb8[3 * dst] = rgb8[3 * src]; o_rgb8[3 * dst + 1] = rgb8[3 * src + 1]; o_rgb8[3 * dst + 2] = rgb8[3 * src + 2]; }
    if (o_tri_id && tri_id) o_tri_id[dst] = tri_id[src];
    if (o_t && t) o_t[dst] = t[src];
}

// ------------------------------------------------- peer flags (fused NVLink gather) ----
// In the peer-store gather every rank's frame kernel writes its tiles straight into rank 0's
// row-major image over NVLink (CUDA IPC mapping), so the "gather" is only a completion signal:
// a flag word per rank in rank 0's memory.  k_flag_set publishes a frame sequence number after
// everything the stream did before it (kernel boundary + system fence); k_flag_wait spins until
// all `n` flags reached it, bounded by a wall-clock timeout so a dead peer cannot hang the GPU.
__global__ void k_flag_set(volatile unsigned* flags, int stride, int n, unsigned seq) {
    __threadfence_system();
    for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) flags[(size_t)i * stride] = seq;
    __threadfence_system();
}
// waits for flags[(r0 + a) * rank_stride + b * stride] >= seq, a < nranks, b < nper
__global__ void k_flag_wait(const volatile unsigned* flags, int rank_stride, int nranks, int stride, int nper, unsigned seq, unsigned long long timeout_ns, unsigned* err) {
    for (int i = (int)threadIdx.x; i < nranks * nper; i += (int)blockDim.x) {
        const int a = i / nper, b = i - a * nper;
        const volatile unsigned* f = flags + (size_t)a * rank_stride + (size_t)b * stride;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(*f - seq) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { atomicExch(err, 1u + (unsigned)a); break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

// After a timed-out wait: release every stream-ordered wait on the chunk flags (cuStreamWaitValue32 has no timeout).
__global__ void k_flag_unblock(const unsigned* err, volatile unsigned* flags, int stride, int n, unsigned seq) {
    if (*err == 0u) return;
    for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) flags[(size_t)i * stride] = seq;
    __threadfence_system();
}

} // namespace

cudaError_t rt_launch_flag_unblock(const unsigned* err, unsigned* flags, int stride, int n, unsigned seq, cudaStream_t stream) {
    k_flag_unblock<<<1, 64, 0, stream>>>(err, flags, stride, n, seq);
    return cudaGetLastError();
}
cudaError_t rt_launch_flag_set(unsigned* flags, int stride, int n, unsigned seq, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    k_flag_set<<<1, 32, 0, stream>>>(flags, stride, n, seq);
    return cudaGetLastError();
}
cudaError_t rt_launch_flag_wait(const unsigned* flags, int rank_stride, int nranks, int stride, int nper, unsigned seq, unsigned long long timeout_ns, unsigned* err, cudaStream_t stream) {
    if (nranks <= 0 || nper <= 0) return cudaSuccess;
    k_flag_wait<<<1, 128, 0, stream>>>(flags, rank_stride, nranks, stride, nper, seq, timeout_ns, err);
    return cudaGetLastError();
}

// Grid of the persistent kernel: every SM filled to the kernel's occupancy (never more blocks than tiles).
// (All instantiations share one function-pointer type, so the cache is keyed by the pointer.)
static cudaError_t persist_grid(void (*kernel)(const FrameParams), int tiles, unsigned* grid) {
    struct Entry { void (*k)(const FrameParams); int blocks; };
    static Entry cache[32];
    static int ncache = 0;
    int blocks = 0;
    for (int i = 0; i < ncache; ++i) if (cache[i].k == kernel) blocks = cache[i].blocks;
    if (!blocks) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RT_BLOCK_THREADS, 0);
        if (e != cudaSuccess) return e;
        blocks = (per_sm < 1 ? 1 : per_sm) * (sms < 1 ? 1 : sms);
        if (ncache < 32) { cache[ncache].k = kernel; cache[ncache].blocks = blocks; ++ncache; }
    }
    long long g = blocks;
    if (g > tiles) g = tiles;
    *grid = (unsigned)(g < 1 ? 1 : g);
    return cudaSuccess;
}
#define RT_LAUNCH_PERSIST(...) do { unsigned g_ = 1; cudaError_t e_ = persist_grid(k_render_persist<__VA_ARGS__>, fp.local_tiles, &g_); if (e_ != cudaSuccess) return e_; \
                                    k_render_persist<__VA_ARGS__><<<g_, block, 0, stream>>>(fp); } while (0)

// Which frames run on the persistent kernel (see rt_kernels.h).
bool rt_render_is_persistent(const FrameParams& fp, int variant, bool banded) {
    if (fp.accel != RT_ACCEL_BVH || !fp.wide || fp.mode == RT_MODE_HW2_CPU) return false;
    // Bounce frames (max_depth > 1) run on the per-ray kernel.  Measured alternative (round 2): segment 0 as packets in this kernel, the
    // rest of every path per lane with a private stack — bit-identical frames, but SLOWER than the per-ray kernel (frog.json as shipped,
    // 1920x1080, depth 8, diffuse bounces: 1.09 vs 0.87 ms stock view, 2.66 vs 1.55 ms frame-filling view): the scalar continuation
    // inherits the packet kernel's 56-register budget and spills ~100 words per thread where the per-ray kernel has 124 registers.
    if (fp.mode != RT_MODE_HW1 && fp.max_depth > 1) return false;
    if (variant == RT_VARIANT_DEFAULT || variant == RT_VARIANT_STATS) return banded && fp.world > 1;
    return variant == RT_VARIANT_PERSIST || variant == RT_VARIANT_PERSIST_EXACT_MT || variant == RT_VARIANT_PERSIST_OCC8 || variant == RT_VARIANT_PERSIST_OCC10;
}

template <int MODE>
static cudaError_t launch_mode(const FrameParams& fp, int variant, cudaStream_t stream) {
    dim3 grid((unsigned)fp.local_tiles), block(RT_BLOCK_THREADS);
    if (fp.accel == RT_ACCEL_BRUTE) { k_render_brute<MODE><<<grid, block, 0, stream>>>(fp); return cudaGetLastError(); }
    const bool fast = fp.fast_slab != 0;
    if (fp.persist) {
        const bool grouped = fp.sample_group > 1;
        switch (variant) {
        case RT_VARIANT_STATS:
            if (grouped) { if (fast) RT_LAUNCH_PERSIST(MODE, true, true, RT_FRUSTUM_MINB, true, true); else RT_LAUNCH_PERSIST(MODE, true, false, RT_FRUSTUM_MINB, true, true); }
            else { if (fast) RT_LAUNCH_PERSIST(MODE, true, true, RT_FRUSTUM_MINB, false, true); else RT_LAUNCH_PERSIST(MODE, true, false, RT_FRUSTUM_MINB, false, true); }
            break;
        case RT_VARIANT_PERSIST_EXACT_MT: RT_LAUNCH_PERSIST(MODE, false, true, RT_FRUSTUM_MINB, false, false); break;
        case RT_VARIANT_PERSIST_OCC8:     RT_LAUNCH_PERSIST(MODE, false, true, 8, false, true); break;
        case RT_VARIANT_PERSIST_OCC10:    RT_LAUNCH_PERSIST(MODE, false, true, 10, false, true); break;
        default:
            if (grouped) { if (fast) RT_LAUNCH_PERSIST(MODE, false, true, RT_FRUSTUM_MINB, true, true); else RT_LAUNCH_PERSIST(MODE, false, false, RT_FRUSTUM_MINB, true, true); }
            else { if (fast) RT_LAUNCH_PERSIST(MODE, false, true, RT_FRUSTUM_MINB, false, true); else RT_LAUNCH_PERSIST(MODE, false, false, RT_FRUSTUM_MINB, false, true); }
        }
        return cudaGetLastError();
    }
    // Bounce rays (max_depth > 1) are incoherent: the frame goes to the per-ray kernel, whose sample function
    // carries TraceRayIterative's full loop; the packet kernels implement depth 1.
    if (MODE != RT_MODE_HW1 && fp.max_depth > 1)
        variant = (variant == RT_VARIANT_STATS || variant == RT_VARIANT_PER_RAY_STATS || variant == RT_VARIANT_FRUSTUM_STATS || variant == RT_VARIANT_PACKET_STATS) ? RT_VARIANT_PER_RAY_STATS : RT_VARIANT_PER_RAY;
    // default: the frustum traversal, one block per tile; no 8-wide view (it could not be allocated): per-lane packet traversal
    if (variant == RT_VARIANT_DEFAULT || variant == RT_VARIANT_PERSIST || variant == RT_VARIANT_PERSIST_EXACT_MT || variant == RT_VARIANT_PERSIST_OCC8 || variant == RT_VARIANT_PERSIST_OCC10)
        variant = fp.wide ? RT_VARIANT_FRUSTUM : RT_VARIANT_PACKET;
    if (variant == RT_VARIANT_STATS) variant = fp.wide ? RT_VARIANT_FRUSTUM_STATS : RT_VARIANT_PACKET_STATS;
    if ((variant == RT_VARIANT_FRUSTUM || variant == RT_VARIANT_FRUSTUM_STATS) && !fp.wide) return cudaErrorInvalidValue;
    switch (variant) {
    case RT_VARIANT_PACKET_OCC6: case RT_VARIANT_PACKET_OCC10: case RT_VARIANT_PACKET_PREFETCH:   // retired round-1 experiments: the plain per-lane kernel
    case RT_VARIANT_PACKET_PIXEL_MAJOR:
    case RT_VARIANT_PACKET:
        if (fp.sample_group > 1) {
            if (fast) k_render_packet<MODE, false, true, 8, false, true><<<grid, block, 0, stream>>>(fp);
            else k_render_packet<MODE, false, false, 8, false, true><<<grid, block, 0, stream>>>(fp);
        } else {
            if (fast) k_render_packet<MODE, false, true, 8><<<grid, block, 0, stream>>>(fp);
            else k_render_packet<MODE, false, false, 8><<<grid, block, 0, stream>>>(fp);
        }
        break;
    case RT_VARIANT_PACKET_STATS:
        if (fp.sample_group > 1) {
            if (fast) k_render_packet<MODE, true, true, 8, false, true><<<grid, block, 0, stream>>>(fp);
            else k_render_packet<MODE, true, false, 8, false, true><<<grid, block, 0, stream>>>(fp);
        } else {
            if (fast) k_render_packet<MODE, true, true, 8><<<grid, block, 0, stream>>>(fp);
            else k_render_packet<MODE, true, false, 8><<<grid, block, 0, stream>>>(fp);
        }
        break;
    case RT_VARIANT_PACKET_EXACT_SLAB: k_render_packet<MODE, false, false, 8><<<grid, block, 0, stream>>>(fp); break;
    case RT_VARIANT_FRUSTUM:            // round-1 default: block-per-tile launch of the frustum traversal
        if (fp.sample_group > 1) {
            if (fast) k_render_packet<MODE, false, true, RT_TILE_MINB, false, true, true><<<grid, block, 0, stream>>>(fp);
            else k_render_packet<MODE, false, false, RT_TILE_MINB, false, true, true><<<grid, block, 0, stream>>>(fp);
        } else {
            if (fast) k_render_packet<MODE, false, true, RT_TILE_MINB, false, false, true><<<grid, block, 0, stream>>>(fp);
            else k_render_packet<MODE, false, false, RT_TILE_MINB, false, false, true><<<grid, block, 0, stream>>>(fp);
        }
        break;
    case RT_VARIANT_FRUSTUM_STATS:
        if (fp.sample_group > 1) {
            if (fast) k_render_packet<MODE, true, true, RT_TILE_MINB, false, true, true><<<grid, block, 0, stream>>>(fp);
            else k_render_packet<MODE, true, false, RT_TILE_MINB, false, true, true><<<grid, block, 0, stream>>>(fp);
        } else {
            if (fast) k_render_packet<MODE, true, true, RT_TILE_MINB, false, false, true><<<grid, block, 0, stream>>>(fp);
            else k_render_packet<MODE, true, false, RT_TILE_MINB, false, false, true><<<grid, block, 0, stream>>>(fp);
        }
        break;
    case RT_VARIANT_PER_RAY:       k_render_bvh<MODE, false><<<grid, block, 0, stream>>>(fp); break;
    case RT_VARIANT_PER_RAY_STATS: k_render_bvh<MODE, true><<<grid, block, 0, stream>>>(fp); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t rt_launch_render(const FrameParams& fp, int kernel_variant, cudaStream_t stream, int* launches) {
    if (fp.local_tiles <= 0) { if (launches) *launches = 0; return cudaSuccess; }
    if (launches) *launches = 1;
    switch (fp.mode) {
    case RT_MODE_HW1:     return launch_mode<RT_MODE_HW1>(fp, kernel_variant, stream);
    case RT_MODE_HW2_BVH: return launch_mode<RT_MODE_HW2_BVH>(fp, kernel_variant, stream);
    case RT_MODE_HW2_CPU: {   // CPUOnly renderer's contract (N1): per-ray kernel over the BVH (mirror recursion is incoherent)
        if (fp.accel != RT_ACCEL_BVH) return cudaErrorInvalidValue;
        dim3 grid((unsigned)fp.local_tiles), block(RT_BLOCK_THREADS);
        if (kernel_variant == RT_VARIANT_STATS || kernel_variant == RT_VARIANT_PER_RAY_STATS) k_render_bvh<RT_MODE_HW2_CPU, true><<<grid, block, 0, stream>>>(fp);
        else k_render_bvh<RT_MODE_HW2_CPU, false><<<grid, block, 0, stream>>>(fp);
        return cudaGetLastError();
    }
    default:              return cudaErrorInvalidValue;
    }
}

cudaError_t rt_launch_unpack(const FrameParams& fp, int src_rank, const float* rgb, const uint8_t* rgb8,
                             const int32_t* tri_id, const float* t, float* o_rgb, uint8_t* o_rgb8,
                             int32_t* o_tri_id, float* o_t, cudaStream_t stream) {
    const int src_tiles = fp.local_tiles;        // every rank has the same number of tile slots
    if (src_tiles <= 0) return cudaSuccess;
    k_unpack<<<src_tiles, RT_BLOCK_THREADS, 0, stream>>>(fp, src_rank, rgb, rgb8, tri_id, t, o_rgb, o_rgb8, o_tri_id, o_t, src_tiles);
    return cudaGetLastError();
}
