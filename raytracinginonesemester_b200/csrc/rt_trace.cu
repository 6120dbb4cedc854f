// rt_trace.cu — ray generation + closest hit + shading + resolve, one fused kernel per frame.
//
// Replaces renderBatchCUDA / normalizeCUDA / render() of the reference
// (HW2/HW2/GPUandCPU/include/query.cu:12-167) and the HW1 pixel loop (HW1/src/render.cpp:72-124).
// Not a port: the reference walks a 16-byte-node + separate-AABB LBVH with a 512-entry local
// stack and fp64 slabs per thread; here a thread block owns a 16x8 pixel tile (4 warps of 8x4
// pixels), nodes are single 64-byte lines carrying both child boxes, the traversal stack is a
// bank-conflict-free shared-memory column per thread, triangles are 48-byte pre-differenced
// blocks in leaf order, and shadow rays are any-hit queries in the same kernel.  The
// Möller–Trumbore and shading arithmetic is exactly rounded (rt_math.h) so hit ids, t and
// colours equal the reference CPU build.
#include "rt_kernels.h"
#include "rt_trace_core.h"

#include <cuda_runtime.h>

namespace {

__device__ __forceinline__ Pixel map_pixel(const FrameParams& P) { return rt_map_pixel(P, (int)blockIdx.x, (int)threadIdx.x); }

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
    TraceStats st{0, 0, 0};
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

} // namespace

cudaError_t rt_launch_render(const FrameParams& fp, int kernel_variant, cudaStream_t stream, int* launches) {
    const bool stats = kernel_variant == RT_VARIANT_STATS;
    if (fp.local_tiles <= 0) { if (launches) *launches = 0; return cudaSuccess; }
    dim3 grid((unsigned)fp.local_tiles), block(RT_BLOCK_THREADS);
    const bool brute = fp.accel == RT_ACCEL_BRUTE;
    switch (fp.mode) {
    case RT_MODE_HW1:
        if (brute) k_render_brute<RT_MODE_HW1><<<grid, block, 0, stream>>>(fp);
        else if (stats) k_render_bvh<RT_MODE_HW1, true><<<grid, block, 0, stream>>>(fp);
        else k_render_bvh<RT_MODE_HW1, false><<<grid, block, 0, stream>>>(fp);
        break;
    case RT_MODE_HW2_BVH:
        if (brute) k_render_brute<RT_MODE_HW2_BVH><<<grid, block, 0, stream>>>(fp);
        else if (stats) k_render_bvh<RT_MODE_HW2_BVH, true><<<grid, block, 0, stream>>>(fp);
        else k_render_bvh<RT_MODE_HW2_BVH, false><<<grid, block, 0, stream>>>(fp);
        break;
    default:
        return cudaErrorInvalidValue;
    }
    if (launches) *launches = 1;
    return cudaGetLastError();
}

cudaError_t rt_launch_unpack(const FrameParams& fp, int src_rank, const float* rgb, const uint8_t* rgb8,
                             const int32_t* tri_id, const float* t, float* o_rgb, uint8_t* o_rgb8,
                             int32_t* o_tri_id, float* o_t, cudaStream_t stream) {
    const int total = fp.tiles_x * fp.tiles_y;
    const int src_tiles = (total - src_rank + fp.world - 1) / fp.world;
    if (src_tiles <= 0) return cudaSuccess;
    k_unpack<<<src_tiles, RT_BLOCK_THREADS, 0, stream>>>(fp, src_rank, rgb, rgb8, tri_id, t, o_rgb, o_rgb8, o_tri_id, o_t, src_tiles);
    return cudaGetLastError();
}
