// rt_build.cu — device BVH construction: triangle bounds + 63-bit Morton keys, radix sort,
// Karras-2012 hierarchy, atomic bottom-up refit, subtree collapse into multi-triangle leaves
// and emission of the flattened 64-byte-node / 48-byte-triangle-block arena.
//
// Replaces calculateAABBs + buildBVH + buildTrianglesKernel of the reference
// (HW2/HW2/GPUandCPU/include/bvh.cu:7-206, bvh.h:131-289, src/main.cu:19-41).  Not a port:
// the reference keeps a 2P-1 array of 16-byte topology nodes next to a 2P-1 array of AABBs,
// 30-bit codes (1024^3 cells, saturating at ~1M triangles — SURVEY §7 H5) and one triangle
// per leaf; this build uses 63-bit codes, collapses subtrees of <= leaf_max triangles into
// contiguous leaf ranges (Karras ranges are contiguous in sorted order, so no extra
// permutation is needed), compacts the surviving internal nodes with a prefix sum and writes
// each node once as a single line holding both children's padded boxes.  Closest-hit results
// do not depend on the topology (canonical min-t / min-id rule), only speed does.
#include "rt_kernels.h"
#include "rt_build_core.h"

#include <cub/cub.cuh>

namespace {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// Float min/max through integer atomics: non-negative floats order like signed ints, negative floats like unsigned ints
// reversed.  -0.0f (bit pattern 0x80000000 = INT_MIN) satisfies `v >= 0.f` but would win every signed atomicMin and lose
// every signed atomicMax, so the value is canonicalised first (v + 0.0f maps -0.0 to +0.0; the branch is on the sign BIT).
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    v += 0.0f;
    if (!(__float_as_uint(v) >> 31)) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    v += 0.0f;
    if (!(__float_as_uint(v) >> 31)) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}

__global__ void k_init_bounds(Bounds* b) {
    for (int a = 0; a < 3; ++a) { b->lo[a] = INFINITY; b->hi[a] = -INFINITY; }
}

// Scene bounds (thrust::reduce of main.cu:264-270): grid-stride accumulation in registers, warp shuffle + shared-memory
// block reduction, six float atomics per block (min/max are exact, so the order of the reduction does not matter).
// (One atomic set per warp — 1.9M contended atomics at 10M triangles — took 1.25 ms, DRAM at 2 % of peak.)
__global__ void k_scene_bounds(BuildParams bp, Bounds* scene) {
    __shared__ float s_red[8][6];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < bp.num_tris; i += gridDim.x * blockDim.x) {
        f3 a, b, c; uint32_t ia, ib, ic; float l[3], h[3];
        rt_tri_verts(bp, i, a, b, c, ia, ib, ic);
        rt_tri_box(a, b, c, l, h);
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], l[k]); hi[k] = fmaxf(hi[k], h[k]); }
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) for (int k = 0; k < 3; ++k) { s_red[warp][k] = lo[k]; s_red[warp][3 + k] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int nw = (int)(blockDim.x >> 5), k = (int)threadIdx.x;
        float v = s_red[0][k];
        for (int w = 1; w < nw; ++w) v = k < 3 ? fminf(v, s_red[w][k]) : fmaxf(v, s_red[w][k]);
        if (k < 3) { if (v != INFINITY) atomic_min_f(&scene->lo[k], v); }
        else { if (v != -INFINITY) atomic_max_f(&scene->hi[k - 3], v); }
    }
}

// Replaces ComputeMortonCodes (bvh.cu:34-55).
__global__ void k_morton(BuildParams bp, const Bounds* scene, uint64_t* keys, uint32_t* vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bp.num_tris) return;
    f3 a, b, c; uint32_t ia, ib, ic;
    rt_tri_verts(bp, i, a, b, c, ia, ib, ic);
    keys[i] = rt_morton63(a, b, c, *scene);
    vals[i] = i;
}

__global__ void k_karras(const uint64_t* __restrict__ keys, int n, Topo* topo, uint32_t* parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const Topo tp = rt_karras_node(keys, n, i);
    topo[i] = tp;
    parent[tp.left] = (uint32_t)i;
    parent[tp.right] = (uint32_t)i;
    if (i == 0) parent[0] = 0xFFFFFFFFu;
}

// Padded leaf boxes + bottom-up merge: the second thread to reach a node merges its children
// (replaces the atomicCAS refit of bvh.cu:172-203).
__global__ void k_refit(BuildParams bp, const uint32_t* __restrict__ vals, int n, const Topo* __restrict__ topo,
                        const uint32_t* __restrict__ parent, const Bounds* scene, float4* blo, float4* bhi, int* flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    f3 a, b, c; uint32_t ia, ib, ic;
    rt_tri_verts(bp, vals[k], a, b, c, ia, ib, ic);
    float lo[3], hi[3];
    rt_padded_leaf_box(a, b, c, *scene, lo, hi);
    uint32_t node = (uint32_t)(n - 1 + k);
    blo[node] = make_float4(lo[0], lo[1], lo[2], 0.f);
    bhi[node] = make_float4(hi[0], hi[1], hi[2], 0.f);
    if (n == 1) return;
    __threadfence();
    uint32_t p = parent[node];
    while (p != 0xFFFFFFFFu) {
        if (atomicAdd(&flags[p], 1) == 0) return;      // first arrival: the sibling subtree is not done yet
        __threadfence();
        const Topo tp = topo[p];
        float4 l0 = __ldcg(&blo[tp.left]), h0 = __ldcg(&bhi[tp.left]);
        float4 l1 = __ldcg(&blo[tp.right]), h1 = __ldcg(&bhi[tp.right]);
        blo[p] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
        bhi[p] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        __threadfence();
        p = parent[p];
    }
}

__global__ void k_keep_flags(const Topo* __restrict__ topo, int n, uint32_t leaf_max, uint32_t* keep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    keep[i] = (RT_TOPO_LAST(topo[i]) - topo[i].first + 1u) > leaf_max ? 1u : 0u;
}

__global__ void k_emit_nodes(const Topo* __restrict__ topo, int n, const uint32_t* __restrict__ keep,
                             const uint32_t* __restrict__ newidx, const uint32_t* __restrict__ parent,
                             const float4* __restrict__ blo, const float4* __restrict__ bhi, BvhNode* nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !keep[i]) return;
    const Topo tp = topo[i];
    const BvhNode nd = rt_make_node(blo[tp.left], bhi[tp.left], blo[tp.right], bhi[tp.right],
                                    rt_child_ref(tp.left, n, topo, keep, newidx), rt_child_ref(tp.right, n, topo, keep, newidx),
                                    tp.first | RT_NODE_DEPTH3_BITS(rt_node_depth(parent, (uint32_t)i)),
                                    (RT_TOPO_LAST(tp) - tp.first + 1u) | RT_NODE_AXIS_BITS(RT_TOPO_AXIS(tp)));
    float4* dst = reinterpret_cast<float4*>(nodes + newidx[i]);
    const float4* src = reinterpret_cast<const float4*>(&nd);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
}

// Root for scenes with <= leaf_max triangles: one node, child 0 = everything, child 1 = absent.
__global__ void k_emit_single(int n, const float4* __restrict__ blo, const float4* __restrict__ bhi, BvhNode* nodes) {
    if (blockIdx.x || threadIdx.x) return;
    float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
    for (int k = 0; k < n; ++k) {
        float4 l = blo[n - 1 + k], h = bhi[n - 1 + k];
        lo.x = fminf(lo.x, l.x); lo.y = fminf(lo.y, l.y); lo.z = fminf(lo.z, l.z);
        hi.x = fmaxf(hi.x, h.x); hi.y = fmaxf(hi.y, h.y); hi.z = fmaxf(hi.z, h.z);
    }
    const float4 elo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), ehi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
    nodes[0] = rt_make_node(lo, hi, elo, ehi, rt_leaf_ref(0u, (uint32_t)n), rt_leaf_ref(0u, 1u), 0u, (uint32_t)n);
}

// Triangle blocks in slot order (vals == NULL: identity order).
__global__ void k_pack_tris(BuildParams bp, const uint32_t* __restrict__ vals, TriBlock* geom, TriBlock* shade) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= bp.num_tris) return;
    rt_pack_tri(bp, vals ? vals[k] : k, geom + k, shade ? shade + k : nullptr);
}

inline unsigned blocks_for(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

// first position whose vertex index is out of range (atomicMin over 64-bit positions; ~0 = none)
__global__ void k_validate_indices(const uint32_t* __restrict__ idx, size_t n, uint32_t num_vertices, unsigned long long* bad) {
    unsigned long long first = ~0ull;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (idx[i] >= num_vertices && (unsigned long long)i < first) first = (unsigned long long)i;
    if (first != ~0ull) atomicMin(bad, first);
}

} // namespace

cudaError_t rt_validate_indices(const uint32_t* idx, size_t n, uint32_t num_vertices, unsigned long long* bad, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    unsigned nb = blocks_for(n, 256);
    if (nb > 148u * 16u) nb = 148u * 16u;
    k_validate_indices<<<nb, 256, 0, stream>>>(idx, n, num_vertices, bad);
    return cudaGetLastError();
}

// Per-object transform bake in place (main.cu:75-96), one thread per vertex of the object.
__global__ void k_bake_transform(float* pos, float* nrm, size_t first, size_t count, BakeXform T) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float* p = pos + 3 * (first + i);
    const f3 q = rt_bake_point(mk3(p[0], p[1], p[2]), T);
    p[0] = q.x; p[1] = q.y; p[2] = q.z;
    if (nrm) {
        float* n = nrm + 3 * (first + i);
        const f3 m = rt_bake_normal(mk3(n[0], n[1], n[2]), T);
        n[0] = m.x; n[1] = m.y; n[2] = m.z;
    }
}

cudaError_t rt_bake_transform(float* pos, float* nrm, size_t first, size_t count, const BakeXform& T, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    k_bake_transform<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(pos, nrm, first, count, T);
    return cudaGetLastError();
}

// ---- 8-wide view, compact ----
// The frustum traversal hops three BVH2 levels at a time, so only the BVH2 nodes at every third depth ever serve as wide nodes
// (a leaf met early ends its path).  Only those get a WideNode, and their references to each other are indices into the compact
// array (round 1 built one WideNode per BVH2 node: 139 of 270 MB on C4, 1.39 GB on C5, most of it never read).  Which third is
// free to choose: with phase f the root hops f levels (3 when f = 0) and every later wide node sits at depth = f (mod 3).  A
// complete tree would make the three choices differ 4:2:1; on a real LBVH the leaf depths spread over many levels and the
// thirds come out equal to 2 % (C4: 33.5 / 32.8 / 33.7 % of the nodes), so the caller normally asks for phase 0.  Derived data,
// rebuilt by every rank from its copy of the nodes: each node carries its depth modulo 3 (RT_NODE_DEPTH3, k_emit_nodes).
__global__ void k_depth_census(const BvhNode* __restrict__ nodes, uint32_t num_nodes, unsigned* __restrict__ census) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned cls = i < num_nodes ? RT_NODE_DEPTH3(nodes[i].first_slot) : 3u;
    for (unsigned c = 0; c < 3; ++c) {
        const unsigned n = __popc(__ballot_sync(0xffffffffu, cls == c));
        if ((threadIdx.x & 31) == 0 && n) atomicAdd(&census[c], n);
    }
}
__global__ void k_wide_flags(const BvhNode* __restrict__ nodes, uint32_t num_nodes, unsigned phase, uint32_t* __restrict__ need) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < num_nodes) need[i] = (i == 0 || RT_NODE_DEPTH3(nodes[i].first_slot) == phase) ? 1u : 0u;
}
// One thread per needed BVH2 node: expand three levels (the root: `root_levels`) into the node's eight wide entries
// (rt_wide_node, rt_build_core.h) and translate the inner references into compact indices.
__global__ void k_build_wide(const BvhNode* __restrict__ nodes, uint32_t num_nodes, const uint32_t* __restrict__ need,
                             const uint32_t* __restrict__ widx, int root_levels, WideNode* __restrict__ wide) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_nodes || !need[i]) return;
    WideNode w = rt_wide_node(nodes, i, i == 0 ? root_levels : 3);
    for (int k = 0; k < 8; ++k) if (w.e[k].ref >= 0) w.e[k].ref = (int32_t)widx[w.e[k].ref];
    wide[widx[i]] = w;
}

// `phase`: 0, 1 or 2 forces that third; anything else picks the smallest.
cudaError_t rt_build_wide(const BvhNode* nodes, uint32_t num_nodes, int phase, WideNode** wide_out, uint32_t* count_out, cudaStream_t stream) {
    *wide_out = nullptr; *count_out = 0;
    if (num_nodes == 0) return cudaSuccess;
    uint32_t *d_need = nullptr, *d_widx = nullptr; unsigned* d_census = nullptr; void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t err = cudaSuccess;
    WideNode* wide = nullptr;
    unsigned census[3] = {0, 0, 0};
#define CKW(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
    CKW(cudaMallocAsync(&d_need, sizeof(uint32_t) * (size_t)num_nodes, stream));
    CKW(cudaMallocAsync(&d_widx, sizeof(uint32_t) * (size_t)num_nodes, stream));
    CKW(cudaMallocAsync(&d_census, sizeof census, stream));
    CKW(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_need, d_widx, (int)num_nodes, stream));
    CKW(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 16, stream));
    CKW(cudaMemsetAsync(d_census, 0, sizeof census, stream));
    k_depth_census<<<blocks_for(num_nodes, 256), 256, 0, stream>>>(nodes, num_nodes, d_census);
    CKW(cudaMemcpyAsync(census, d_census, sizeof census, cudaMemcpyDeviceToHost, stream));
    CKW(cudaStreamSynchronize(stream));
    {
        // wide nodes per phase: the class itself, plus the root when it is not in the class
        const unsigned size[3] = {census[0], census[1] + 1u, census[2] + 1u};
        if (phase < 0 || phase > 2) { phase = 0; for (int f = 1; f < 3; ++f) if (size[f] < size[phase]) phase = f; }
        const uint32_t count = size[phase];
        k_wide_flags<<<blocks_for(num_nodes, 256), 256, 0, stream>>>(nodes, num_nodes, (unsigned)phase, d_need);
        CKW(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_need, d_widx, (int)num_nodes, stream));
        CKW(cudaMalloc(&wide, sizeof(WideNode) * (size_t)count));
        k_build_wide<<<blocks_for(num_nodes, 128), 128, 0, stream>>>(nodes, num_nodes, d_need, d_widx, phase == 0 ? 3 : phase, wide);
        CKW(cudaGetLastError());
        *wide_out = wide; *count_out = count;
        wide = nullptr;
    }
done:
#undef CKW
    if (wide) cudaFree(wide);
    cudaFreeAsync(d_need, stream); cudaFreeAsync(d_widx, stream); cudaFreeAsync(d_census, stream); cudaFreeAsync(d_tmp, stream);
    return err;
}

cudaError_t rt_pack_triangles(const BuildParams& bp, TriBlock* geom, TriBlock* shade, cudaStream_t stream) {
    if (bp.num_tris == 0) return cudaSuccess;
    k_pack_tris<<<blocks_for(bp.num_tris, 256), 256, 0, stream>>>(bp, nullptr, geom, shade);
    return cudaGetLastError();
}

cudaError_t rt_build_bvh(const BuildParams& bp, BvhNode** nodes_out, TriBlock* geom, TriBlock* shade,
                         BuildResult* res, cudaStream_t stream) {
    const int n = (int)bp.num_tris;
    if (n <= 0) return cudaErrorInvalidValue;
    const unsigned T = 256;
    const uint32_t leaf_max = bp.leaf_max < 1 ? 1 : (bp.leaf_max > 8 ? 8 : bp.leaf_max);

    Bounds* d_scene = nullptr;
    uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
    uint32_t *d_vals = nullptr, *d_vals2 = nullptr, *d_parent = nullptr, *d_keep = nullptr, *d_newidx = nullptr;
    Topo* d_topo = nullptr;
    float4 *d_blo = nullptr, *d_bhi = nullptr;
    int* d_flags = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0, tmp2 = 0;
    const size_t nn = 2 * (size_t)n - 1;
    cudaError_t err = cudaSuccess;
    BvhNode* nodes = nullptr;
    uint32_t num_nodes = 0;

#define CKG(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
    CKG(cudaMallocAsync(&d_scene, sizeof(Bounds), stream));
    CKG(cudaMallocAsync(&d_keys, sizeof(uint64_t) * n, stream));
    CKG(cudaMallocAsync(&d_keys2, sizeof(uint64_t) * n, stream));
    CKG(cudaMallocAsync(&d_vals, sizeof(uint32_t) * n, stream));
    CKG(cudaMallocAsync(&d_vals2, sizeof(uint32_t) * n, stream));
    CKG(cudaMallocAsync(&d_parent, sizeof(uint32_t) * nn, stream));
    CKG(cudaMallocAsync(&d_topo, sizeof(Topo) * (size_t)(n > 1 ? n - 1 : 1), stream));
    CKG(cudaMallocAsync(&d_keep, sizeof(uint32_t) * (size_t)n, stream));
    CKG(cudaMallocAsync(&d_newidx, sizeof(uint32_t) * (size_t)n, stream));
    CKG(cudaMallocAsync(&d_blo, sizeof(float4) * nn, stream));
    CKG(cudaMallocAsync(&d_bhi, sizeof(float4) * nn, stream));
    CKG(cudaMallocAsync(&d_flags, sizeof(int) * (size_t)n, stream));
    CKG(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 63, stream));
    CKG(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_keep, d_newidx, n, stream));
    if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
    CKG(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 16, stream));

    k_init_bounds<<<1, 1, 0, stream>>>(d_scene);
    { unsigned nb = blocks_for(n, T); if (nb > 148u * 8u) nb = 148u * 8u; k_scene_bounds<<<nb, T, 0, stream>>>(bp, d_scene); }
    k_morton<<<blocks_for(n, T), T, 0, stream>>>(bp, d_scene, d_keys, d_vals);
    CKG(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 63, stream));
    CKG(cudaMemsetAsync(d_flags, 0, sizeof(int) * (size_t)n, stream));
    if (n > 1) k_karras<<<blocks_for(n - 1, T), T, 0, stream>>>(d_keys2, n, d_topo, d_parent);
    k_refit<<<blocks_for(n, T), T, 0, stream>>>(bp, d_vals2, n, d_topo, d_parent, d_scene, d_blo, d_bhi, d_flags);
    k_pack_tris<<<blocks_for(n, T), T, 0, stream>>>(bp, d_vals2, geom, shade);

    if ((uint32_t)n > leaf_max) {
        CKG(cudaMemsetAsync(d_keep, 0, sizeof(uint32_t) * (size_t)n, stream));
        k_keep_flags<<<blocks_for(n - 1, T), T, 0, stream>>>(d_topo, n, leaf_max, d_keep);
        CKG(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_keep, d_newidx, n, stream));
        uint32_t last_idx = 0;   // d_keep[n-1] == 0, so newidx[n-1] is the number of kept nodes
        CKG(cudaMemcpyAsync(&last_idx, d_newidx + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CKG(cudaStreamSynchronize(stream));
        num_nodes = last_idx;
        CKG(cudaMalloc(&nodes, sizeof(BvhNode) * (size_t)num_nodes));
        k_emit_nodes<<<blocks_for(n - 1, T), T, 0, stream>>>(d_topo, n, d_keep, d_newidx, d_parent, d_blo, d_bhi, nodes);
    } else {
        num_nodes = 1;
        CKG(cudaMalloc(&nodes, sizeof(BvhNode)));
        k_emit_single<<<1, 32, 0, stream>>>(n, d_blo, d_bhi, nodes);
    }
    CKG(cudaGetLastError());
    {
        Bounds hb;
        CKG(cudaMemcpyAsync(&hb, d_scene, sizeof(Bounds), cudaMemcpyDeviceToHost, stream));
        CKG(cudaStreamSynchronize(stream));
        if (res) {
            res->num_nodes = num_nodes;
            res->num_leaves = (uint32_t)n > leaf_max ? num_nodes + 1u : 1u;   // every kept node has two children, each a kept node or a leaf
            for (int k = 0; k < 3; ++k) { res->scene_min[k] = hb.lo[k]; res->scene_max[k] = hb.hi[k]; }
        }
    }
    *nodes_out = nodes;
    nodes = nullptr;
done:
    if (nodes) cudaFree(nodes);
    cudaFreeAsync(d_scene, stream); cudaFreeAsync(d_keys, stream); cudaFreeAsync(d_keys2, stream);
    cudaFreeAsync(d_vals, stream); cudaFreeAsync(d_vals2, stream); cudaFreeAsync(d_parent, stream);
    cudaFreeAsync(d_topo, stream); cudaFreeAsync(d_keep, stream); cudaFreeAsync(d_newidx, stream);
    cudaFreeAsync(d_blo, stream); cudaFreeAsync(d_bhi, stream); cudaFreeAsync(d_flags, stream);
    cudaFreeAsync(d_tmp, stream);
    return err;
#undef CKG
}
