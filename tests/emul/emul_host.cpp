// emul_host.cpp — TEST INFRASTRUCTURE: runs the product's host/device-compilable per-element
// functions (csrc/rt_build_core.h, rt_trace_core.h) sequentially on the CPU so the BVH build,
// the flattened layout and the traversal/shading logic can be checked against the oracle in a
// container without a GPU (pytest -m "not gpu").  Nothing in the product links this file; the
// product itself has no CPU path.  Built by tests/emul/build_emul.py with
// g++ -O2 -ffp-contract=off (so the X* macros round like the device intrinsics).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../raytracinginonesemester_b200/csrc/rt_build_core.h"
#include "../../raytracinginonesemester_b200/csrc/rt_trace_core.h"
#include "../../raytracinginonesemester_b200/csrc/rt_obj_core.h"
#include <cstdlib>
#include <string>

struct EmuScene {
    std::vector<BvhNode> nodes;
    std::vector<TriBlock> geom, shade;
    std::vector<rt_material> materials;
    uint32_t num_tris = 0;
    bool has_normals = false;
};

extern "C" {

// Mirrors rt_build_bvh (csrc/rt_build.cu) step for step with the same per-element functions.
void* emu_build(const rt_scene* sc, uint32_t leaf_max) {
    const int n = (int)sc->num_triangles;
    if (n <= 0) return nullptr;
    if (leaf_max < 1) leaf_max = 1;
    if (leaf_max > 8) leaf_max = 8;
    EmuScene* es = new EmuScene;
    es->num_tris = (uint32_t)n;
    BuildParams bp{};
    es->has_normals = sc->normals != nullptr;
    bp.positions = sc->positions; bp.normals = sc->normals; bp.indices = sc->indices; bp.obj_ids = sc->tri_obj_ids;
    bp.num_tris = (uint32_t)n; bp.leaf_max = leaf_max;
    if (sc->num_materials > 0) es->materials.assign(sc->materials, sc->materials + sc->num_materials);

    Bounds scene;
    for (int k = 0; k < 3; ++k) { scene.lo[k] = INFINITY; scene.hi[k] = -INFINITY; }
    for (int i = 0; i < n; ++i) {
        f3 a, b, c; uint32_t ia, ib, ic; float lo[3], hi[3];
        rt_tri_verts(bp, (uint32_t)i, a, b, c, ia, ib, ic);
        rt_tri_box(a, b, c, lo, hi);
        for (int k = 0; k < 3; ++k) { scene.lo[k] = fminf(scene.lo[k], lo[k]); scene.hi[k] = fmaxf(scene.hi[k], hi[k]); }
    }
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> vals(n);
    for (int i = 0; i < n; ++i) {
        f3 a, b, c; uint32_t ia, ib, ic;
        rt_tri_verts(bp, (uint32_t)i, a, b, c, ia, ib, ic);
        keys[i] = rt_morton63(a, b, c, scene);
    }
    std::iota(vals.begin(), vals.end(), 0u);
    std::stable_sort(vals.begin(), vals.end(), [&](uint32_t x, uint32_t y) { return keys[x] < keys[y]; });  // radix sort is stable
    std::vector<uint64_t> skeys(n);
    for (int i = 0; i < n; ++i) skeys[i] = keys[vals[i]];

    const size_t nn = 2 * (size_t)n - 1;
    std::vector<Topo> topo(n > 1 ? n - 1 : 1);
    std::vector<uint32_t> parent(nn, 0xFFFFFFFFu);
    for (int i = 0; i + 1 < n; ++i) {
        topo[i] = rt_karras_node(skeys.data(), n, i);
        parent[topo[i].left] = (uint32_t)i; parent[topo[i].right] = (uint32_t)i;
    }
    std::vector<float4> blo(nn), bhi(nn);
    std::vector<int> flags(n, 0);
    for (int k = 0; k < n; ++k) {
        f3 a, b, c; uint32_t ia, ib, ic; float lo[3], hi[3];
        rt_tri_verts(bp, vals[k], a, b, c, ia, ib, ic);
        rt_padded_leaf_box(a, b, c, scene, lo, hi);
        blo[n - 1 + k] = make_float4(lo[0], lo[1], lo[2], 0.f);
        bhi[n - 1 + k] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    for (int k = 0; k < n && n > 1; ++k) {        // same second-arrival walk as k_refit
        uint32_t p = parent[n - 1 + k];
        while (p != 0xFFFFFFFFu) {
            if (flags[p]++ == 0) break;
            const Topo tp = topo[p];
            blo[p] = make_float4(fminf(blo[tp.left].x, blo[tp.right].x), fminf(blo[tp.left].y, blo[tp.right].y), fminf(blo[tp.left].z, blo[tp.right].z), 0.f);
            bhi[p] = make_float4(fmaxf(bhi[tp.left].x, bhi[tp.right].x), fmaxf(bhi[tp.left].y, bhi[tp.right].y), fmaxf(bhi[tp.left].z, bhi[tp.right].z), 0.f);
            p = parent[p];
        }
    }
    es->geom.resize(n); es->shade.resize(n);
    for (int k = 0; k < n; ++k) rt_pack_tri(bp, vals[k], &es->geom[k], &es->shade[k]);

    if ((uint32_t)n > leaf_max) {
        std::vector<uint32_t> keep(n, 0), newidx(n, 0);
        for (int i = 0; i + 1 < n; ++i) keep[i] = (RT_TOPO_LAST(topo[i]) - topo[i].first + 1u) > leaf_max ? 1u : 0u;
        uint32_t run = 0;
        for (int i = 0; i < n; ++i) { newidx[i] = run; run += keep[i]; }
        es->nodes.resize(run);
        for (int i = 0; i + 1 < n; ++i) {
            if (!keep[i]) continue;
            const Topo tp = topo[i];
            es->nodes[newidx[i]] = rt_make_node(blo[tp.left], bhi[tp.left], blo[tp.right], bhi[tp.right],
                                                rt_child_ref(tp.left, n, topo.data(), keep.data(), newidx.data()),
                                                rt_child_ref(tp.right, n, topo.data(), keep.data(), newidx.data()),
                                                tp.first | RT_NODE_DEPTH3_BITS(rt_node_depth(parent.data(), (uint32_t)i)),
                                                (RT_TOPO_LAST(tp) - tp.first + 1u) | RT_NODE_AXIS_BITS(RT_TOPO_AXIS(tp)));
        }
    } else {
        float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
        for (int k = 0; k < n; ++k) {
            lo.x = fminf(lo.x, blo[n - 1 + k].x); lo.y = fminf(lo.y, blo[n - 1 + k].y); lo.z = fminf(lo.z, blo[n - 1 + k].z);
            hi.x = fmaxf(hi.x, bhi[n - 1 + k].x); hi.y = fmaxf(hi.y, bhi[n - 1 + k].y); hi.z = fmaxf(hi.z, bhi[n - 1 + k].z);
        }
        const float4 elo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), ehi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
        es->nodes.push_back(rt_make_node(lo, hi, elo, ehi, rt_leaf_ref(0u, (uint32_t)n), rt_leaf_ref(0u, 1u), 0u, (uint32_t)n));
    }
    return es;
}

// Adopt a BVH downloaded from the device (rt_debug_download_bvh) + shading data rebuilt on the host.
void* emu_adopt(const void* nodes64, uint32_t num_nodes, const void* geom48, uint32_t num_tris, const rt_scene* sc) {
    EmuScene* es = new EmuScene;
    es->num_tris = num_tris;
    es->has_normals = sc->normals != nullptr;
    es->nodes.resize(num_nodes); es->geom.resize(num_tris); es->shade.resize(num_tris);
    memcpy(es->nodes.data(), nodes64, sizeof(BvhNode) * (size_t)num_nodes);
    memcpy(es->geom.data(), geom48, sizeof(TriBlock) * (size_t)num_tris);
    BuildParams bp{};
    bp.positions = sc->positions; bp.normals = sc->normals; bp.indices = sc->indices; bp.obj_ids = sc->tri_obj_ids;
    bp.num_tris = num_tris;
    std::vector<TriBlock> scratch(1);
    for (uint32_t k = 0; k < num_tris; ++k) {
        int tri = RT_F2I(es->geom[k].g[3]);
        rt_pack_tri(bp, (uint32_t)tri, &scratch[0], &es->shade[k]);
    }
    if (sc->num_materials > 0) es->materials.assign(sc->materials, sc->materials + sc->num_materials);
    return es;
}

void emu_free(void* h) { delete (EmuScene*)h; }
uint32_t emu_num_nodes(void* h) { return (uint32_t)((EmuScene*)h)->nodes.size(); }
void emu_export(void* h, void* nodes64, void* geom48) {
    EmuScene* es = (EmuScene*)h;
    if (nodes64) memcpy(nodes64, es->nodes.data(), sizeof(BvhNode) * es->nodes.size());
    if (geom48) memcpy(geom48, es->geom.data(), sizeof(TriBlock) * es->geom.size());
}

// 8-wide view of the emulated BVH, node by node with the product's rt_wide_node (what k_build_wide runs per thread).
void emu_wide(void* h, void* wide256) {
    EmuScene* es = (EmuScene*)h;
    WideNode* w = (WideNode*)wide256;
    for (size_t i = 0; i < es->nodes.size(); ++i) w[i] = rt_wide_node(es->nodes.data(), (uint32_t)i);
}

// The compact, phased view (rt_build_wide, rt_build.cu) step by step on the host: census of the depth classes, phase choice,
// flags, exclusive scan, expansion with translated references.  Returns the number of wide nodes; *phase_io < 0 = smallest.
uint32_t emu_wide_compact(void* h, int* phase_io, void* wide256) {
    EmuScene* es = (EmuScene*)h;
    const uint32_t nn = (uint32_t)es->nodes.size();
    unsigned census[3] = {0, 0, 0};
    for (uint32_t i = 0; i < nn; ++i) census[RT_NODE_DEPTH3(es->nodes[i].first_slot)]++;
    const unsigned size[3] = {census[0], census[1] + 1u, census[2] + 1u};
    int phase = *phase_io;
    if (phase < 0 || phase > 2) { phase = 0; for (int f = 1; f < 3; ++f) if (size[f] < size[phase]) phase = f; }
    *phase_io = phase;
    std::vector<uint32_t> need(nn), widx(nn);
    uint32_t run = 0;
    for (uint32_t i = 0; i < nn; ++i) { need[i] = (i == 0 || RT_NODE_DEPTH3(es->nodes[i].first_slot) == (unsigned)phase) ? 1u : 0u; widx[i] = run; run += need[i]; }
    WideNode* out = (WideNode*)wide256;
    for (uint32_t i = 0; i < nn; ++i) {
        if (!need[i]) continue;
        WideNode w = rt_wide_node(es->nodes.data(), i, i == 0 ? (phase == 0 ? 3 : phase) : 3);
        for (int k = 0; k < 8; ++k) if (w.e[k].ref >= 0) w.e[k].ref = (int32_t)widx[w.e[k].ref];
        out[widx[i]] = w;
    }
    return run == size[phase] ? run : 0xFFFFFFFFu;
}

// Structural check of a flattened BVH: every slot reachable exactly once, child boxes contain
// their triangles, refs in range.  Returns 0 when sound, else a negative code.
int emu_validate(void* h) {
    EmuScene* es = (EmuScene*)h;
    std::vector<int> seen(es->num_tris, 0);
    std::vector<int> stack{0};
    std::vector<int> visited(es->nodes.size(), 0);
    while (!stack.empty()) {
        int ni = stack.back(); stack.pop_back();
        if (ni < 0 || (size_t)ni >= es->nodes.size()) return -1;
        if (visited[ni]++) return -2;
        const BvhNode& nd = es->nodes[ni];
        for (int c = 0; c < 2; ++c) {
            const float* cc = nd.q + 6 * c; const float* hh = cc + 3;     // centre, half-extent
            int32_t ref = c ? nd.ref1 : nd.ref0;
            if (hh[0] < 0.f) continue;   // absent child
            const float lo[3] = {cc[0] - hh[0], cc[1] - hh[1], cc[2] - hh[2]}, hi[3] = {cc[0] + hh[0], cc[1] + hh[1], cc[2] + hh[2]};
            if (ref >= 0) { stack.push_back(ref); continue; }
            uint32_t first = rt_leaf_first(ref), cnt = rt_leaf_count(ref);
            if (first + cnt > es->num_tris) return -3;
            for (uint32_t s = first; s < first + cnt; ++s) {
                if (seen[s]++) return -4;
                const float* g = es->geom[s].g;
                for (int k = 0; k < 3; ++k) {
                    float v0 = g[k], v1 = g[k] + g[4 + k], v2 = g[k] + g[8 + k];
                    float mn = fminf(v0, fminf(v1, v2)), mx = fmaxf(v0, fmaxf(v1, v2));
                    if (mn < lo[k] - 1e-6f * (1.f + fabsf(mn)) || mx > hi[k] + 1e-6f * (1.f + fabsf(mx))) return -5;
                }
            }
        }
    }
    for (uint32_t s = 0; s < es->num_tris; ++s) if (seen[s] != 1) return -6;
    return 0;
}

// Renders through rt_sample_bvh exactly as k_render_bvh does (stack stride 1), tile by tile with
// the kernel's own pixel mapping.  world > 1: the planes in img are this rank's tile-packed buffers
// (local_tiles*128 elements).  stats[0..3] = primary rays, shadow rays, node visits, triangle
// tests; stats[4] = deepest stack use.
int emu_render_rank(void* h, const rt_frame* fr, rt_image* img, uint64_t* stats, int rank, int world, int chunks_per_rank) {
    EmuScene* es = (EmuScene*)h;
    FrameParams P{};
    P.cam = fr->cam; P.mode = fr->mode; P.accel = fr->accel; P.W = fr->width; P.H = fr->height; P.spp = fr->spp;
    P.max_depth = fr->max_depth; P.diffuse_bounce = fr->diffuse_bounce ? 1 : 0; P.shadows = fr->shadows; P.quantiser = fr->quantiser; P.num_lights = fr->num_lights;
    P.num_materials = (int)es->materials.size(); P.has_normals = es->has_normals ? 1 : 0;
    memcpy(P.miss, fr->miss_color, sizeof P.miss);
    P.nodes = es->nodes.data(); P.geom = es->geom.data(); P.shade = es->has_normals ? es->shade.data() : nullptr; P.num_tris = es->num_tris;
    P.materials = es->materials.empty() ? nullptr : es->materials.data();
    P.lights = fr->lights; P.jitter = fr->jitter;
    if (fr->mode == RT_MODE_HW2_CPU) { P.light_radius = fr->light_radius; P.light_samples = fr->light_shadow_samples; }
    P.rng_seed = fr->rng_seed;
    P.tiles_x = (P.W + RT_TILE_W - 1) / RT_TILE_W; P.tiles_y = (P.H + RT_TILE_H - 1) / RT_TILE_H;
    P.rank = rank; P.world = world; P.packed = world > 1 ? 1 : 0;
    P.chunk_tiles = rt_chunk_tiles(P.tiles_x * P.tiles_y, world, chunks_per_rank);
    P.local_tiles = rt_tiles_of_rank(P.tiles_x * P.tiles_y, world, chunks_per_rank);
    P.rgb = img->rgb; P.rgb8 = img->rgb8; P.tri_id = img->tri_id; P.t = img->t;
    unsigned long long tot[5] = {0, 0, 0, 0, 0};
    uint32_t stk[RT_STACK_DEPTH];
    for (int lt = 0; lt < P.local_tiles; ++lt)
        for (int tid = 0; tid < RT_BLOCK_THREADS; ++tid) {
            const Pixel px = rt_map_pixel(P, lt, tid);
            if (!px.inside) continue;
            f3 accum = mk3(0.f, 0.f, 0.f);
            Hit first; rt_hit_reset(first);
            unsigned np = 0, ns = 0;
            TraceStats st{0, 0, 0, 0, 0};
            for (int s = 0; s < P.spp; ++s) {
                Hit hh;
                f3 color = fr->mode == RT_MODE_HW1 ? rt_sample_bvh<RT_MODE_HW1, 1, true>(P, px.x, px.y, s, stk, hh, np, ns, &st)
                         : fr->mode == RT_MODE_HW2_CPU ? rt_sample_bvh<RT_MODE_HW2_CPU, 1, true>(P, px.x, px.y, s, stk, hh, np, ns, &st)
                         : rt_sample_bvh<RT_MODE_HW2_BVH, 1, true>(P, px.x, px.y, s, stk, hh, np, ns, &st);
                if (s == 0) first = hh;
                accum = xadd3(accum, color);
            }
            rt_write_pixel(P, px.out, accum, first);
            tot[0] += np; tot[1] += ns; tot[2] += st.nodes; tot[3] += st.tris;
            if (st.max_sp > tot[4]) tot[4] = st.max_sp;
        }
    img->width = P.W; img->height = P.H; img->rays_primary = tot[0]; img->rays_shadow = tot[1];
    if (stats) for (int k = 0; k < 5; ++k) stats[k] = tot[k];
    return P.local_tiles;
}
int emu_render(void* h, const rt_frame* fr, rt_image* img, uint64_t* stats) {
    emu_render_rank(h, fr, img, stats, 0, 1, 0);
    return 0;
}
// Host mirror of k_unpack for one packed u8 rgb plane of rank src_rank.
void emu_unpack_rgb8(int W, int H, int world, int chunks_per_rank, int src_rank, const uint8_t* packed, uint8_t* image) {
    FrameParams P{};
    P.W = W; P.H = H; P.world = world;
    P.tiles_x = (W + RT_TILE_W - 1) / RT_TILE_W; P.tiles_y = (H + RT_TILE_H - 1) / RT_TILE_H;
    P.chunk_tiles = rt_chunk_tiles(P.tiles_x * P.tiles_y, world, chunks_per_rank);
    const int n = rt_tiles_of_rank(P.tiles_x * P.tiles_y, world, chunks_per_rank);
    for (int lt = 0; lt < n; ++lt)
        for (int e = 0; e < RT_BLOCK_THREADS; ++e) {
            long long di = rt_unpack_index(P, src_rank, lt, e);
            if (di < 0) continue;
            for (int c = 0; c < 3; ++c) image[3 * di + c] = packed[3 * ((size_t)lt * RT_BLOCK_THREADS + e) + c];
        }
}


// ---- division-free front end of the triangle test (rt_core.h, rt_moller_trumbore_lazy_r) vs the reference-order test ----
// Single probe, for the reference's own unit vectors.  mode 0 = HW1 contract, 1 = HW2-BVH contract.
int emu_mt_lazy(int mode, const float* o, const float* d, int normalise, const float* v0, const float* v1, const float* v2, float ra_scale, float* t_out) {
    Ray r; r.o = ld3(o); r.d = normalise ? xunit(ld3(d)) : ld3(d);      // both reference Ray constructors / get_ray normalise
    const f3 a = ld3(v0), e1 = xsub3(ld3(v1), a), e2 = xsub3(ld3(v2), a);
    const int m = mode == 0 ? RT_MODE_HW1 : RT_MODE_HW2_BVH;
    float t, u, v;
    const bool hit = rt_moller_trumbore_lazy_r(r, a, e1, e2, rt_det_eps(m), rt_tmin(m), FLT_MAX, ra_scale, t, u, v);
    if (t_out) *t_out = hit ? t : -1.0f;
    return hit ? 1 : 0;
}

static inline uint64_t xs64(uint64_t& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double urand(uint64_t& s) { return (double)(xs64(s) >> 11) * (1.0 / 9007199254740992.0); }

// n random probes aimed at the places where the two tests could disagree: hit points within 10^-9..10^-1 of an edge, a
// vertex or the t limits, triangles of size 10^-3..10^3 (also slivers), origins near and far, unit and non-unit directions;
// every probe is judged with the approximate reciprocal at its worst (x (1 -+ 2^-22)) and exact.  Returns the number of
// probes where the lazy test rejected something the exact test accepts or returned different (t, u, v): must be 0.
// stats[0] = exact accepts, [1] = probes that left before the IEEE divide, [2] = exact rejects that still paid for it.
uint64_t emu_mt_lazy_sweep(uint64_t n, uint64_t seed, uint64_t* stats) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull, bad = 0, acc = 0, early = 0, late = 0;
    const float scales[3] = {1.0f - 2.3841858e-7f, 1.0f, 1.0f + 2.3841858e-7f};
    for (uint64_t i = 0; i < n; ++i) {
        const double size = pow(10.0, -3.0 + 6.0 * urand(s));
        double V[3][3];
        for (int k = 0; k < 3; ++k) for (int c = 0; c < 3; ++c) V[k][c] = (urand(s) - 0.5) * size;
        if (xs64(s) % 8 == 0) for (int c = 0; c < 3; ++c) V[2][c] = V[0][c] + (V[1][c] - V[0][c]) * urand(s) + (urand(s) - 0.5) * size * 1e-4;   // sliver
        const double off = (urand(s) - 0.5) * size * (xs64(s) % 4 == 0 ? 100.0 : 1.0);
        for (int k = 0; k < 3; ++k) for (int c = 0; c < 3; ++c) V[k][c] += off;
        // barycentric target near the boundary
        const double eps = pow(10.0, -9.0 + 8.0 * urand(s)) * (xs64(s) & 1 ? 1.0 : -1.0);
        double bu = urand(s), bv = urand(s) * (1.0 - bu);
        switch (xs64(s) % 6) {
            case 0: bu = eps; break;                       // edge u = 0
            case 1: bv = eps; break;                       // edge v = 0
            case 2: bv = 1.0 - bu + eps; break;            // edge u + v = 1
            case 3: bu = eps; bv = eps * urand(s); break;  // vertex 0
            case 4: bu = 1.0 + eps; bv = eps * urand(s); break;
            default: break;                                // interior / anywhere
        }
        double Pt[3];
        for (int c = 0; c < 3; ++c) Pt[c] = V[0][c] + bu * (V[1][c] - V[0][c]) + bv * (V[2][c] - V[0][c]);
        double O[3], D[3], len = 0;
        const double dist = size * pow(10.0, -4.0 + 6.0 * urand(s));
        for (int c = 0; c < 3; ++c) { D[c] = urand(s) - 0.5; len += D[c] * D[c]; }
        len = sqrt(len) + 1e-300;
        const double dscale = (xs64(s) & 1) ? 1.0 : pow(10.0, -2.0 + 4.0 * urand(s));     // non-unit directions too (HW1 rays are not all unit)
        const double tsign = (xs64(s) % 16 == 0) ? -1.0 : 1.0;                                 // sometimes behind the origin
        for (int c = 0; c < 3; ++c) { D[c] /= len; O[c] = Pt[c] - tsign * dist * D[c]; D[c] *= dscale; }
        Ray r; r.o = mk3((float)O[0], (float)O[1], (float)O[2]); r.d = mk3((float)D[0], (float)D[1], (float)D[2]);
        const f3 a = mk3((float)V[0][0], (float)V[0][1], (float)V[0][2]);
        const f3 e1 = xsub3(mk3((float)V[1][0], (float)V[1][1], (float)V[1][2]), a), e2 = xsub3(mk3((float)V[2][0], (float)V[2][1], (float)V[2][2]), a);
        const int m = (xs64(s) & 1) ? RT_MODE_HW1 : RT_MODE_HW2_BVH;
        const float tmin = rt_tmin(m);
        // t limits: unbounded, or close to the true t (the closest-hit loop passes its best t as tmax)
        float tmax = FLT_MAX;
        if (xs64(s) % 3 == 0) tmax = (float)(dist / dscale * (1.0 + eps));
        float t0, u0, v0;
        const bool h0 = rt_moller_trumbore(r, a, e1, e2, rt_det_eps(m), tmin, tmax, t0, u0, v0);
        acc += h0;
        for (int k = 0; k < 3; ++k) {
            float t1, u1, v1;
            const bool h1 = rt_moller_trumbore_lazy_r(r, a, e1, e2, rt_det_eps(m), tmin, tmax, scales[k], t1, u1, v1);
            if (h0 != h1 || (h0 && (t0 != t1 || u0 != u1 || v0 != v1))) ++bad;
        }
        if (!h0) {   // did the front end save the divide?  (re-run its comparisons with the exact reciprocal)
            const f3 pvec = xcross(r.d, e2); const float det = xdot(e1, pvec);
            if (fabsf(det) < rt_det_eps(m)) { ++early; continue; }
            const float ra = 1.0f / det; const f3 tvec = xsub3(r.o, a); const float ua = xdot(tvec, pvec) * ra;
            const f3 qvec = xcross(tvec, e1); const float va = xdot(r.d, qvec) * ra, ta = xdot(e2, qvec) * ra;
            if (ua < -RT_LAZY_TINY || ua > 1.00000095367431640625f || va < -RT_LAZY_TINY || ua + va > 1.0000019073486328125f ||
                ta < 0.999996f * tmin - RT_LAZY_TINY || ta > 1.000004f * tmax + RT_LAZY_TINY) ++early; else ++late;
        }
    }
    if (stats) { stats[0] = acc; stats[1] = early; stats[2] = late; }
    return bad;
}


// ---- device OBJ parser, per-line functions (rt_obj_core.h) on the host ----
// parse_real over every '\n'-separated literal of `buf` against glibc's strtof on the same characters.
// out[0] = literals converted on the "device" and equal to strtof bit for bit, out[1] = handed to the host ("hard"),
// out[2] = refused forms (inf / nan / hex), out[3] = no conversion on both sides, out[4] = VALUE mismatches,
// out[5] = end-of-token mismatches (incl. conversion on one side only), out[6] = index of the first mismatch (or -1).
void emu_obj_real_sweep(const char* buf, uint64_t n, long long* out) {
    for (int k = 0; k < 7; ++k) out[k] = 0;
    out[6] = -1;
    uint64_t s = 0; long long idx = 0;
    while (s < n) {
        uint64_t e = s;
        while (e < n && buf[e] != '\n') ++e;
        const uint64_t end = e < n ? e + 1 : e;                 // the line keeps its newline, like fgets
        std::string line(buf + s, buf + end);
        char* ep = nullptr;
        const float want = strtof(line.c_str(), &ep);
        const uint32_t want_end = (uint32_t)(ep - line.c_str());
        DCur c{buf + s, 0u, (uint32_t)(end - s)};
        float got = 0.f; uint32_t tok0 = 0;
        const int r = parse_real(c, got, tok0);
        bool bad_end = false, bad_val = false;
        if (r == 0) { if (want_end != 0) bad_end = true; else out[3]++; }
        else if (r == 3) out[2]++;
        else {
            if (c.i != want_end) bad_end = true;
            if (r == 2) out[1]++;
            else if (memcmp(&got, &want, 4) != 0) bad_val = true;
            else out[0]++;
        }
        if (bad_val) out[4]++;
        if (bad_end) out[5]++;
        if ((bad_val || bad_end) && out[6] < 0) out[6] = idx;
        s = end; ++idx;
    }
}
// parse_face on one line (after the 'f'); corners: 12 ints (v, t, n) x 4.  Returns the corner count.
int emu_obj_face(const char* line, uint32_t n, uint32_t nv, uint32_t nt, uint32_t nn, int* corners) {
    DCur c{line, 0u, n};
    Corner k[4];
    const int nc = parse_face(c, k, nv, nt, nn);
    for (int j = 0; j < nc; ++j) { corners[3 * j] = k[j].v; corners[3 * j + 1] = k[j].t; corners[3 * j + 2] = k[j].n; }
    return nc;
}

} // extern "C"
