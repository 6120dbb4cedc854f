// rt_params.h — parameter blocks shared by the host layer, the kernels and the host-side
// emulation used by the CPU tests (no CUDA runtime types here).
#pragma once

#include "rt_core.h"

// Screen-space tile: the unit of rank ownership (contiguous chunks of tiles, see rt_global_tile) and of
// one thread block (4 warps, each an 8x4 pixel sub-tile).
#define RT_TILE_W 16
#define RT_TILE_H 8
#define RT_BLOCK_THREADS (RT_TILE_W * RT_TILE_H)
#define RT_STACK_DEPTH 32
#define RT_MAX_BANDS 16              // completion flags per rank and frame
#define RT_CPU_MAX_DEPTH 16          // RT_MODE_HW2_CPU: deepest mirror recursion kept on the per-thread fold stack

struct FrameParams {
    rt_camera cam;
    int mode, accel, W, H, spp, max_depth, shadows, quantiser, num_lights, num_materials, diffuse_bounce, has_normals;
    float miss[3];
    // scene arena
    const BvhNode* nodes;
    const WideNode* wide;           // 8-wide view of nodes (frustum traversal); NULL when not built
    const TriBlock* geom;
    const TriBlock* shade;
    uint32_t num_tris;
    const rt_material* materials;   // may be NULL
    const rt_light* lights;
    const float* jitter;            // [2*spp] or NULL
    const float* light_radius;      // RT_MODE_HW2_CPU soft shadows: [num_lights] disk radii or NULL (point lights)
    const int* light_samples;       // ... [num_lights] shadow samples per lit hit or NULL (1)
    uint32_t rng_seed;              // ... mixed into the per-pixel hash RNG seed
    // tile sharding
    int tiles_x, tiles_y, rank, world;
    int local_tiles;                // tile slots of this rank (= grid size)
    int tile_offset;                // first tile slot of this launch (band-pipelined frames launch one kernel per band)
    int chunk_tiles;                // tiles per ownership chunk (see rt_global_tile)
    int packed;                     // 1: planes are this rank's tile-packed buffers (NCCL gather); 0: row-major image
                                    //    (single GPU, or rank 0's image written in place over NVLink peer memory)
    // output planes
    float* rgb; uint8_t* rgb8; int32_t* tri_id; float* t;
    unsigned long long* counters;   // [0] primary rays, [1] shadow rays, [2] node visits, [3] triangle tests (stats variants)
    int fast_slab;                  // 1: ray origins are close enough to the scene for rt_slab_fma (host decides)
    float scene_c[3], scene_r2;     // scene bounding sphere (centre, padded squared radius): packets that miss it skip the traversal set-up
    float frustum_eps;              // frustum traversal: absolute slack of the plane tests = 1.6e-5 x largest coordinate in play (host)
    int sample_group;               // packet kernel: samples of one pixel traced side by side (power of two dividing spp, <= 32)
    int persist;                    // 1: this frame runs on the persistent kernel (host decides: rt_render_is_persistent)
    // persistent kernel (rt_trace.cu, k_render_persist): work queue + in-kernel completion protocol
    unsigned* queue;                // PersistCtl, zeroed by the host before the launch
    int num_chunks;                 // the rank's tile slots are cut into num_chunks bands whose completion is published (0: none);
    int band_end[RT_MAX_BANDS];     // band j = slots [band_end[j-1], band_end[j]) (band_end[-1] = 0): big bands first, small ones last, so that
                                    // the copy of the last band — the only one that cannot overlap rendering — is short
    int peer_stores;                // 1: the output planes live in another GPU's memory (fused gather, ranks != 0)
    unsigned seq;                   // frame sequence number written into flag words
    unsigned* flags;                // this rank's band flags: band j -> flags[j * RT_PEER_FLAG_STRIDE] (own memory, or rank 0's over NVLink)
    const unsigned* flags_base;     // rank 0 of the fused gather: the whole flag block (all ranks' band flags), else unused
    int wait_ranks;                 // rank 0 of the fused gather: ranks 1..wait_ranks are awaited by the last warp; else 0
    const unsigned* ready_in;       // ranks != 0 of the fused gather: rank 0's `ready` word, awaited before the first store; else NULL
    unsigned* ready_out;            // rank 0 of the fused gather: where `ready = seq` is published at kernel start; else NULL
    unsigned long long* host_counters; // persistent kernel: when set, the last block out copies counters[0..1] (primary / shadow rays of the frame)
                                    // here — mapped page-locked host memory — so that rt_render_into needs no device->host read afterwards
    unsigned long long peer_timeout_ns;
    unsigned* peer_err;             // set when a wait above timed out
};

struct BuildParams {
    const float* positions; const float* normals; const uint32_t* indices; const int32_t* obj_ids;
    uint32_t num_tris;
    uint32_t leaf_max;              // max triangles per leaf (<= 8)
};

// Tile ownership.  The frame's tiles in row-major tile order are cut into world*chunks_per_rank contiguous
// chunks of `chunk_tiles` tiles (horizontal bands of the image); chunk c belongs to rank c % world.  Contiguous
// bands keep each GPU's working set to its share of the scene (the round-robin single-tile interleave made
// every GPU walk the whole BVH); several bands per rank spread non-uniform images over the ranks.
// A rank's local tile slots are its chunks back to back: slot = j*chunk_tiles + i  <->  global tile
// (j*world + rank)*chunk_tiles + i; slots past the end of the frame are padding (rendered by nobody).
#define RT_DEFAULT_CHUNKS_PER_RANK 4
RT_HD int rt_chunk_tiles(int total, int world, int chunks_per_rank) {
    const int c = world * (chunks_per_rank > 0 ? chunks_per_rank : RT_DEFAULT_CHUNKS_PER_RANK);
    const int s = (total + c - 1) / c;
    return s > 0 ? s : 1;
}
// number of local tile slots of every rank (padding included)
RT_HD int rt_tiles_of_rank(int total, int world, int chunks_per_rank) {
    if (world <= 1) return total;
    return (chunks_per_rank > 0 ? chunks_per_rank : RT_DEFAULT_CHUNKS_PER_RANK) * rt_chunk_tiles(total, world, chunks_per_rank);
}
RT_HD int rt_global_tile(const FrameParams& P, int rank, int ltile) {
    if (P.world <= 1) return ltile;
    const int j = ltile / P.chunk_tiles, i = ltile - j * P.chunk_tiles;
    const long long g = ((long long)j * P.world + rank) * P.chunk_tiles + i;
    return g < (long long)P.tiles_x * P.tiles_y ? (int)g : -1;
}

// Pixel owned by thread `tid` of the block working on rank-local tile slot `ltile`.  Within a tile each warp
// covers an 8x4 pixel patch.  `out` is the index into the output planes: row-major pixel index
// when P.packed == 0, tile-packed (ltile*128 + ly*16 + lx) when P.packed == 1.
struct Pixel { int x, y; bool inside; size_t out; };
RT_HD Pixel rt_map_pixel(const FrameParams& P, int ltile, int tid) {
    const int gtile = rt_global_tile(P, P.rank, ltile);
    const int g = gtile < 0 ? 0 : gtile;
    const int tx = g % P.tiles_x, ty = g / P.tiles_x;
    const int warp = tid >> 5, lane = tid & 31;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    Pixel px;
    px.x = tx * RT_TILE_W + lx; px.y = ty * RT_TILE_H + ly;
    px.inside = (gtile >= 0) && (px.x < P.W) && (px.y < P.H);
    px.out = (P.packed == 0) ? ((size_t)px.y * P.W + px.x)
                            : ((size_t)ltile * RT_BLOCK_THREADS + (size_t)ly * RT_TILE_W + lx);
    return px;
}
// Unpack: element `e` (= ly*16 + lx) of packed tile slot `ltile` of rank `src_rank` -> row-major pixel
// index, or -1 when the element is padding outside the frame.
RT_HD long long rt_unpack_index(const FrameParams& P, int src_rank, int ltile, int e) {
    const int gtile = rt_global_tile(P, src_rank, ltile);
    if (gtile < 0) return -1;
    const int tx = gtile % P.tiles_x, ty = gtile / P.tiles_x;
    const int x = tx * RT_TILE_W + e % RT_TILE_W, y = ty * RT_TILE_H + e / RT_TILE_W;
    if (x >= P.W || y >= P.H) return -1;
    return (long long)y * P.W + x;
}
