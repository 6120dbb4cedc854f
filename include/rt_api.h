/* rt_api.h — C ABI of the B200 ray-casting hot path.
 *
 * This is the drop-in boundary for the HW1 / HW2 renderers of the reference
 * class repository.  The reference has no FFI layer of its own; its de-facto
 * operator interface is
 *
 *     render(numTriangles, W, H, cam, missColor, max_depth, spp, nodes, aabbs,
 *            triangles, triObjectIds, objectMaterials, numObjectMaterials,
 *            lights, numLights, diffuse_bounce, output)
 *                         HW2/HW2/GPUandCPU/include/query.h:13-29
 *     AccStruct::BVH::calculateAABBs / buildBVH
 *                         HW2/HW2/GPUandCPU/include/bvh.h:412-433
 *     the pixel loop inlined in main()
 *                         HW1/src/render.cpp:60-124
 *
 * Every export below replaces a slice of that interface (file:line cited per
 * function).  Conventions: extern "C"; plain pointers and sizes; caller-owned
 * host buffers are copied on upload and filled on download; int status
 * (0 = ok, <0 = error, text via rt_last_error); one host thread per context;
 * all device work runs on the context's own non-default stream; rt_render is
 * asynchronous, rt_download_image blocks.  There is no CPU fallback: every
 * entry point fails with RT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RT_API_H
#define RT_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_API_VERSION 5

/* status codes */
#define RT_OK            0
#define RT_ERR_ARG      -1   /* bad argument / NULL / out of range                */
#define RT_ERR_CUDA     -2   /* CUDA runtime failure (message has the CUDA text) */
#define RT_ERR_STATE    -3   /* call order: render before upload, etc.           */
#define RT_ERR_NCCL     -4   /* NCCL failure                                      */
#define RT_ERR_NOMEM    -5
#define RT_ERR_UNSUPPORTED -6 /* input outside what the device OBJ parser handles (see rt_dmesh_parse_obj) */

/* rt_frame.mode — which reference renderer's contract the frame follows
 * (SURVEY §8a "mode parameters"). */
#define RT_MODE_HW1      0   /* HW1/src/render.cpp + HW1/include/{ray,raytracer}.h        */
#define RT_MODE_HW2_BVH  1   /* HW2/HW2/GPUandCPU/include/{query,shader,brdf}.h           */
#define RT_MODE_HW2_CPU  2   /* HW2/HW2/CPUOnly/include/{ray,raytracer,brdf}.h: direct light + mirror bounces
                                (TraceRay with diffuse_bounce == false and point lights — the deterministic part;
                                its diffuse bounces and disk lights draw from std::random_device).  Sample position
                                = pixel index + jitter (the reference uses +0.5 at 1 spp, render.cpp:127-131);
                                rt_light.intensity_f; normals == NULL means per-triangle face normals
                                (render.cpp:88-96); miss colour is the sky gradient (raytracer.h:224-230). */

/* rt_frame.accel */
#define RT_ACCEL_BRUTE   0   /* test every triangle (HW1/src/render.cpp:89-107)   */
#define RT_ACCEL_BVH     1   /* BVH traversal (query.h:224-311)                    */

/* rt_frame.outputs / rt_image: which planes are produced */
#define RT_OUT_RGB_F32   1u  /* float rgb[3*W*H], the reference's Vec3 image       */
#define RT_OUT_RGB8      2u  /* uint8 rgb8[3*W*H], quantised on device             */
#define RT_OUT_TRI_ID    4u  /* int32 tri_id[W*H]: closest-hit triangle of sample 0, -1 = miss */
#define RT_OUT_T         8u  /* float t[W*H]: hit distance of sample 0, -1 = miss  */

/* rt_frame.quantiser: float -> u8 rule used for RT_OUT_RGB8 */
#define RT_QUANT_PPM_LROUND   0  /* ppm_p6.cpp:137-155  lround(clamp01(x)*255), gamma off */
#define RT_QUANT_PPM_GAMMA2   1  /* same with sqrt() first (ppm_p6 default gamma2=true)    */
#define RT_QUANT_HW1_TRUNC    2  /* HW1/src/render.cpp:121-123   (uchar)(255.99f*c)        */
#define RT_QUANT_HW2_TRUNC    3  /* GPUandCPU/src/main.cu:428-430 (uchar)(255*min(c,1))    */
#define RT_QUANT_CPU_TRUNC    4  /* CPUOnly/src/render.cpp:157-163 clamp to [0,1], (uchar)(255.99f*c) */

/* rt_frame.kernel_variant */
#define RT_VARIANT_DEFAULT          0  /* packet kernel (one warp = one 8x4 packet), frustum-culled wide traversal, division-free front end of the
                                          triangle test.  rt_render_into: the PERSISTENT launch (blocks pull tiles from a queue and publish finished
                                          bands themselves, = RT_VARIANT_PERSIST); rt_render: one block per tile (= RT_VARIANT_FRUSTUM) */
#define RT_VARIANT_PACKET_OCC6      1  /* retired round-1 experiments 1, 2, 4: accepted, run RT_VARIANT_PACKET */
#define RT_VARIANT_PACKET_OCC10     2
#define RT_VARIANT_PACKET_EXACT_SLAB 3 /* per-lane packet traversal with the unfused (b-o)*inv slab test */
#define RT_VARIANT_PACKET_PREFETCH  4
#define RT_VARIANT_PACKET_PIXEL_MAJOR 5 /* one sample of 32 pixels per packet even when spp > 1 (per-lane traversal) */
#define RT_VARIANT_FRUSTUM          6  /* one block per 16x8 tile, frustum-culled wide traversal (one lane = one box, 32 boxes per round) */
#define RT_VARIANT_PACKET           7  /* one block per tile, per-lane traversal (every lane slab-tests both children of a node) */
#define RT_VARIANT_PERSIST          8  /* the persistent launch also for rt_render */
#define RT_VARIANT_PERSIST_EXACT_MT 9  /* persistent kernel with the reference-order triangle test (IEEE divide first): A/B of the lazy front end */
#define RT_VARIANT_PERSIST_OCC8    11  /* persistent kernel compiled for >= 8 resident blocks per SM (64 registers) */
#define RT_VARIANT_PERSIST_OCC10   12  /* ... >= 10 resident blocks per SM (48 registers) */
#define RT_VARIANT_PER_RAY         10  /* independent per-thread stack traversal (shared-memory stack) */
#define RT_VARIANT_STATS          100  /* default kernel, also counts node data requested / triangle tests (= RT_VARIANT_FRUSTUM_STATS) */
#define RT_VARIANT_PACKET_STATS   107  /* per-lane packet traversal with the same counters               */
#define RT_VARIANT_FRUSTUM_STATS  106  /* frustum kernel with the same counters                        */
#define RT_VARIANT_PER_RAY_STATS  110  /* per-ray kernel with the same counters                        */

/* rt_scene.build_flags */
#define RT_BUILD_DEFAULT      0u
#define RT_BUILD_NO_BVH       1u  /* brute-force frames only; skip the BVH build  */
#define RT_BUILD_LEAF_MAX(n)  (((uint32_t)(n) & 0xFu) << 8)  /* max triangles per BVH leaf, 1..8 (0 = default 2) */

typedef struct rt_ctx rt_ctx;

/* 13 floats, 52 bytes, field order of HW2/HW2/GPUandCPU/include/material.h:6-20 */
typedef struct rt_material {
    float albedo[3];
    float kd;
    float specular_color[3];
    float ks;
    float shininess;
    float kr;
    float emission[3];
} rt_material;

/* 28 bytes, HW2/HW2/GPUandCPU/include/scene.h:21-25 (intensity is an int there).
 * RT_MODE_HW1 uses position+color only (HW1/include/raytracer.h:8-11). */
typedef struct rt_light {
    float   position[3];
    float   color[3];
    union {
        int32_t intensity;      /* RT_MODE_HW1 / RT_MODE_HW2_BVH: int, as in GPUandCPU/include/scene.h:24 */
        float   intensity_f;    /* RT_MODE_HW2_CPU: float, as in CPUOnly/include/raytracer.h:40-42 */
    };
} rt_light;

/* Optional per-object transform baked ON THE DEVICE during rt_upload_scene, replacing the host loop of
 * applyObjectTransform (GPUandCPU/src/main.cu:57-96): for the vertices [first_vertex, first_vertex + num_vertices)
 * p' = Rz(Ry(Rx(p * scale))) + position, n' = normalize(Rz(Ry(Rx(n / scale)))) or (0,0,1) when degenerate —
 * the same operation order and rounding (the six sin/cos values are taken on the host with the C library's
 * sinf/cosf, like the reference).  Ranges must not overlap. */
typedef struct rt_object_transform {
    uint64_t first_vertex, num_vertices;
    float    position[3];
    float    rotation_deg[3];
    float    scale[3];
} rt_object_transform;

/* Indexed triangle mesh exactly as the reference loaders hand it to render():
 * MeshView (GPUandCPU/include/MeshOBJ.h:24-40) + objectMaterials (main.cu:165-190). */
typedef struct rt_scene {
    const float*       positions;      /* [3*num_vertices]                         */
    const float*       normals;        /* [3*num_vertices] or NULL (zero normals)  */
    uint64_t           num_vertices;
    const uint32_t*    indices;        /* [3*num_triangles]                        */
    uint64_t           num_triangles;
    const int32_t*     tri_obj_ids;    /* [num_triangles] or NULL                  */
    const rt_material* materials;      /* [num_materials] or NULL                  */
    int32_t            num_materials;
    uint32_t           build_flags;
    const rt_object_transform* transforms;   /* [num_transforms] or NULL: positions/normals are already in world space */
    int32_t            num_transforms;
} rt_scene;

/* The four vectors Camera::initialize leaves behind
 * (GPUandCPU/include/camera.h:72-94, HW1/include/camera.h:55-92). */
typedef struct rt_camera {
    float center[3];
    float pixel00_loc[3];
    float pixel_delta_u[3];
    float pixel_delta_v[3];
} rt_camera;

typedef struct rt_frame {
    int32_t         mode;          /* RT_MODE_*                                    */
    int32_t         accel;         /* RT_ACCEL_*                                   */
    int32_t         width, height;
    rt_camera       cam;
    const rt_light* lights;
    int32_t         num_lights;
    float           miss_color[3]; /* HW2 constant miss colour (query.h:181-184)   */
    int32_t         spp;           /* samples per pixel (>=1)                      */
    const float*    jitter;        /* [2*spp] sub-pixel offsets; NULL = pixel centre */
    int32_t         max_depth;     /* TraceRayIterative's maxDepth (query.h:156-220): 1 = primary + direct light
                                      (+ shadow ray); > 1 adds mirror / hash-RNG diffuse bounces (HW2_BVH mode) */
    int32_t         shadows;       /* HW2 modes: trace the IsInShadow ray (shader.h:44-62) */
    uint32_t        outputs;       /* RT_OUT_* mask                                */
    int32_t         quantiser;     /* RT_QUANT_*                                   */
    int32_t         kernel_variant;/* 0 = default; >0 selects an experimental traversal kernel */
    int32_t         diffuse_bounce;/* max_depth > 1: settings.diffuse_bounce (query.h:193-209): 1 = pick a diffuse or
                                      a mirror bounce with the per-pixel hash RNG, 0 = mirror bounces only */
    /* RT_MODE_HW2_CPU only — soft shadows of the CPUOnly renderer (CPUOnly/include/raytracer.h:37-46, 121-168): light i is a
     * disk of radius light_radius[i] facing the shaded point, sampled light_shadow_samples[i] times per lit hit
     * (random_in_unit_disk, make_basis); visibility = unoccluded / samples scales the light's contribution.  NULL arrays or a
     * radius of 0 = point light, one shadow ray (the reference's own rule, :128-130).  The reference draws its samples from a
     * process-wide std::mt19937 seeded by std::random_device, in pixel order on one thread, so no two of its runs agree; here the
     * samples come from the reference's OTHER generator, the per-pixel hash RNG of GPUandCPU/include/query.h:32-48, seeded with
     * (x, y, sample) and rng_seed: frames are reproducible and equal the oracle's bit for bit; against the reference the parity is
     * statistical (tests: per-pixel mean and spread of 64 reference runs). */
    const float*    light_radius;          /* [num_lights] or NULL */
    const int32_t*  light_shadow_samples;  /* [num_lights] or NULL (= 1) */
    uint32_t        rng_seed;
} rt_frame;

typedef struct rt_image {
    float*    rgb;        /* [3*W*H] or NULL */
    uint8_t*  rgb8;       /* [3*W*H] or NULL */
    int32_t*  tri_id;     /* [W*H]   or NULL */
    float*    t;          /* [W*H]   or NULL */
    int32_t   width, height;       /* filled */
    uint64_t  rays_primary;        /* filled: closest-hit queries traced (this rank): primary rays + bounce rays */
    uint64_t  rays_shadow;         /* filled: shadow queries traced (this rank)      */
    float     gpu_ms;              /* filled: device time of the last rt_render     */
} rt_image;

typedef struct rt_build_info {
    uint64_t num_triangles;
    uint64_t num_nodes;        /* flattened 64-byte BVH2 nodes                      */
    uint64_t num_leaves;
    uint64_t arena_bytes;      /* nodes + triangle blocks + shading attributes      */
    float    build_ms;         /* device time of the BVH build                      */
    float    upload_ms;        /* host->device copies                               */
    float    scene_min[3], scene_max[3];
} rt_build_info;

/* -- lifetime ------------------------------------------------------------ */
int  rt_api_version(void);
/* Context bound to CUDA device `device` (replaces the cudaMalloc block of
 * GPUandCPU/src/main.cu:199-252). */
int  rt_create(rt_ctx** out, int device);
/* One context over n GPUs of one box driven by ONE host thread of ONE process — what an unmodified single-process
 * main() (GPUandCPU/src/main.cu:98) can use: rt_upload_scene builds on devices[0] and copies the arena to the others
 * over NVLink, rt_render launches every device's share of the frame (screen-space bands, see rt_comm_set_sharding) and
 * the fused peer-store gather delivers it to devices[0], rt_download_image reads it there; rt_render_into makes every
 * device copy its own bands straight into the caller's buffers.  Same kernels and the same in-kernel handshake as the
 * process-per-GPU layout below (rt_comm_init), with peer pointers (cudaDeviceEnablePeerAccess) instead of CUDA IPC and
 * no NCCL.  Needs peer access from every device to devices[0].  rt_image ray counts are totals over the devices. */
int  rt_create_multi(rt_ctx** out, const int* devices, int n);
int  rt_destroy(rt_ctx* ctx);
/* Message of the last failing call on this thread (ctx may be NULL). */
const char* rt_last_error(const rt_ctx* ctx);

/* -- multi-GPU (one process per GPU; screen-space tile sharding) --------- */
/* 128-byte NCCL unique id produced on rank 0 and handed to every rank by the host. */
int  rt_comm_unique_id(void* id128);
int  rt_comm_init(rt_ctx* ctx, int rank, int world, const void* id128);
int  rt_comm_rank(const rt_ctx* ctx, int* rank, int* world);
/* How the ranks' tiles reach rank 0 every frame.  Collective: every rank calls rt_comm_set_gather
 * with the same mode after rt_comm_init (which selects RT_GATHER_AUTO). */
#define RT_GATHER_AUTO  0   /* peer stores when every rank can map rank 0's image (CUDA IPC over NVLink), else NCCL */
#define RT_GATHER_NCCL  1   /* tile-packed planes, one grouped ncclSend/ncclRecv per plane, unpack kernel on rank 0 */
#define RT_GATHER_PEER  2   /* fused: the frame kernel stores its pixels straight into rank 0's row-major image
                               over NVLink; a per-rank flag word in rank 0's memory signals completion */
int  rt_comm_set_gather(rt_ctx* ctx, int mode);
/* Screen-space sharding: the frame's 16x8-pixel tiles (row-major) are cut into world*chunks_per_rank contiguous
 * chunks — horizontal bands — and chunk c is rendered by rank c % world.  0 selects the default (4).  Every rank
 * must use the same value. */
int  rt_comm_set_sharding(rt_ctx* ctx, int chunks_per_rank);
int  rt_comm_gather_mode(const rt_ctx* ctx, int* mode);   /* mode in effect: RT_GATHER_NCCL or RT_GATHER_PEER */
/* Upper bound (seconds, default 120) of every device-side wait on another rank in the fused gather: a rank's frame kernel
 * waiting for rank 0 to release its image (rank 0 publishes that when ITS host calls rt_render for the same frame, so the
 * bound also limits how far apart the ranks' hosts may call rt_render), and rank 0 waiting for the other ranks' bands.
 * A wait that runs out makes the next rt_sync / rt_download_image on that rank fail with RT_ERR_STATE; the late rank
 * stores nothing into rank 0's image.  Not collective. */
int  rt_comm_set_timeout(rt_ctx* ctx, double seconds);

/* -- scene --------------------------------------------------------------- */
/* Pack triangles, build the BVH on the device and (world>1) broadcast the arena
 * from rank 0; ranks != 0 may pass scene == NULL to receive.  Replaces
 * calculateAABBs + buildBVH + buildTrianglesKernel
 * (bvh.cu:60-90, 93-206; main.cu:19-41, 254-293, 347-358). */
int  rt_upload_scene(rt_ctx* ctx, const rt_scene* scene);
int  rt_build_info_get(const rt_ctx* ctx, rt_build_info* info);

/* -- frame --------------------------------------------------------------- */
/* Asynchronous: ray generation + closest hit + shading (+ tile gather to rank 0).
 * Replaces render() (query.cu:79-167) and the HW1 pixel loop (render.cpp:72-116). */
int  rt_render(rt_ctx* ctx, const rt_frame* frame);
/* Blocking: waits for the frame and fills the requested planes (rank 0 holds
 * the gathered image).  Replaces the D2H copy + 8-bit conversion of
 * main.cu:374, 426-431 / render.cpp:119-124. */
int  rt_download_image(rt_ctx* ctx, rt_image* img);
/* rt_render + rt_download_image in one blocking call, with the same result.  On a single-GPU context the frame is
 * rendered in horizontal bands on several streams and every finished band is copied to the caller's buffers while
 * the following bands are still rendering (pinned host memory makes the copies asynchronous), which hides most of
 * the device->host transfer that the reference pays after its kernel (GPUandCPU/src/main.cu:370-378). */
int  rt_render_into(rt_ctx* ctx, const rt_frame* frame, rt_image* img);
/* Multi-GPU contexts: a host buffer shared by every rank's process (POSIX shared memory, page-locked on every rank).
 * Collective; *ptr is this rank's mapping of the same `bytes` bytes.  When EVERY rank calls rt_render_into with plane
 * pointers at the same offsets inside its mapping, each rank copies the bands it rendered straight into the buffer over
 * its own PCIe link while later bands render — no gather to rank 0, no single D2H link (what the reference does after its
 * kernel, GPUandCPU/src/main.cu:374, times the number of GPUs).  rt_render_into returns on rank 0 when every rank's bands
 * have landed; the image is then valid in rank 0's mapping until rank 0 calls rt_render_into again.  After such a frame
 * rt_download_image delivers ray counts and times only.  One image per context; creating another releases the first. */
int  rt_host_image_create(rt_ctx* ctx, size_t bytes, void** ptr);
int  rt_host_image_destroy(rt_ctx* ctx);
/* Blocks until the last rt_render finished; returns its device time. */
int  rt_sync(rt_ctx* ctx, float* gpu_ms);
/* Device times of the last frame on this rank: the whole rt_render (frame kernel + tile delivery to rank 0)
 * and the frame kernel alone.  Blocks like rt_sync. */
int  rt_frame_times(rt_ctx* ctx, float* total_ms, float* kernel_ms);
/* The context's CUDA stream (a cudaStream_t), for hosts that order their own device work or events
 * against rt_render / rt_download_image.  The reference serialises everything on the default stream
 * (CHECK_CUDA, GPUandCPU/include/imports.h:40-47). */
int  rt_stream_handle(const rt_ctx* ctx, void** cuda_stream);
int  rt_device_of(const rt_ctx* ctx, int* device);       /* CUDA device ordinal of the context (rank 0's for rt_create_multi) */

/* Traversal work of the last frame rendered with a *_STATS variant, summed over primary and shadow
 * queries of this rank: per-ray BVH node visits and triangle tests, and the memory requests behind
 * them — 64-byte node lines and 48-byte triangle blocks fetched (one per warp visit in the packet
 * kernel, one per lane in the per-ray kernel).  The roofline's algorithmic bytes per launch are
 * 64*node_lines + 48*tri_blocks + output bytes.  Blocks like rt_sync. */
int  rt_frame_stats(rt_ctx* ctx, uint64_t* node_visits, uint64_t* tri_tests, uint64_t* node_lines, uint64_t* tri_blocks);

/* -- host helpers (pure host arithmetic, no device work) ------------------ */
/* Camera::initialize restated with the reference's mixed fp64/fp32 rounding
 * (GPUandCPU/include/camera.h:72-94; HW1/include/camera.h:55-92 is identical
 * for valid sizes).  Returns RT_ERR_ARG when width/height < 1 (HW1 throws). */
int  rt_camera_init(rt_camera* out, const float pos[3], const float look_at[3],
                    const float up[3], double focal_length_mm,
                    double sensor_height_mm, int width, int height);
/* The CPUOnly renderer's camera (CPUOnly/include/camera.h:64-104): explicit sensor width instead of the
 * aspect-derived one.  Returns RT_ERR_ARG when width/height < 1 (the reference throws). */
int  rt_camera_init_cpuonly(rt_camera* out, const float pos[3], const float look_at[3], const float up[3],
                            double focal_length_mm, double sensor_height_mm, double sensor_width_mm, int width, int height);
/* jittered_samples(spp, seed) of GPUandCPU/include/antialias.h:12-27:
 * std::mt19937 + uniform_real_distribution<float>(0,1) - 0.5.  out[2*spp]. */
int  rt_jitter_table(float* out, int spp, uint32_t seed, int centered);

/* -- host mesh ingest (pure host code) ------------------------------------------ */
/* The reference loaders' output format for callers that do not link C++: OBJ -> unified indexed mesh
 * (vertices de-duplicated by their v/vt/vn reference in order of first use, quads split (0,1,2),(0,2,3),
 * one object id per o/g tag: GPUandCPU/include/MeshOBJ.h:260-427, HW1/src/MeshOBJ.cpp:143-281), the
 * per-object transform of GPUandCPU/src/main.cu:57-96 and AppendMesh (MeshOBJ.h:429-466). */
typedef struct rt_mesh rt_mesh;
int  rt_mesh_load_obj(const char* path, int32_t* next_object_id, rt_mesh** out);
int  rt_mesh_create(rt_mesh** out);                 /* empty mesh, target of rt_mesh_append */
void rt_mesh_free(rt_mesh* mesh);
int  rt_mesh_counts(const rt_mesh* mesh, uint64_t* num_vertices, uint64_t* num_normals, uint64_t* num_triangles);
int  rt_mesh_copy(const rt_mesh* mesh, float* positions, float* normals, uint32_t* indices, int32_t* tri_obj_ids);
int  rt_mesh_transform(rt_mesh* mesh, const float position[3], const float rotation_deg[3], const float scale[3]);
int  rt_mesh_append(rt_mesh* dst, const rt_mesh* src);
const char* rt_mesh_last_error(void);

/* -- OBJ ingest on the device ------------------------------------ */
/* The same loader contract as rt_mesh_load_obj (LoadOBJ_ToMesh, GPUandCPU/include/MeshOBJ.h:260-427), executed on the GPU:
 * `text` (host memory, `nbytes` bytes: the file as it is on disk) is copied to the context's device and parsed there;
 * positions / normals / indices / per-triangle object ids come out as DEVICE arrays.  rt_upload_scene accepts device pointers
 * in rt_scene (cudaMemcpyDefault), so rt_dmesh_arrays feeds it directly, transforms included (rt_scene.transforms are
 * baked on the device) — the mesh never exists in host memory.  Identical, bit for bit, to what rt_mesh_load_obj returns
 * for the same bytes (tests/test_gpu_ingest.py); numbers follow strtof (the few literals fp64 cannot decide are converted by
 * the host's own strtof, rt_dmesh_stats counts them).  Refused with RT_ERR_UNSUPPORTED: inf / nan / hexadecimal literals,
 * lines of 1024 bytes or more (the reference's fgets buffer would split them), files of 2 GiB or more.
 * *next_object_id: in = id of the file's first object, out = first unused id (NULL = start at 0). */
typedef struct rt_dmesh rt_dmesh;
int  rt_dmesh_parse_obj(rt_ctx* ctx, const char* text, uint64_t nbytes, int32_t* next_object_id, rt_dmesh** out);
int  rt_dmesh_create(rt_dmesh** out);               /* empty mesh, target of rt_dmesh_append */
void rt_dmesh_free(rt_dmesh* mesh);
int  rt_dmesh_counts(const rt_dmesh* mesh, uint64_t* num_vertices, uint64_t* num_normals, uint64_t* num_triangles);
int  rt_dmesh_arrays(const rt_dmesh* mesh, const float** positions, const float** normals, const uint32_t** indices, const int32_t** tri_obj_ids);
int  rt_dmesh_copy(const rt_dmesh* mesh, float* positions, float* normals, uint32_t* indices, int32_t* tri_obj_ids);   /* to host memory */
int  rt_dmesh_append(rt_dmesh* dst, const rt_dmesh* src);                     /* AppendMesh (MeshOBJ.h:429-466) on the device */
int  rt_dmesh_stats(const rt_dmesh* mesh, float* parse_ms, uint64_t* lines, uint64_t* host_converted_numbers);
const char* rt_dmesh_last_error(void);

/* -- introspection for tests / profiling ---------------------------------- */
/* Copies the flattened BVH back to the host (nodes: 64 B each; tri blocks: 48 B
 * each, leaf order; tri_ids: original triangle id per block). Any pointer may be
 * NULL; counts come from rt_build_info_get. */
/* Single-GPU contexts only: render just the tiles rank `rank` of a `world`-rank job would own (tile-packed
 * planes, no delivery) so that per-rank load balance and kernel tails can be measured on one GPU.  Planes of
 * such frames cannot be downloaded; ray counts and times can.  world = 0 switches it off. */
int  rt_debug_set_shard(rt_ctx* ctx, int rank, int world);
int  rt_debug_download_bvh(rt_ctx* ctx, void* nodes64, void* tri_blocks48, int32_t* tri_ids);

#ifdef __cplusplus
}
#endif
#endif /* RT_API_H */
