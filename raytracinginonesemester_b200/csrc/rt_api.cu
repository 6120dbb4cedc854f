// rt_api.cu — host side of the C ABI declared in include/rt_api.h.
//
// Owns the context (device, stream, events), the scene arena (64-byte nodes, 48-byte triangle
// geometry/shading blocks, materials), the per-frame output planes and the NCCL communicator.
// There is no CPU rendering path here: without a usable CUDA device every entry point returns
// RT_ERR_CUDA.
#include "rt_kernels.h"
#include "rt_build_core.h"

#include <cuda.h>           // driver API types only; cuStreamWaitValue32 is fetched with cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <nccl.h>   // types and prototypes only: the library is resolved at run time, see NcclApi below

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {
thread_local std::string g_last_error;

// NCCL is bound lazily, on the first multi-GPU call, and to whichever libnccl.so.2 the process
// already has (a host that also runs torch.distributed has torch's bundled NCCL loaded; linking a
// second copy at load time makes the two fight over one SONAME).  Single-GPU contexts never load it.
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
    std::string why;
    NcclApi() {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define RT_SYM(f) f = reinterpret_cast<decltype(f)>(dlsym(h, "nccl" #f)); if (!f) { why = "libnccl.so.2 lacks nccl" #f; return; }
        RT_SYM(GetUniqueId) RT_SYM(CommInitRank) RT_SYM(CommDestroy) RT_SYM(Broadcast) RT_SYM(AllReduce) RT_SYM(Send) RT_SYM(Recv)
        RT_SYM(GroupStart) RT_SYM(GroupEnd) RT_SYM(GetErrorString)
#undef RT_SYM
        ok = true;
    }
};
NcclApi& nccl() { static NcclApi api; return api; }

// Stream-ordered wait on a 32-bit word in device memory (no SM resources, unlike a spinning kernel, so it cannot queue
// behind the frame kernel's pending blocks).  NULL when the driver does not offer it: callers fall back to k_flag_wait.
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValue32Fn stream_wait_value32() {
    static StreamWaitValue32Fn fn = []() -> StreamWaitValue32Fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
        return (StreamWaitValue32Fn)p;
    }();
    return fn;
}

typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWriteValue32Fn stream_write_value32() {
    static StreamWriteValue32Fn fn = []() -> StreamWriteValue32Fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
        return (StreamWriteValue32Fn)p;
    }();
    return fn;
}

struct HostOut { void* host; uint32_t bit; size_t bpp; const void* dev; const char* name; };

struct Plane {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
} // namespace

#define RT_BANDS 8
#define RT_BAND_STREAMS 4
#define RT_SHM_HEADER 4096            // shared host image: header with one 64-byte line per rank (done word), planes behind it
#define RT_COUNTER_BYTES 64           // 8 x u64 ray / traversal counters, followed by the persistent kernel's PersistCtl

struct rt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;   // whole frame / frame kernel only
    std::string err;
    // scene
    WideNode* wide = nullptr;            // 8-wide view of nodes, built after the BVH2 (every rank builds its own)
    BvhNode* nodes = nullptr; TriBlock* geom = nullptr; TriBlock* shade = nullptr; rt_material* materials = nullptr;
    uint32_t num_tris = 0, num_nodes = 0; int num_materials = 0; bool has_scene = false, has_bvh = false, has_normals = false;
    rt_build_info info{};
    // frame
    Plane lights, jitter, counters;
    Plane loc_rgb, loc_rgb8, loc_id, loc_t;       // this rank's planes (row-major if world==1, tile-packed otherwise)
    Plane img_rgb, img_rgb8, img_id, img_t;       // rank 0, world>1: gathered row-major image
    Plane stage_rgb, stage_rgb8, stage_id, stage_t; // rank 0, world>1: receive staging for ranks 1..world-1
    FrameParams fp{};
    uint32_t outputs = 0;
    bool frame_valid = false;
    int launches = 0;
    // comm
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    Plane xfer;                                   // small device staging buffer for host<->NCCL messages
    // peer-store gather (RT_GATHER_PEER): rank 0 exports its image planes and a flag block through CUDA IPC;
    // the other ranks map them and their frame kernels write rank 0's image in place over NVLink.
    bool peer = false;                            // mode in effect
    unsigned* flags = nullptr;                    // rank 0: own allocation; others: IPC mapping of rank 0's
    unsigned* peer_err = nullptr;                 // local: set by a flag wait that timed out
    void* peer_img[4] = {nullptr, nullptr, nullptr, nullptr};   // ranks != 0: mapped rank-0 planes (rgb, rgb8, id, t)
    size_t gather_cap[4] = {0, 0, 0, 0};          // every rank tracks rank 0's plane capacities identically
    std::vector<void*> retired;                   // rank 0: outgrown exported planes, freed at teardown
    unsigned seq = 0;                             // frame sequence number of the peer protocol
    int chunks_per_rank = 0;                      // tile ownership bands per rank (0 = RT_DEFAULT_CHUNKS_PER_RANK)
    // rt_render_into: band-pipelined render + download (kernels of later bands overlap the D2H copy of earlier ones)
    cudaStream_t band_stream[RT_BAND_STREAMS] = {}; cudaStream_t copy_stream = nullptr, sig_stream = nullptr;
    cudaEvent_t band_ev[RT_BANDS] = {}; cudaEvent_t copy_done = nullptr;
    // shared host image (rt_host_image_create): one POSIX shared-memory segment mapped and page-locked by every rank; each
    // rank copies the bands IT rendered straight into it over its own PCIe link (rt_render_into), no gather to rank 0
    char* shm_base = nullptr; size_t shm_bytes = 0;       // whole mapping: RT_SHM_HEADER bytes of per-rank done words, then the planes
    unsigned* hflags = nullptr; unsigned* hflags_dev = nullptr;   // this rank's band flags in mapped pinned HOST memory (+ device alias): rt_render_into polls them
    unsigned hseq = 0;                                   // frame sequence number of the shared-host protocol
    bool last_host_direct = false;                       // the last frame went out through the shared host image
    // rt_create_multi: one process, one host thread, several GPUs.  The context handed to the caller is rank 0 and owns the others;
    // peers are reached through unified addressing (cudaDeviceEnablePeerAccess) instead of IPC mappings, nothing goes through NCCL.
    std::vector<rt_ctx*> members;                        // ranks 1..n-1 (rank 0 of a group only)
    bool inproc = false;                                 // this context is a rank of such a group
    rt_ctx* group_root = nullptr;                        // ... and this is its rank 0
    unsigned long long* pin_dev = nullptr;               // device alias of pin; pin[8..10]: ray counters + sequence number written by the persistent kernel's last block,
                                                         // pin[12]: 'copies done' word written by a stream memory operation (host-polled end of rt_render_into)
    unsigned long long* pin = nullptr;                   // 128 bytes of page-locked host memory: small device->host reads (ray counters, error word)
                                                         // land here — into pageable memory each of them is a staged, synchronous copy of tens of microseconds
    bool fast_finish = false;                            // the rt_render_into in flight ends with host-polled words (no stream synchronisation)
    HostOut pump_outs[4] = {}; unsigned pump_seq = 0; int pump_next = 0; bool last_pumped = false;   // band copies of the rt_render_into in flight (band_pump)
    unsigned long long peer_timeout_ns = 120ull * 1000000000ull;   // bound of every in-kernel / flag-kernel wait on another rank (rt_comm_set_timeout)
    int dbg_rank = 0, dbg_world = 0;              // rt_debug_set_shard: render one rank's share on a single GPU (timing only)
};

namespace {

int fail(rt_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_last_error = buf;
    if (c) c->err = buf;
    return code;
}
#define CU(c, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail((c), RT_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)
#define NC(c, x) do { if (!nccl().ok) return fail((c), RT_ERR_NCCL, "%s", nccl().why.c_str()); ncclResult_t r_ = nccl().x; if (r_ != ncclSuccess) return fail((c), RT_ERR_NCCL, "nccl%s: %s", #x, nccl().GetErrorString(r_)); } while (0)

void free_scene(rt_ctx* c) {
    if (c->nodes) cudaFree(c->nodes);
    if (c->wide) cudaFree(c->wide);
    c->wide = nullptr;
    if (c->geom) cudaFree(c->geom);
    if (c->shade) cudaFree(c->shade);
    if (c->materials) cudaFree(c->materials);
    c->nodes = nullptr; c->geom = nullptr; c->shade = nullptr; c->materials = nullptr;
    c->num_tris = c->num_nodes = 0; c->num_materials = 0; c->has_scene = c->has_bvh = false;
}


// Device bytes of the scene arrays (the 8-wide view adds its own in build_wide_nodes).  The per-vertex-normal blocks exist
// only when the mesh has normals.
static uint64_t scene_arena_bytes(const rt_ctx* c) {
    return sizeof(BvhNode) * (uint64_t)c->num_nodes + (c->has_normals ? 2u : 1u) * sizeof(TriBlock) * (uint64_t)c->num_tris +
           sizeof(rt_material) * (uint64_t)c->num_materials;
}

// 8-wide view of the BVH2 for the frustum traversal: derived data, rebuilt by every rank from its copy of the nodes.
int build_wide_nodes(rt_ctx* c) {
    if (c->wide) { cudaFree(c->wide); c->wide = nullptr; }
    if (!c->has_bvh || !c->num_nodes) return RT_OK;
    cudaEvent_t e0, e1;
    CU(c, cudaEventCreate(&e0)); CU(c, cudaEventCreate(&e1));
    CU(c, cudaEventRecord(e0, c->stream));
    uint32_t count = 0;
    // Phase 0 (wide nodes at depth 0, 3, 6, ...): on the LBVHs measured the three thirds are the same size to 2 % and the
    // full first hop renders 1.5 % faster (tools/wide_probe.py).  RT_B200_WIDE_PHASE = 1, 2 or -1 (smallest) is a measurement knob.
    const char* ph = getenv("RT_B200_WIDE_PHASE");
    const cudaError_t we = rt_build_wide(c->nodes, c->num_nodes, ph && *ph ? atoi(ph) : 0, &c->wide, &count, c->stream);
    if (we == cudaErrorMemoryAllocation) {
        (void)cudaGetLastError();          // not enough memory for the 8-wide view: frames use the per-lane traversal
        c->wide = nullptr;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return RT_OK;
    }
    CU(c, we);
    CU(c, cudaEventRecord(e1, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c->info.build_ms += ms;
    c->info.arena_bytes += sizeof(WideNode) * (uint64_t)count;
    return RT_OK;
}

struct SceneHeader { int32_t status; uint32_t num_tris, num_nodes, num_leaves, num_materials, has_bvh, has_normals; float smin[3], smax[3]; };

// Broadcast of the built arena from rank 0 (scene + BVH travel once per scene over NVLink).  Collective: rank 0 ALWAYS
// gets here, also when its own upload failed (root_status != RT_OK) — the header carries the status, so every rank
// returns an error together instead of the others waiting forever in ncclBroadcast.
int broadcast_scene(rt_ctx* c, int root_status) {
    SceneHeader h{};
    if (c->rank == 0) {
        h.status = root_status;
        h.num_tris = c->num_tris; h.num_nodes = c->num_nodes; h.num_leaves = (uint32_t)c->info.num_leaves; h.num_materials = (uint32_t)c->num_materials; h.has_bvh = c->has_bvh; h.has_normals = c->has_normals;
        memcpy(h.smin, c->info.scene_min, sizeof h.smin); memcpy(h.smax, c->info.scene_max, sizeof h.smax);
    }
    SceneHeader* dh = nullptr;
    CU(c, cudaMalloc(&dh, sizeof h));
    if (c->rank == 0) CU(c, cudaMemcpyAsync(dh, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    NC(c, Broadcast(dh, dh, sizeof h, ncclUint8, 0, c->comm, c->stream));
    CU(c, cudaMemcpyAsync(&h, dh, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(dh);
    if (h.status != RT_OK) {
        if (c->rank == 0) return root_status;                  // (its message is already set)
        free_scene(c);
        return fail(c, RT_ERR_STATE, "rt_upload_scene: rank 0 failed to upload the scene (status %d); no scene on this rank", h.status);
    }
    if (c->rank != 0) {
        free_scene(c);
        c->num_tris = h.num_tris; c->num_nodes = h.num_nodes; c->num_materials = (int)h.num_materials; c->has_bvh = h.has_bvh != 0; c->has_normals = h.has_normals != 0;
        if (c->num_nodes) CU(c, cudaMalloc(&c->nodes, sizeof(BvhNode) * (size_t)c->num_nodes));
        CU(c, cudaMalloc(&c->geom, sizeof(TriBlock) * (size_t)c->num_tris));
        if (c->has_normals) CU(c, cudaMalloc(&c->shade, sizeof(TriBlock) * (size_t)c->num_tris));
        if (c->num_materials) CU(c, cudaMalloc(&c->materials, sizeof(rt_material) * (size_t)c->num_materials));
        memset(&c->info, 0, sizeof c->info);
        c->info.num_triangles = h.num_tris; c->info.num_nodes = h.num_nodes; c->info.num_leaves = h.num_leaves;
        memcpy(c->info.scene_min, h.smin, sizeof h.smin); memcpy(c->info.scene_max, h.smax, sizeof h.smax);
    }
    NC(c, GroupStart());
    if (c->num_nodes) NC(c, Broadcast(c->nodes, c->nodes, sizeof(BvhNode) * (size_t)c->num_nodes, ncclUint8, 0, c->comm, c->stream));
    NC(c, Broadcast(c->geom, c->geom, sizeof(TriBlock) * (size_t)c->num_tris, ncclUint8, 0, c->comm, c->stream));
    if (c->has_normals) NC(c, Broadcast(c->shade, c->shade, sizeof(TriBlock) * (size_t)c->num_tris, ncclUint8, 0, c->comm, c->stream));
    if (c->num_materials) NC(c, Broadcast(c->materials, c->materials, sizeof(rt_material) * (size_t)c->num_materials, ncclUint8, 0, c->comm, c->stream));
    NC(c, GroupEnd());
    CU(c, cudaStreamSynchronize(c->stream));
    c->has_scene = true;
    c->info.arena_bytes = scene_arena_bytes(c);
    return build_wide_nodes(c);
}

// ---- small host-side collectives over the library's communicator (setup paths only) ----
int bcast_bytes(rt_ctx* c, void* host, size_t n) {           // rank 0's bytes -> every rank's `host`
    CU(c, c->xfer.reserve(n < 256 ? 256 : n));
    if (c->rank == 0) CU(c, cudaMemcpyAsync(c->xfer.p, host, n, cudaMemcpyHostToDevice, c->stream));
    NC(c, Broadcast(c->xfer.p, c->xfer.p, n, ncclUint8, 0, c->comm, c->stream));
    if (c->rank != 0) CU(c, cudaMemcpyAsync(host, c->xfer.p, n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RT_OK;
}
int all_min(rt_ctx* c, int v, int* out) {                     // also a true barrier
    CU(c, c->xfer.reserve(256));
    CU(c, cudaMemcpyAsync(c->xfer.p, &v, sizeof v, cudaMemcpyHostToDevice, c->stream));
    NC(c, AllReduce(c->xfer.p, c->xfer.p, 1, ncclInt32, ncclMin, c->comm, c->stream));
    CU(c, cudaMemcpyAsync(out, c->xfer.p, sizeof v, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RT_OK;
}

const size_t kFlagBytes = sizeof(unsigned) * RT_PEER_CHUNK_FLAG(RT_PEER_MAX_RANKS, 0);   // layout: rt_kernels.h

// Drops every peer mapping / exported allocation.  `collective`: all ranks are here and the communicator is
// healthy, so rank 0 can wait for the importers to unmap before it frees (otherwise the exported memory is
// left to process teardown: freeing it under a live import is undefined).
void peer_teardown(rt_ctx* c, bool collective) {
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->rank != 0) {
        for (void*& p : c->peer_img) { if (p) cudaIpcCloseMemHandle(p); p = nullptr; }
        if (c->flags) cudaIpcCloseMemHandle(c->flags);
        c->flags = nullptr;
    }
    bool safe = c->world == 1;
    if (c->world > 1 && collective && c->comm && nccl().ok) { int dummy = 0; safe = all_min(c, 1, &dummy) == RT_OK; }
    if (c->rank == 0) {
        Plane* planes[4] = {&c->img_rgb, &c->img_rgb8, &c->img_id, &c->img_t};
        for (int i = 0; i < 4; ++i) {
            if (!c->gather_cap[i]) continue;                    // exported plane
            if (safe) planes[i]->release(); else { planes[i]->p = nullptr; planes[i]->cap = 0; }   // else leaked on purpose
        }
        if (safe) { for (void* p : c->retired) cudaFree(p); if (c->flags) cudaFree(c->flags); }
        c->retired.clear();
        c->flags = nullptr;
    }
    for (size_t& g : c->gather_cap) g = 0;
    c->peer = false;
}

void shm_release(rt_ctx* c) {
    if (c->shm_base) { cudaHostUnregister(c->shm_base); munmap(c->shm_base, c->shm_bytes); }
    c->shm_base = nullptr; c->shm_bytes = 0;
}

int peer_timeout(rt_ctx* c, unsigned code) {
    cudaMemset(c->peer_err, 0, sizeof(unsigned));
    c->frame_valid = false;
    return fail(c, RT_ERR_STATE, "peer-store gather: rank %d waited more than %llu s for %s", c->rank, c->peer_timeout_ns / 1000000000ull,
                c->rank == 0 ? "another rank's completion flag" : "rank 0 to start the frame");
    (void)code;
}

// Collective.  Tries to put the peer-store gather in place: rank 0 exports the flag block, everybody maps it.
int peer_setup(rt_ctx* c) {
    c->peer = false;
    if (c->world <= 1) return RT_OK;
    if (!c->peer_err) { CU(c, cudaMalloc(&c->peer_err, sizeof(unsigned))); CU(c, cudaMemset(c->peer_err, 0, sizeof(unsigned))); }
    cudaIpcMemHandle_t h; memset(&h, 0, sizeof h);
    int ok = 1;
    if (c->rank == 0) {
        if (cudaMalloc(&c->flags, kFlagBytes) != cudaSuccess || cudaMemset(c->flags, 0, kFlagBytes) != cudaSuccess ||
            cudaIpcGetMemHandle(&h, c->flags) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    }
    int rc = bcast_bytes(c, &h, sizeof h);
    if (rc != RT_OK) return rc;
    if (c->rank != 0) {
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
        c->flags = (unsigned*)p;
    }
    int all = 0;
    rc = all_min(c, ok, &all);
    if (rc != RT_OK) return rc;
    c->seq = 0;
    if (!all) {
        if (c->rank != 0 && c->flags) cudaIpcCloseMemHandle(c->flags);
        if (c->rank == 0 && c->flags) c->retired.push_back(c->flags);
        c->flags = nullptr;
        return RT_OK;                                           // stays on the NCCL gather
    }
    c->peer = true;
    return RT_OK;
}

// Collective (every rank evaluates the same growth condition): rank 0 (re)allocates the image planes the
// frame needs and exports them; the other ranks map them.  On any failure all ranks fall back to NCCL.
int peer_planes(rt_ctx* c, uint32_t outputs, size_t npix) {
    static const uint32_t bits[4] = {RT_OUT_RGB_F32, RT_OUT_RGB8, RT_OUT_TRI_ID, RT_OUT_T};
    static const size_t bpp[4] = {12, 3, 4, 4};
    Plane* planes[4] = {&c->img_rgb, &c->img_rgb8, &c->img_id, &c->img_t};
    struct Msg { cudaIpcMemHandle_t h[4]; int ok; } m;
    memset(&m, 0, sizeof m); m.ok = 1;
    bool grow[4]; bool any = false;
    for (int i = 0; i < 4; ++i) { grow[i] = (outputs & bits[i]) && bpp[i] * npix + 16 > c->gather_cap[i]; any = any || grow[i]; }
    if (!any) return RT_OK;
    CU(c, cudaStreamSynchronize(c->stream));                    // nobody is still writing the old planes
    for (int i = 0; i < 4; ++i) {
        if (!grow[i]) continue;
        const size_t need = bpp[i] * npix + 16;
        if (c->rank == 0) {
            if (planes[i]->p) { c->retired.push_back(planes[i]->p); planes[i]->p = nullptr; planes[i]->cap = 0; }
            if (cudaMalloc(&planes[i]->p, need) != cudaSuccess) { m.ok = 0; cudaGetLastError(); planes[i]->p = nullptr; continue; }
            planes[i]->cap = need;
            if (cudaIpcGetMemHandle(&m.h[i], planes[i]->p) != cudaSuccess) { m.ok = 0; cudaGetLastError(); }
        } else if (c->peer_img[i]) { cudaIpcCloseMemHandle(c->peer_img[i]); c->peer_img[i] = nullptr; }
    }
    int rc = bcast_bytes(c, &m, sizeof m);
    if (rc != RT_OK) return rc;
    int ok = m.ok;
    if (c->rank != 0 && ok) {
        for (int i = 0; i < 4; ++i) {
            if (!grow[i]) continue;
            if (cudaIpcOpenMemHandle(&c->peer_img[i], m.h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); c->peer_img[i] = nullptr; }
        }
    }
    int all = 0;
    rc = all_min(c, ok, &all);
    if (rc != RT_OK) return rc;
    if (!all) { peer_teardown(c, true); return RT_OK; }
    for (int i = 0; i < 4; ++i) if (grow[i]) c->gather_cap[i] = bpp[i] * npix + 16;
    return RT_OK;
}

// In-process group (rt_create_multi): rank 0 (re)allocates the planes, the others use the same pointers.  Called on every
// rank of the group with the same arguments, rank 0 first.
int inproc_planes(rt_ctx* c, rt_ctx* root, uint32_t outputs, size_t npix) {
    static const uint32_t bits[4] = {RT_OUT_RGB_F32, RT_OUT_RGB8, RT_OUT_TRI_ID, RT_OUT_T};
    static const size_t bpp[4] = {12, 3, 4, 4};
    Plane* planes[4] = {&root->img_rgb, &root->img_rgb8, &root->img_id, &root->img_t};
    for (int i = 0; i < 4; ++i) {
        if (!(outputs & bits[i])) continue;
        if (c == root) {
            if (bpp[i] * npix + 16 > planes[i]->cap) {
                for (rt_ctx* m : root->members) { cudaSetDevice(m->device); cudaStreamSynchronize(m->stream); }   // nobody still writes the old plane
                CU(c, cudaSetDevice(c->device));
                CU(c, cudaStreamSynchronize(c->stream));
                CU(c, planes[i]->reserve(bpp[i] * npix + 16));
            }
        } else c->peer_img[i] = planes[i]->p;
    }
    return RT_OK;
}

} // namespace

extern "C" {

int rt_api_version(void) { return RT_API_VERSION; }

const char* rt_last_error(const rt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int rt_create(rt_ctx** out, int device) {
    if (!out) return fail(nullptr, RT_ERR_ARG, "rt_create: out is NULL");
    *out = nullptr;
    int count = 0;
    CU(nullptr, cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(nullptr, RT_ERR_ARG, "rt_create: device %d out of range (%d visible)", device, count);
    CU(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(nullptr, RT_ERR_CUDA, "rt_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    {   // The BVH build takes its scratch (~60 B per triangle) from the stream-ordered pool; with the default release
        // threshold of 0 the pool hands the memory back at every synchronisation and each build pays the physical
        // allocation again (C5: 56 ms instead of 17 ms).  Keep it.
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    rt_ctx* c = new (std::nothrow) rt_ctx;
    if (!c) return fail(nullptr, RT_ERR_NOMEM, "rt_create: out of host memory");
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&c->evk0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->evk1);
    if (e == cudaSuccess) e = c->counters.reserve(RT_COUNTER_BYTES + RT_PERSIST_CTL_BYTES);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->pin, 128, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e == cudaSuccess) { memset(c->pin, 0, 128); e = cudaHostGetDevicePointer((void**)&c->pin_dev, c->pin, 0); }
    if (e != cudaSuccess) { int r = fail(nullptr, RT_ERR_CUDA, "rt_create: %s", cudaGetErrorString(e)); delete c; return r; }
    *out = c;
    return RT_OK;
}

int rt_create_multi(rt_ctx** out, const int* devices, int n) {
    if (!out || !devices || n < 1 || n > RT_PEER_MAX_RANKS) return fail(nullptr, RT_ERR_ARG, "rt_create_multi: 1..%d devices", RT_PEER_MAX_RANKS);
    *out = nullptr;
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j)
        if (devices[i] == devices[j]) return fail(nullptr, RT_ERR_ARG, "rt_create_multi: device %d listed twice", devices[i]);
    rt_ctx* root = nullptr;
    int rc = rt_create(&root, devices[0]);
    if (rc != RT_OK || n == 1) { *out = root; return rc; }
    auto bail = [&](int code) { rt_destroy(root); return code; };
    root->rank = 0; root->world = n; root->inproc = true; root->group_root = root;
    for (int i = 1; i < n; ++i) {
        rt_ctx* m = nullptr;
        rc = rt_create(&m, devices[i]);
        if (rc != RT_OK) return bail(rc);
        m->rank = i; m->world = n; m->inproc = true; m->group_root = root;
        root->members.push_back(m);
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) != cudaSuccess || !can)
            return bail(fail(nullptr, RT_ERR_CUDA, "rt_create_multi: device %d cannot address device %d's memory (no NVLink / P2P path)", devices[i], devices[0]));
        cudaSetDevice(devices[i]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bail(fail(nullptr, RT_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)));
        cudaGetLastError();
        cudaSetDevice(devices[0]);
        e = cudaDeviceEnablePeerAccess(devices[i], 0);           // rank 0 -> rank i: the scene copy
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bail(fail(nullptr, RT_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)));
        cudaGetLastError();
    }
    // the fused gather's flag block lives on rank 0; every rank addresses it directly
    cudaSetDevice(devices[0]);
    if (cudaMalloc(&root->flags, kFlagBytes) != cudaSuccess || cudaMemset(root->flags, 0, kFlagBytes) != cudaSuccess)
        return bail(fail(nullptr, RT_ERR_CUDA, "rt_create_multi: flag block"));
    std::vector<rt_ctx*> all{root};
    all.insert(all.end(), root->members.begin(), root->members.end());
    for (rt_ctx* m : all) {
        cudaSetDevice(m->device);
        if (cudaMalloc(&m->peer_err, sizeof(unsigned)) != cudaSuccess || cudaMemset(m->peer_err, 0, sizeof(unsigned)) != cudaSuccess)
            return bail(fail(nullptr, RT_ERR_CUDA, "rt_create_multi: error word"));
        m->flags = root->flags; m->peer = true; m->seq = 0;
    }
    cudaSetDevice(devices[0]);
    *out = root;
    return RT_OK;
}

int rt_destroy(rt_ctx* c) {
    if (!c) return RT_OK;
    if (c->inproc && c->group_root == c) {                       // a group: the other ranks first (they address rank 0's memory)
        for (rt_ctx* m : c->members) { cudaSetDevice(m->device); if (m->stream) cudaStreamSynchronize(m->stream); }
        for (rt_ctx* m : c->members) { m->flags = nullptr; for (void*& p : m->peer_img) p = nullptr; m->peer = false; m->inproc = false; rt_destroy(m); }
        c->members.clear();
        cudaSetDevice(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        if (c->flags) cudaFree(c->flags);
        c->flags = nullptr; c->peer = false; c->inproc = false; c->world = 1;
        for (size_t& g : c->gather_cap) g = 0;
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    // NOT a collective (ranks may destroy in any order, or after another rank died): mappings are closed, rank 0's
    // exported planes are left to process teardown because an importer may still have them mapped.
    peer_teardown(c, false);
    shm_release(c);
    if (c->hflags) cudaFreeHost(c->hflags);
    if (c->pin) cudaFreeHost(c->pin);
    if (c->peer_err) cudaFree(c->peer_err);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    free_scene(c);
    c->xfer.release();
    Plane* planes[] = {&c->lights, &c->jitter, &c->counters, &c->loc_rgb, &c->loc_rgb8, &c->loc_id, &c->loc_t,
                       &c->img_rgb, &c->img_rgb8, &c->img_id, &c->img_t, &c->stage_rgb, &c->stage_rgb8, &c->stage_id, &c->stage_t};
    for (Plane* p : planes) p->release();
    for (cudaStream_t s : c->band_stream) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->sig_stream) { cudaStreamSynchronize(c->sig_stream); cudaStreamDestroy(c->sig_stream); }
    for (cudaEvent_t e : c->band_ev) if (e) cudaEventDestroy(e);
    if (c->copy_done) cudaEventDestroy(c->copy_done);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evk0) cudaEventDestroy(c->evk0);
    if (c->evk1) cudaEventDestroy(c->evk1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return RT_OK;
}

int rt_comm_unique_id(void* id128) {
    if (!id128) return fail(nullptr, RT_ERR_ARG, "rt_comm_unique_id: NULL");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(nullptr, GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return RT_OK;
}

int rt_comm_init(rt_ctx* c, int rank, int world, const void* id128) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(c, RT_ERR_ARG, "rt_comm_init: bad arguments");
    if (c->inproc) return fail(c, RT_ERR_STATE, "rt_comm_init: this context is a single-process group (rt_create_multi)");
    CU(c, cudaSetDevice(c->device));
    if (world > 8) return fail(c, RT_ERR_ARG, "rt_comm_init: at most 8 ranks (one box)");
    peer_teardown(c, true);
    if (c->comm && nccl().ok) { nccl().CommDestroy(c->comm); c->comm = nullptr; }
    c->rank = rank; c->world = world;
    c->frame_valid = false;
    if (world > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        NC(c, CommInitRank(&c->comm, world, id, rank));
        return peer_setup(c);                                  // RT_GATHER_AUTO
    }
    return RT_OK;
}

int rt_comm_set_gather(rt_ctx* c, int mode) {
    if (!c || mode < RT_GATHER_AUTO || mode > RT_GATHER_PEER) return fail(c, RT_ERR_ARG, "rt_comm_set_gather: bad arguments");
    CU(c, cudaSetDevice(c->device));
    c->frame_valid = false;
    if (c->world <= 1) return RT_OK;
    if (c->inproc) return mode == RT_GATHER_NCCL ? fail(c, RT_ERR_ARG, "rt_comm_set_gather: a single-process group gathers with peer stores only") : RT_OK;
    if (mode == RT_GATHER_NCCL) { peer_teardown(c, true); return RT_OK; }
    if (!c->peer) { int rc = peer_setup(c); if (rc != RT_OK) return rc; }
    if (mode == RT_GATHER_PEER && !c->peer) return fail(c, RT_ERR_CUDA, "rt_comm_set_gather: CUDA IPC peer mapping of rank 0's memory is not available on every rank");
    return RT_OK;
}

int rt_comm_set_sharding(rt_ctx* c, int chunks_per_rank) {
    if (!c || chunks_per_rank < 0 || chunks_per_rank > 4096) return fail(c, RT_ERR_ARG, "rt_comm_set_sharding: bad arguments");
    c->chunks_per_rank = chunks_per_rank;
    c->frame_valid = false;
    for (rt_ctx* m : c->members) { m->chunks_per_rank = chunks_per_rank; m->frame_valid = false; }
    return RT_OK;
}

int rt_comm_set_timeout(rt_ctx* c, double seconds) {
    if (!c || !(seconds > 0.0) || seconds > 86400.0) return fail(c, RT_ERR_ARG, "rt_comm_set_timeout: seconds must be in (0, 86400]");
    c->peer_timeout_ns = (unsigned long long)(seconds * 1e9);
    return RT_OK;
}

int rt_comm_gather_mode(const rt_ctx* c, int* mode) {
    if (!c || !mode) return fail(nullptr, RT_ERR_ARG, "rt_comm_gather_mode: NULL");
    *mode = c->peer ? RT_GATHER_PEER : RT_GATHER_NCCL;
    return RT_OK;
}

int rt_comm_rank(const rt_ctx* c, int* rank, int* world) {
    if (!c) return fail(nullptr, RT_ERR_ARG, "rt_comm_rank: NULL ctx");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return RT_OK;
}

// Rank 0 / single GPU: validate, copy, bake transforms, build.  No collective step in here (see rt_upload_scene).
static int upload_local(rt_ctx* c, const rt_scene* sc) {
    if (!sc->positions || !sc->indices || sc->num_triangles == 0 || sc->num_vertices == 0)
        return fail(c, RT_ERR_ARG, "rt_upload_scene: empty mesh (positions/indices NULL or zero counts)");
    if (sc->num_triangles >= (1ull << 28)) return fail(c, RT_ERR_ARG, "rt_upload_scene: more than 2^28 triangles");
    if (sc->num_materials < 0 || (sc->num_materials > 0 && !sc->materials)) return fail(c, RT_ERR_ARG, "rt_upload_scene: materials");
    const bool timing = getenv("RT_TIMING") != nullptr;        // host-side phase times on stderr (diagnostic)
    auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[rt_upload_scene] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
        tp0 = t;
    };
    if (sc->num_transforms < 0 || (sc->num_transforms > 0 && !sc->transforms)) return fail(c, RT_ERR_ARG, "rt_upload_scene: transforms");
    for (int k = 0; k < sc->num_transforms; ++k)
        if (sc->transforms[k].first_vertex > sc->num_vertices || sc->transforms[k].num_vertices > sc->num_vertices - sc->transforms[k].first_vertex)
            return fail(c, RT_ERR_ARG, "rt_upload_scene: transform %d covers vertices outside the mesh", k);
    if (sc->num_vertices >= (1ull << 32)) return fail(c, RT_ERR_ARG, "rt_upload_scene: more than 2^32 vertices");
    free_scene(c);
    lap("free previous scene");

    const size_t nv = (size_t)sc->num_vertices, nt = (size_t)sc->num_triangles;
    float *d_pos = nullptr, *d_nrm = nullptr; uint32_t* d_idx = nullptr; int32_t* d_obj = nullptr;
    cudaEvent_t e0, e1, e2;
    CU(c, cudaEventCreate(&e0)); CU(c, cudaEventCreate(&e1)); CU(c, cudaEventCreate(&e2));
    int rc = RT_OK;
    auto cleanup = [&]() {
        if (d_pos) cudaFree(d_pos); if (d_nrm) cudaFree(d_nrm); if (d_idx) cudaFree(d_idx); if (d_obj) cudaFree(d_obj);
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    };
#define CUS(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(c, RT_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); cleanup(); free_scene(c); return rc; } } while (0)
    CUS(cudaMalloc(&d_pos, sizeof(float) * 3 * nv));
    CUS(cudaMalloc(&d_idx, sizeof(uint32_t) * 3 * nt));
    if (sc->normals) CUS(cudaMalloc(&d_nrm, sizeof(float) * 3 * nv));
    if (sc->tri_obj_ids) CUS(cudaMalloc(&d_obj, sizeof(int32_t) * nt));
    CUS(cudaMalloc(&c->geom, sizeof(TriBlock) * nt));
    if (sc->normals) CUS(cudaMalloc(&c->shade, sizeof(TriBlock) * nt));      // normals blocks: only for meshes that have normals
    if (sc->num_materials) CUS(cudaMalloc(&c->materials, sizeof(rt_material) * (size_t)sc->num_materials));
    lap("cudaMalloc");
    CUS(cudaEventRecord(e0, c->stream));
    CUS(cudaMemcpyAsync(d_pos, sc->positions, sizeof(float) * 3 * nv, cudaMemcpyDefault, c->stream));
    CUS(cudaMemcpyAsync(d_idx, sc->indices, sizeof(uint32_t) * 3 * nt, cudaMemcpyDefault, c->stream));
    if (d_nrm) CUS(cudaMemcpyAsync(d_nrm, sc->normals, sizeof(float) * 3 * nv, cudaMemcpyDefault, c->stream));
    if (d_obj) CUS(cudaMemcpyAsync(d_obj, sc->tri_obj_ids, sizeof(int32_t) * nt, cudaMemcpyDefault, c->stream));
    if (sc->num_materials) CUS(cudaMemcpyAsync(c->materials, sc->materials, sizeof(rt_material) * (size_t)sc->num_materials, cudaMemcpyHostToDevice, c->stream));
    for (int k = 0; k < sc->num_transforms; ++k) {             // applyObjectTransform on the device (main.cu:75-96)
        const rt_object_transform& o = sc->transforms[k];
        const float d2r = 0.01745329251994329577f;             // deg2rad, main.cu:52-54
        const float rx = o.rotation_deg[0] * d2r, ry = o.rotation_deg[1] * d2r, rz = o.rotation_deg[2] * d2r;
        BakeXform T;
        T.sx = o.scale[0]; T.sy = o.scale[1]; T.sz = o.scale[2];
        T.cx_ = cosf(rx); T.sx_ = sinf(rx); T.cy_ = cosf(ry); T.sy_ = sinf(ry); T.cz_ = cosf(rz); T.sz_ = sinf(rz);
        T.tx = o.position[0]; T.ty = o.position[1]; T.tz = o.position[2];
        CUS(rt_bake_transform(d_pos, d_nrm, (size_t)o.first_vertex, (size_t)o.num_vertices, T, c->stream));
    }
    {   // index range check on the device, on the copy that the build will read (the reference's loaders trust the file,
        // MeshOBJ.h:389-401; a host loop over 30 M indices of C5 took longer than the whole build)
        unsigned long long bad = ~0ull, *d_bad = nullptr;
        CUS(cudaMallocAsync(&d_bad, sizeof bad, c->stream));
        CUS(cudaMemcpyAsync(d_bad, &bad, sizeof bad, cudaMemcpyHostToDevice, c->stream));
        CUS(rt_validate_indices(d_idx, 3 * nt, (uint32_t)nv, d_bad, c->stream));
        CUS(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
        CUS(cudaFreeAsync(d_bad, c->stream));
        CUS(cudaStreamSynchronize(c->stream));
        if (bad != ~0ull) {
            cleanup(); free_scene(c);
            return fail(c, RT_ERR_ARG, "rt_upload_scene: index %llu out of range", bad);
        }
    }
    CUS(cudaEventRecord(e1, c->stream));
    lap("H2D copies + index check");

    BuildParams bp{};
    bp.positions = d_pos; bp.normals = d_nrm; bp.indices = d_idx; bp.obj_ids = d_obj;
    bp.num_tris = (uint32_t)nt;
    uint32_t leaf_max = (sc->build_flags >> 8) & 0xFu;     // bits 8..11: leaf size override (0 = default)
    bp.leaf_max = leaf_max ? leaf_max : 2u;
    BuildResult br{};
    memset(&c->info, 0, sizeof c->info);
    if (sc->build_flags & RT_BUILD_NO_BVH) {
        CUS(rt_pack_triangles(bp, c->geom, c->shade, c->stream));
        c->has_bvh = false; c->num_nodes = 0;
    } else {
        CUS(rt_build_bvh(bp, &c->nodes, c->geom, c->shade, &br, c->stream));
        c->has_bvh = true; c->num_nodes = br.num_nodes; c->info.num_leaves = br.num_leaves;
        memcpy(c->info.scene_min, br.scene_min, sizeof br.scene_min);
        memcpy(c->info.scene_max, br.scene_max, sizeof br.scene_max);
    }
    CUS(cudaEventRecord(e2, c->stream));
    CUS(cudaStreamSynchronize(c->stream));
    lap("build (host view)");
    float up_ms = 0.f, b_ms = 0.f;
    cudaEventElapsedTime(&up_ms, e0, e1);
    cudaEventElapsedTime(&b_ms, e1, e2);
    cleanup();
    lap("free staging");
#undef CUS
    c->num_tris = (uint32_t)nt; c->num_materials = sc->num_materials; c->has_scene = true; c->has_normals = sc->normals != nullptr;
    c->info.num_triangles = nt; c->info.num_nodes = c->num_nodes; c->info.build_ms = b_ms; c->info.upload_ms = up_ms;
    c->info.arena_bytes = scene_arena_bytes(c);
    return RT_OK;
}

int rt_host_image_create(rt_ctx* c, size_t bytes, void** out) {
    if (!c || !out || bytes == 0) return fail(c, RT_ERR_ARG, "rt_host_image_create: bad arguments");
    *out = nullptr;
    CU(c, cudaSetDevice(c->device));
    shm_release(c);
    const size_t total = RT_SHM_HEADER + ((bytes + 4095) & ~(size_t)4095);
    char name[64]; memset(name, 0, sizeof name);
    int ok = 1, fd = -1;
    if (c->rank == 0) {
        static int serial = 0;
        snprintf(name, sizeof name, "/rt_b200_%d_%d", (int)getpid(), ++serial);
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)total) != 0) ok = 0;
    }
    if (c->world > 1) {                                         // collective: the name travels over the library's communicator
        int rc = bcast_bytes(c, name, sizeof name);
        if (rc != RT_OK) { if (fd >= 0) { close(fd); shm_unlink(name); } return rc; }
        if (c->rank != 0) { fd = shm_open(name, O_RDWR, 0600); if (fd < 0) ok = 0; }
    }
    void* p = MAP_FAILED;
    if (ok) p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (fd >= 0) close(fd);
    if (p == MAP_FAILED) ok = 0;
    if (ok && c->rank == 0) memset(p, 0, RT_SHM_HEADER);
    if (ok && cudaHostRegister(p, total, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) { cudaGetLastError(); munmap(p, total); p = MAP_FAILED; ok = 0; }
    int all = ok;
    if (c->world > 1) { int rc = all_min(c, ok, &all); if (rc != RT_OK) all = 0; }      // also the barrier before the name is unlinked
    if (c->rank == 0 && name[0]) shm_unlink(name);              // the mappings keep the segment alive
    if (!all) {
        if (p != MAP_FAILED) { cudaHostUnregister(p); munmap(p, total); }
        return fail(c, RT_ERR_NOMEM, "rt_host_image_create: cannot create / map / page-lock %zu bytes of shared host memory on every rank", total);
    }
    c->shm_base = (char*)p; c->shm_bytes = total; c->hseq = 0;
    *out = c->shm_base + RT_SHM_HEADER;
    return RT_OK;
}

int rt_host_image_destroy(rt_ctx* c) {
    if (!c) return fail(nullptr, RT_ERR_ARG, "rt_host_image_destroy: NULL ctx");
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    shm_release(c);
    return RT_OK;
}

int rt_upload_scene(rt_ctx* c, const rt_scene* sc) {
    if (!c) return fail(nullptr, RT_ERR_ARG, "rt_upload_scene: NULL ctx");
    CU(c, cudaSetDevice(c->device));
    c->frame_valid = false;
    if (c->world > 1 && c->rank != 0) return broadcast_scene(c, RT_OK);   // collective: rank 0's scene (and status) arrives here
    int rc = sc ? upload_local(c, sc)
                : fail(c, RT_ERR_ARG, "rt_upload_scene: scene is NULL (only ranks != 0 of a multi-GPU context may receive)");
    if (c->inproc) {                                            // single-process group: peer copies of the arena, rank by rank
        if (rc != RT_OK) return rc;
        rc = build_wide_nodes(c);
        for (rt_ctx* m : c->members) {
            if (rc != RT_OK) break;
            CU(c, cudaSetDevice(m->device));
            m->frame_valid = false;
            free_scene(m);
            m->num_tris = c->num_tris; m->num_nodes = c->num_nodes; m->num_materials = c->num_materials;
            m->has_bvh = c->has_bvh; m->has_normals = c->has_normals; m->info = c->info;
            if (m->num_nodes) CU(c, cudaMalloc(&m->nodes, sizeof(BvhNode) * (size_t)m->num_nodes));
            CU(c, cudaMalloc(&m->geom, sizeof(TriBlock) * (size_t)m->num_tris));
            if (m->has_normals) CU(c, cudaMalloc(&m->shade, sizeof(TriBlock) * (size_t)m->num_tris));
            if (m->num_materials) CU(c, cudaMalloc(&m->materials, sizeof(rt_material) * (size_t)m->num_materials));
            if (m->num_nodes) CU(c, cudaMemcpyPeerAsync(m->nodes, m->device, c->nodes, c->device, sizeof(BvhNode) * (size_t)m->num_nodes, m->stream));
            CU(c, cudaMemcpyPeerAsync(m->geom, m->device, c->geom, c->device, sizeof(TriBlock) * (size_t)m->num_tris, m->stream));
            if (m->has_normals) CU(c, cudaMemcpyPeerAsync(m->shade, m->device, c->shade, c->device, sizeof(TriBlock) * (size_t)m->num_tris, m->stream));
            if (m->num_materials) CU(c, cudaMemcpyPeerAsync(m->materials, m->device, c->materials, c->device, sizeof(rt_material) * (size_t)m->num_materials, m->stream));
            CU(c, cudaStreamSynchronize(m->stream));
            m->has_scene = true;
            m->info.arena_bytes = scene_arena_bytes(m);
            m->info.build_ms = 0.f;
            rc = build_wide_nodes(m);
            if (rc != RT_OK) fail(c, rc, "rt_upload_scene: rank %d: %s", m->rank, m->err.c_str());
        }
        cudaSetDevice(c->device);
        return rc;
    }
    if (c->world > 1) return broadcast_scene(c, rc);           // also on failure: the other ranks are waiting in the broadcast
    if (rc != RT_OK) return rc;
    return build_wide_nodes(c);
}

int rt_build_info_get(const rt_ctx* c, rt_build_info* info) {
    if (!c || !info) return fail(nullptr, RT_ERR_ARG, "rt_build_info_get: NULL");
    if (!c->has_scene) return fail(const_cast<rt_ctx*>(c), RT_ERR_STATE, "rt_build_info_get: no scene uploaded");
    *info = c->info;
    return RT_OK;
}

} // extern "C"

namespace {
int ensure_band_resources(rt_ctx* c) {
    if (c->copy_stream) return RT_OK;
    for (cudaStream_t& s : c->band_stream) CU(c, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    // The copy stream also runs the one-block flag waits of the multi-GPU path: highest priority, so that such a block is
    // dispatched as soon as any SM slot frees up instead of queueing behind the frame kernel's pending blocks.
    int prio_lo = 0, prio_hi = 0;
    CU(c, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU(c, cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, prio_hi));
    // Chunk-completion flags are written by one-thread kernels: on a high-priority stream too, or they would only get an
    // SM slot after the pending blocks of the following chunk kernels.
    CU(c, cudaStreamCreateWithPriority(&c->sig_stream, cudaStreamNonBlocking, prio_hi));
    for (cudaEvent_t& e : c->band_ev) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming));
    return RT_OK;
}

// Local flag block for single-GPU band pipelining (multi-GPU contexts get theirs from peer_setup).
int ensure_flags(rt_ctx* c) {
    if (c->flags) return RT_OK;
    CU(c, cudaMalloc(&c->flags, kFlagBytes));
    CU(c, cudaMemset(c->flags, 0, kFlagBytes));
    return RT_OK;
}

// Band flags the HOST polls (rt_render_into is a blocking call: its own thread waits for a band's word and issues the band's
// copy — measured faster than stream-ordered cuStreamWaitValue32 waits on the copy stream).
int ensure_host_flags(rt_ctx* c) {
    if (c->hflags) return RT_OK;
    CU(c, cudaHostAlloc((void**)&c->hflags, kFlagBytes, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(c->hflags, 0, kFlagBytes);
    CU(c, cudaHostGetDevicePointer((void**)&c->hflags_dev, c->hflags, 0));
    return RT_OK;
}


// Bands of a rank's L local tile slots (completion flags of the persistent kernel).  Cuts at every ownership-chunk boundary
// (chunks of `chunk` slots: a band must not straddle two chunks, their tiles are not contiguous in the frame) and, when
// `refine`, at 1/2, 3/4, 7/8 ... of the LAST chunk: the copy of the last band is the only one that cannot overlap rendering,
// so it should be short.  Returns the number of bands (<= RT_MAX_BANDS), band_end[] cumulative.
int plan_bands(int L, int chunk, bool refine, int* band_end) {
    int n = 0;
    if (L <= 0) { band_end[0] = 0; return 1; }
    if (chunk < 1) chunk = L;
    int nchunks = (L + chunk - 1) / chunk;
    if (nchunks > RT_MAX_BANDS) { band_end[0] = L; return 1; }        // (callers check; one band = the whole rank)
    for (int k = 1; k < nchunks; ++k) band_end[n++] = k * chunk;
    int lo = (nchunks - 1) * chunk;                                     // refine [lo, L): 1/2, 3/4, 7/8 (every band costs each block a
    for (int k = 0; refine && k < 3 && n + 1 < RT_MAX_BANDS && L - lo >= 64; ++k) {   // barrier and a fence: measured +70 us per frame with 7 bands)
        lo += (L - lo) / 2;
        band_end[n++] = lo;
    }
    band_end[n++] = L;
    return n;
}

// ---- band copies of rt_render_into (the frame's parameters are c->fp, the destinations c->pump_outs) ----
// rows [y0, y1) of every requested plane
int band_copy_rows(rt_ctx* c, size_t y0, size_t y1) {
    const FrameParams& P = c->fp;
    for (const HostOut& o : c->pump_outs) {
        if (!o.host) continue;
        const size_t off = y0 * (size_t)P.W * o.bpp, bytes = (y1 - y0) * (size_t)P.W * o.bpp;
        CU(c, cudaMemcpyAsync((char*)o.host + off, (const char*)o.dev + off, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    return RT_OK;
}
// the pixels of the row-major GLOBAL tiles [g0, g1) (row-major device planes): a partial tile row, whole tile rows (one contiguous
// copy), a partial tile row
int band_copy_tiles(rt_ctx* c, long long g0, long long g1) {
    const FrameParams& P = c->fp;
    const long long total_tiles = (long long)P.tiles_x * P.tiles_y;
    if (g1 > total_tiles) g1 = total_tiles;
    if (g0 >= g1) return RT_OK;
    const int tx = P.tiles_x;
    const int row0 = (int)(g0 / tx), col0 = (int)(g0 % tx), row1 = (int)((g1 - 1) / tx), col1 = (int)((g1 - 1) % tx) + 1;
    struct Piece { int r0, r1, c0, c1; } pieces[3];
    int np = 0;
    if (row0 == row1) pieces[np++] = {row0, row0, col0, col1};
    else {
        int first_full = row0, last_full = row1;
        if (col0 != 0) { pieces[np++] = {row0, row0, col0, tx}; first_full = row0 + 1; }
        if (col1 != tx) { pieces[np++] = {row1, row1, 0, col1}; last_full = row1 - 1; }
        if (first_full <= last_full) pieces[np++] = {first_full, last_full, 0, tx};
    }
    for (int q = 0; q < np; ++q) {
        const size_t x0 = (size_t)pieces[q].c0 * RT_TILE_W, y0 = (size_t)pieces[q].r0 * RT_TILE_H;
        size_t x1 = (size_t)pieces[q].c1 * RT_TILE_W, y1 = (size_t)(pieces[q].r1 + 1) * RT_TILE_H;
        if (x1 > (size_t)P.W) x1 = (size_t)P.W;
        if (y1 > (size_t)P.H) y1 = (size_t)P.H;
        if (x0 >= x1 || y0 >= y1) continue;
        if (x0 == 0 && x1 == (size_t)P.W) { int rc = band_copy_rows(c, y0, y1); if (rc != RT_OK) return rc; continue; }
        for (const HostOut& o : c->pump_outs) {
            if (!o.host) continue;
            const size_t pitch = (size_t)P.W * o.bpp, off = y0 * pitch + x0 * o.bpp;
            CU(c, cudaMemcpy2DAsync((char*)o.host + off, pitch, (const char*)o.dev + off, pitch, (x1 - x0) * o.bpp, y1 - y0, cudaMemcpyDeviceToHost, c->copy_stream));
        }
    }
    return RT_OK;
}
// the same for the rank-local slot range [s0, s1), which must lie inside one ownership chunk
int band_copy_slots(rt_ctx* c, int s0, int s1) {
    const FrameParams& P = c->fp;
    if (s0 >= s1) return RT_OK;
    if (P.world <= 1) return band_copy_tiles(c, s0, s1);
    const int j = s0 / P.chunk_tiles;
    const long long base = ((long long)j * P.world + P.rank) * P.chunk_tiles;
    return band_copy_tiles(c, base + (s0 - j * P.chunk_tiles), base + (s1 - j * P.chunk_tiles));
}
// Host-driven band pipeline: wait for the next band's flag word (written by the frame kernel into mapped host memory), then
// enqueue the band's copies.  wait == false: only the bands that are already published (used to pump several ranks from one
// thread); returns RT_OK with c->pump_next == num_chunks when the rank is done.  Bounded: a kernel that died never writes its flags.
int band_pump(rt_ctx* c, bool wait) {
    const FrameParams& P = c->fp;
    const unsigned seq = c->pump_seq;
    const auto t0 = std::chrono::steady_clock::now();
    static const bool timing = getenv("RT_TIMING") != nullptr;
    CU(c, cudaSetDevice(c->device));
    while (c->pump_next < P.num_chunks) {
        const int j = c->pump_next;
        volatile unsigned* f = c->hflags + RT_PEER_CHUNK_FLAG(0, j);
        unsigned spins = 0;
        while ((int)(*f - seq) < 0) {
            if (!wait) return RT_OK;
            if ((++spins & 0x3fffu) == 0u) {
                const cudaError_t q = cudaStreamQuery(c->stream);
                if (q != cudaErrorNotReady && (int)(*f - seq) < 0) {
                    if (q != cudaSuccess) return fail(c, RT_ERR_CUDA, "rt_render_into: %s", cudaGetErrorString(q));
                    return fail(c, RT_ERR_STATE, "rt_render_into: the frame kernel finished without publishing band %d", j);
                }
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e9 > (double)c->peer_timeout_ns)
                    return fail(c, RT_ERR_STATE, "rt_render_into: band %d not finished after %llu s", j, c->peer_timeout_ns / 1000000000ull);
            }
        }
        if (timing) fprintf(stderr, "[band_pump r%d] band %d of %d published %8.1f us after the pump started\n", c->rank, j, P.num_chunks, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
        int rc = band_copy_slots(c, j ? P.band_end[j - 1] : 0, P.band_end[j]);
        if (rc != RT_OK) return rc;
        ++c->pump_next;
    }
    return RT_OK;
}

// rt_render (into == NULL) and rt_render_into (into != NULL: finished bands of the frame are copied to the caller's host
// buffers while later bands are still rendering).  One frame = ONE launch of the persistent kernel per rank: it
// publishes a flag word per finished band itself (and, on several GPUs, runs the ready / done handshake with rank 0),
// the copy stream waits on those words with cuStreamWaitValue32.  Kernels that do not publish their own completion
// (brute force, per-ray, the block-per-tile variants) get the flag kernels around them.
int render_impl(rt_ctx* c, const rt_frame* fr, rt_image* into, bool* pipelined) {
    if (pipelined) *pipelined = false;
    if (!c || !fr) return fail(c, RT_ERR_ARG, "rt_render: NULL argument");
    CU(c, cudaSetDevice(c->device));
    if (!c->has_scene) return fail(c, RT_ERR_STATE, "rt_render: no scene uploaded");
    if (fr->width < 1 || fr->height < 1) return fail(c, RT_ERR_ARG, "rt_render: pixel_width/pixel_height must be >= 1");
    if ((uint64_t)fr->width * (uint64_t)fr->height > (1ull << 31)) return fail(c, RT_ERR_ARG, "rt_render: image too large");
    if (fr->spp < 1) return fail(c, RT_ERR_ARG, "rt_render: spp must be >= 1");
    if (fr->mode != RT_MODE_HW1 && fr->mode != RT_MODE_HW2_BVH && fr->mode != RT_MODE_HW2_CPU) return fail(c, RT_ERR_ARG, "rt_render: unsupported mode %d", fr->mode);
    if (fr->mode == RT_MODE_HW2_CPU) {
        if (fr->accel != RT_ACCEL_BVH) return fail(c, RT_ERR_ARG, "rt_render: RT_MODE_HW2_CPU runs over the BVH (RT_ACCEL_BVH)");
        if (fr->diffuse_bounce) return fail(c, RT_ERR_ARG, "rt_render: RT_MODE_HW2_CPU has no deterministic diffuse bounce (the reference draws from std::random_device)");
        if (fr->max_depth > RT_CPU_MAX_DEPTH) return fail(c, RT_ERR_ARG, "rt_render: RT_MODE_HW2_CPU supports max_depth <= %d", RT_CPU_MAX_DEPTH);
    }
    if (fr->accel != RT_ACCEL_BRUTE && fr->accel != RT_ACCEL_BVH) return fail(c, RT_ERR_ARG, "rt_render: unsupported accel %d", fr->accel);
    if (fr->accel == RT_ACCEL_BRUTE && fr->mode != RT_MODE_HW1 && fr->max_depth > 1)
        return fail(c, RT_ERR_ARG, "rt_render: bounces (max_depth > 1) need RT_ACCEL_BVH");
    if (fr->accel == RT_ACCEL_BVH && !c->has_bvh) return fail(c, RT_ERR_STATE, "rt_render: scene was uploaded with RT_BUILD_NO_BVH");
    if (fr->num_lights < 0 || (fr->num_lights > 0 && !fr->lights)) return fail(c, RT_ERR_ARG, "rt_render: lights");
    if (fr->mode == RT_MODE_HW1 && fr->num_lights < 1) return fail(c, RT_ERR_ARG, "rt_render: HW1 mode needs one light");
    if (fr->quantiser < RT_QUANT_PPM_LROUND || fr->quantiser > RT_QUANT_CPU_TRUNC) return fail(c, RT_ERR_ARG, "rt_render: quantiser");
    if (fr->light_radius) {
        for (int i = 0; i < fr->num_lights; ++i) {
            if (!(fr->light_radius[i] >= 0.0f)) return fail(c, RT_ERR_ARG, "rt_render: light_radius[%d] must be >= 0", i);
            if (fr->light_radius[i] > 0.0f && fr->mode != RT_MODE_HW2_CPU)
                return fail(c, RT_ERR_ARG, "rt_render: disk lights (light_radius > 0) belong to RT_MODE_HW2_CPU (the CPUOnly renderer's ShadowVisibility)");
            if (fr->light_shadow_samples && fr->light_shadow_samples[i] > 4096) return fail(c, RT_ERR_ARG, "rt_render: light_shadow_samples[%d] > 4096", i);
        }
    }
    const uint32_t outputs = fr->outputs ? fr->outputs : RT_OUT_RGB_F32;
    if (into) {   // checked before any device work or collective step, so that an error leaves every rank in the same state
        const HostOut chk[4] = {{into->rgb, RT_OUT_RGB_F32, 12, nullptr, "rgb"}, {into->rgb8, RT_OUT_RGB8, 3, nullptr, "rgb8"},
                                {into->tri_id, RT_OUT_TRI_ID, 4, nullptr, "tri_id"}, {into->t, RT_OUT_T, 4, nullptr, "t"}};
        for (const HostOut& o : chk)
            if (o.host && !(outputs & o.bit)) return fail(c, RT_ERR_STATE, "rt_render_into: plane '%s' was not requested in rt_frame.outputs", o.name);
    }

    FrameParams& P = c->fp;
    memset(&P, 0, sizeof P);
    P.cam = fr->cam; P.mode = fr->mode; P.accel = fr->accel; P.W = fr->width; P.H = fr->height; P.spp = fr->spp;
    P.max_depth = fr->max_depth; P.diffuse_bounce = fr->diffuse_bounce ? 1 : 0; P.shadows = fr->shadows; P.quantiser = fr->quantiser; P.num_lights = fr->num_lights;
    P.num_materials = c->num_materials; P.has_normals = c->has_normals ? 1 : 0;
    P.sample_group = 1;
    while (P.sample_group < 32 && fr->spp % (2 * P.sample_group) == 0) P.sample_group *= 2;
    if (fr->kernel_variant == RT_VARIANT_PACKET_PIXEL_MAJOR || fr->kernel_variant == RT_VARIANT_PERSIST_EXACT_MT ||
        fr->kernel_variant == RT_VARIANT_PERSIST_OCC8 || fr->kernel_variant == RT_VARIANT_PERSIST_OCC10 || fr->kernel_variant == RT_VARIANT_PACKET_EXACT_SLAB)
        P.sample_group = 1;   // experimental variants: pixel-major
    memcpy(P.miss, fr->miss_color, sizeof P.miss);
    P.nodes = c->nodes; P.wide = c->wide; P.geom = c->geom; P.shade = c->shade; P.num_tris = c->num_tris; P.materials = c->materials;
    P.tiles_x = (P.W + RT_TILE_W - 1) / RT_TILE_W; P.tiles_y = (P.H + RT_TILE_H - 1) / RT_TILE_H;
    P.rank = c->rank; P.world = c->world;
    const bool dbg_shard = c->world == 1 && c->dbg_world > 1;
    if (dbg_shard) { P.rank = c->dbg_rank; P.world = c->dbg_world; }
    // rt_render_into with every plane inside the shared host image (rt_host_image_create): each rank renders into its own
    // row-major planes and copies the bands it owns to the host itself; nothing is gathered on rank 0's device
    bool host_direct = false;
    if (c->world > 1 && into && c->shm_base) {
        const int F0 = c->chunks_per_rank > 0 ? c->chunks_per_rank : RT_DEFAULT_CHUNKS_PER_RANK;
        const size_t np = (size_t)fr->width * fr->height;
        const void* pl[4] = {into->rgb, into->rgb8, into->tri_id, into->t};
        const size_t bp[4] = {12, 3, 4, 4};
        int n_in = 0, n_out = 0;
        for (int i = 0; i < 4; ++i) {
            if (!pl[i]) continue;
            const char* a = (const char*)pl[i];
            if (a >= c->shm_base + RT_SHM_HEADER && a + bp[i] * np <= c->shm_base + c->shm_bytes) ++n_in; else ++n_out;
        }
        if (n_in && n_out) return fail(c, RT_ERR_ARG, "rt_render_into: some planes are inside the shared host image and some are not");
        host_direct = n_in > 0 && F0 <= RT_PEER_MAX_CHUNKS;
    }
    if (c->inproc && into && (into->rgb || into->rgb8 || into->tri_id || into->t)) host_direct = true;   // one process: every rank can write the caller's buffers
    if (c->world > 1 && c->peer && !host_direct) {
        int rc = c->inproc ? inproc_planes(c, c->group_root, outputs, (size_t)fr->width * fr->height)
                           : peer_planes(c, outputs, (size_t)fr->width * fr->height);   // collective; may fall back to the NCCL gather
        if (rc != RT_OK) return rc;
    }
    const bool peer = c->world > 1 && c->peer && !host_direct;
    P.packed = ((c->world > 1 && !peer && !host_direct) || dbg_shard) ? 1 : 0;
    const int total_tiles = P.tiles_x * P.tiles_y;
    P.chunk_tiles = rt_chunk_tiles(total_tiles, P.world, c->chunks_per_rank);
    P.local_tiles = rt_tiles_of_rank(total_tiles, P.world, c->chunks_per_rank);
    {   // fused slab test only when the camera is within 8 scene extents of the scene (rt_core.h, rt_slab_fma)
        float ext = 0.f, far = 0.f;
        for (int k = 0; k < 3; ++k) {
            ext = fmaxf(ext, c->info.scene_max[k] - c->info.scene_min[k]);
            ext = fmaxf(ext, fmaxf(fabsf(c->info.scene_max[k]), fabsf(c->info.scene_min[k])));
            far = fmaxf(far, fabsf(fr->cam.center[k]));
        }
        P.fast_slab = (c->has_bvh && far <= 8.0f * ext) ? 1 : 0;
        P.frustum_eps = 1.6e-5f * fmaxf(ext, far) + 1e-30f;
        float r2 = 0.f;
        for (int k = 0; k < 3; ++k) {
            P.scene_c[k] = 0.5f * c->info.scene_min[k] + 0.5f * c->info.scene_max[k];
            const float e = c->info.scene_max[k] - c->info.scene_min[k];
            r2 += 0.25f * e * e;
        }
        P.scene_r2 = r2 * 1.03f + 1e-30f;   // 3 %: far beyond the 2^-17 padding of the node boxes
        {   // rounding of (c-o).d and |c-o|^2 must stay far below the padding: coordinates <= 1000 r, camera within 30 r
            const float r = sqrtf(r2);
            float dc2 = 0.f, maxabs = 0.f;
            for (int k = 0; k < 3; ++k) {
                const float d = fr->cam.center[k] - P.scene_c[k];
                dc2 += d * d;
                maxabs = fmaxf(maxabs, fmaxf(fabsf(fr->cam.center[k]), fmaxf(fabsf(c->info.scene_min[k]), fabsf(c->info.scene_max[k]))));
            }
            if (!(r > 0.f) || !(maxabs <= 1000.f * r) || !(dc2 <= 900.f * r2)) P.scene_r2 = INFINITY;
        }
    }

    if (fr->num_lights) {
        const size_t nl = (size_t)fr->num_lights, lbytes = (sizeof(rt_light) * nl + 15) & ~(size_t)15;
        CU(c, c->lights.reserve(lbytes + 8 * nl));              // lights, then (soft shadows) radii and sample counts
        CU(c, cudaMemcpyAsync(c->lights.p, fr->lights, sizeof(rt_light) * nl, cudaMemcpyHostToDevice, c->stream));
        if (fr->light_radius && fr->mode == RT_MODE_HW2_CPU) {
            float* d_rad = (float*)((char*)c->lights.p + lbytes);
            CU(c, cudaMemcpyAsync(d_rad, fr->light_radius, 4 * nl, cudaMemcpyHostToDevice, c->stream));
            P.light_radius = d_rad;
            if (fr->light_shadow_samples) {
                int* d_ns = (int*)(d_rad + nl);
                CU(c, cudaMemcpyAsync(d_ns, fr->light_shadow_samples, 4 * nl, cudaMemcpyHostToDevice, c->stream));
                P.light_samples = d_ns;
            }
        }
    }
    P.rng_seed = fr->rng_seed;
    P.lights = (const rt_light*)c->lights.p;
    if (fr->jitter) {
        CU(c, c->jitter.reserve(sizeof(float) * 2 * (size_t)fr->spp));
        CU(c, cudaMemcpyAsync(c->jitter.p, fr->jitter, sizeof(float) * 2 * (size_t)fr->spp, cudaMemcpyHostToDevice, c->stream));
        P.jitter = (const float*)c->jitter.p;
    }
    const size_t npix_full = (size_t)P.W * P.H;
    const size_t npix_loc = (P.world == 1 || host_direct) ? npix_full : (size_t)P.local_tiles * RT_BLOCK_THREADS;
    if (peer) {                                    // every rank writes rank 0's row-major image in place
        const bool root = c->rank == 0;
        if (outputs & RT_OUT_RGB_F32) P.rgb = (float*)(root ? c->img_rgb.p : c->peer_img[0]);
        if (outputs & RT_OUT_RGB8) P.rgb8 = (uint8_t*)(root ? c->img_rgb8.p : c->peer_img[1]);
        if (outputs & RT_OUT_TRI_ID) P.tri_id = (int32_t*)(root ? c->img_id.p : c->peer_img[2]);
        if (outputs & RT_OUT_T) P.t = (float*)(root ? c->img_t.p : c->peer_img[3]);
    } else {
        if (outputs & RT_OUT_RGB_F32) { CU(c, c->loc_rgb.reserve(12 * npix_loc + 16)); P.rgb = (float*)c->loc_rgb.p; }
        if (outputs & RT_OUT_RGB8) { CU(c, c->loc_rgb8.reserve(3 * npix_loc + 16)); P.rgb8 = (uint8_t*)c->loc_rgb8.p; }
        if (outputs & RT_OUT_TRI_ID) { CU(c, c->loc_id.reserve(4 * npix_loc + 16)); P.tri_id = (int32_t*)c->loc_id.p; }
        if (outputs & RT_OUT_T) { CU(c, c->loc_t.reserve(4 * npix_loc + 16)); P.t = (float*)c->loc_t.p; }
    }
    // ray counters (8 x u64) and the persistent kernel's work queue share one allocation and one memset per frame
    P.counters = (unsigned long long*)c->counters.p;
    P.queue = (unsigned*)((char*)c->counters.p + RT_COUNTER_BYTES);
    CU(c, cudaMemsetAsync(c->counters.p, 0, RT_COUNTER_BYTES + RT_PERSIST_CTL_BYTES, c->stream));
    P.peer_timeout_ns = c->peer_timeout_ns;
    P.peer_err = c->peer_err;
    // RT_VARIANT_DEFAULT: the persistent kernel where the completion protocol has to live inside the kernel (banded peer-store
    // gather); frames that go out through the shared host image / the caller's buffers use the faster block-per-tile kernel, one
    // launch per band (below), like single-GPU rt_render_into.
    const bool persistent = rt_render_is_persistent(P, fr->kernel_variant, into != nullptr && !host_direct);
    P.persist = persistent ? 1 : 0;
    const HostOut outs[4] = {{into ? into->rgb : nullptr, RT_OUT_RGB_F32, 12, P.rgb, "rgb"}, {into ? into->rgb8 : nullptr, RT_OUT_RGB8, 3, P.rgb8, "rgb8"},
                             {into ? into->tri_id : nullptr, RT_OUT_TRI_ID, 4, P.tri_id, "tri_id"}, {into ? into->t : nullptr, RT_OUT_T, 4, P.t, "t"}};
    c->pump_outs[0] = outs[0]; c->pump_outs[1] = outs[1]; c->pump_outs[2] = outs[2]; c->pump_outs[3] = outs[3];
    auto copy_rows = [&](size_t y0, size_t y1) -> int { return band_copy_rows(c, y0, y1); };
    auto copy_slots = [&](int s0, int s1) -> int { return band_copy_slots(c, s0, s1); };
    auto pump_bands = [&](unsigned seq) -> int {
        c->pump_seq = seq; c->pump_next = 0; c->last_pumped = true;
        if (c->inproc) return RT_OK;                 // single-process group: the caller pumps every rank's bands together (rt_render_into)
        return band_pump(c, true);
    };
    (void)copy_rows; (void)copy_slots;

    CU(c, cudaEventRecord(c->ev0, c->stream));
    int launches = 0;
    c->last_host_direct = false;
    c->last_pumped = false;
    c->fast_finish = false;
    if (host_direct) {
        int rc = ensure_band_resources(c);
        if (rc != RT_OK) return rc;
        if (rc == RT_OK) rc = ensure_host_flags(c);
        if (rc != RT_OK) return rc;
        const unsigned seq = ++c->hseq;
        volatile unsigned* hdr = (volatile unsigned*)c->shm_base;
        // the caller of rank 0 is done with the previous frame's image once it is back in here: rank 0 says so, the others
        // wait for it before anything of this frame can land in the shared buffer (host-side handshake through the header)
        if (c->inproc) { /* one thread drives every rank: nothing to wait for */ }
        else if (c->rank == 0) { __sync_synchronize(); hdr[0] = seq; }
        else {
            const auto t0 = std::chrono::steady_clock::now();
            while ((int)(hdr[0] - seq) < 0) {
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e9 > (double)c->peer_timeout_ns)
                    return fail(c, RT_ERR_STATE, "rt_render_into: rank %d waited more than %llu s for rank 0 to enter the frame", c->rank, c->peer_timeout_ns / 1000000000ull);
            }
        }
        P.num_chunks = plan_bands(P.local_tiles, P.chunk_tiles, true, P.band_end);
        P.seq = seq;
        P.flags = c->hflags_dev + RT_PEER_CHUNK_FLAG(0, 0);
        if (persistent) P.host_counters = c->pin_dev + 8;
        int l = 0;
        CU(c, cudaEventRecord(c->evk0, c->stream));
        StreamWriteValue32Fn band_w32 = stream_write_value32();
        const bool band_launches = !persistent && P.num_chunks <= RT_BANDS && P.accel == RT_ACCEL_BVH;
        if (band_launches) {
            // Block-per-tile kernel (the faster one): one launch per band on a few streams, so that the hardware scheduler
            // fills the tail of one band with the blocks of the next; the band's flag word is written into mapped host memory by a
            // stream memory operation behind its kernel (no flag kernel, no SM), and this thread's pump copies the band out.
            int start = 0;
            for (int k = 0; k < P.num_chunks; ++k) {
                FrameParams Q = P;
                Q.tile_offset = P.tile_offset + start; Q.local_tiles = P.band_end[k] - start;
                start = P.band_end[k];
                cudaStream_t s = c->band_stream[k % RT_BAND_STREAMS];
                CU(c, cudaStreamWaitEvent(s, c->evk0, 0));
                int lk = 0;
                if (Q.local_tiles > 0) CU(c, rt_launch_render(Q, fr->kernel_variant, s, &lk));
                launches += lk;
                unsigned* flag = P.flags + (size_t)k * RT_PEER_FLAG_STRIDE;
                if (!band_w32 || band_w32((CUstream)s, (CUdeviceptr)flag, seq, 0) != CUDA_SUCCESS) {
                    CU(c, rt_launch_flag_set(flag, RT_PEER_FLAG_STRIDE, 1, seq, s));
                    ++launches;
                }
                CU(c, cudaEventRecord(c->band_ev[k], s));
                CU(c, cudaStreamWaitEvent(c->stream, c->band_ev[k], 0));
            }
        } else {
            CU(c, rt_launch_render(P, fr->kernel_variant, c->stream, &l));
            launches += l;
            if (!persistent) { CU(c, rt_launch_flag_set(P.flags, RT_PEER_FLAG_STRIDE, P.num_chunks, seq, c->stream)); ++launches; }
        }
        CU(c, cudaEventRecord(c->evk1, c->stream));
        CU(c, cudaStreamWaitEvent(c->copy_stream, c->evk0, 0));
        rc = pump_bands(seq);
        if (rc != RT_OK) return rc;
        if (!c->inproc) {   // this rank's bands are in the shared buffer: say so in the header (a store to mapped host memory, after the copies in stream order)
            unsigned* dhdr = nullptr;
            CU(c, cudaHostGetDevicePointer((void**)&dhdr, c->shm_base, 0));
            StreamWriteValue32Fn write32 = stream_write_value32();          // a stream memory operation: no kernel launch, no SM
            if (!write32 || write32((CUstream)c->copy_stream, (CUdeviceptr)(dhdr + 16 * (1 + c->rank)), seq, 0) != CUDA_SUCCESS) {
                CU(c, rt_launch_flag_set(dhdr + 16 * (1 + c->rank), 1, 1, seq, c->copy_stream));
                ++launches;
            }
        }
        // end of the frame without a stream synchronisation: a stream memory operation writes "copies done" into page-locked host
        // memory behind the last copy; rt_render_into polls that word and takes the ray counts from the words the kernel's last
        // block wrote (measured: the event + small device->host reads + cudaStreamSynchronize it replaces cost 70-170 us per frame)
        c->fast_finish = false;
        if (persistent && !c->inproc) {
            StreamWriteValue32Fn w32 = stream_write_value32();
            if (w32 && w32((CUstream)c->copy_stream, (CUdeviceptr)(c->pin_dev + 12), seq, 0) == CUDA_SUCCESS) c->fast_finish = true;
        } else if (band_launches && !c->inproc) {
            // (the pump has seen every band's flag, i.e. every band kernel has finished: the ray counters are final)
            StreamWriteValue32Fn w32 = stream_write_value32();
            if (w32) {
                CU(c, cudaMemcpyAsync(c->pin + 8, c->counters.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->copy_stream));
                if (w32((CUstream)c->copy_stream, (CUdeviceptr)(c->pin_dev + 12), seq, 0) == CUDA_SUCCESS) { c->fast_finish = true; c->pin[10] = seq; }
            }
        }
        CU(c, cudaEventRecord(c->copy_done, c->copy_stream));
        if (!c->fast_finish) CU(c, cudaStreamWaitEvent(c->stream, c->copy_done, 0));
        if (pipelined) *pipelined = true;
        c->last_host_direct = !c->inproc;
    } else if (peer) {
        // Frame k may overwrite rank 0's image only once rank 0 is done with frame k-1 (its downloads are stream-ordered
        // before this point): rank 0 publishes `ready = k`, the others wait for it — inside the persistent kernel.
        const unsigned seq = ++c->seq;
        const int F = c->chunks_per_rank > 0 ? c->chunks_per_rank : RT_DEFAULT_CHUNKS_PER_RANK;
        // one flag per ownership band when rank 0 pipelines its copies behind them (rt_render_into), else one per rank: every band
        // costs every block a barrier and a system-scope fence over NVLink (measured at 8 GPUs: 0.35 ms per-rank kernel with 4
        // bands, where the round-1 kernel without them took 0.30 ms)
        const bool banded = into != nullptr && F <= RT_PEER_MAX_CHUNKS;
        P.num_chunks = plan_bands(P.local_tiles, banded ? P.chunk_tiles : P.local_tiles, false, P.band_end);   // one flag per ownership band, else one per rank
        P.seq = seq;
        P.flags = c->flags + RT_PEER_CHUNK_FLAG(c->rank, 0);
        const bool root = c->rank == 0;
        P.peer_stores = root ? 0 : 1;
        if (persistent) {
            if (root) { P.ready_out = c->flags; P.wait_ranks = c->world - 1; P.flags_base = c->flags; }
            else P.ready_in = c->flags;
        } else {
            if (root) CU(c, rt_launch_flag_set(c->flags, 1, 1, seq, c->stream));
            else CU(c, rt_launch_flag_wait(c->flags, 0, 1, 0, 1, seq, c->peer_timeout_ns, c->peer_err, c->stream));
            ++launches;
        }
        const bool pipe = into != nullptr && root && banded;
        if (pipe) { int rc = ensure_band_resources(c); if (rc != RT_OK) return rc; }
        int l = 0;
        CU(c, cudaEventRecord(c->evk0, c->stream));
        CU(c, rt_launch_render(P, fr->kernel_variant, c->stream, &l));
        launches += l;
        if (!persistent) { CU(c, rt_launch_flag_set(P.flags, RT_PEER_FLAG_STRIDE, P.num_chunks, seq, c->stream)); ++launches; }
        if (root && !persistent) {
            CU(c, rt_launch_flag_wait(c->flags + RT_PEER_CHUNK_FLAG(1, 0), RT_PEER_MAX_CHUNKS * RT_PEER_FLAG_STRIDE, c->world - 1, RT_PEER_FLAG_STRIDE, P.num_chunks,
                                      seq, c->peer_timeout_ns, c->peer_err, c->stream));
            ++launches;
        }
        CU(c, cudaEventRecord(c->evk1, c->stream));             // rank 0: every rank's pixels have landed
        if (pipe) {
            // Rank 0 copies the image to the host group by group: group j = band j of every rank = a contiguous range of
            // row-major tiles; once all its flags arrived, the pixel rows it completes go out while later bands render.
            CU(c, cudaStreamWaitEvent(c->copy_stream, c->evk0, 0));
            StreamWaitValue32Fn wait32 = stream_wait_value32();
            size_t y_prev = 0;
            for (int j = 0; j < F; ++j) {
                if (wait32) {
                    for (int r = 0; r < c->world; ++r)
                        if (wait32((CUstream)c->copy_stream, (CUdeviceptr)(c->flags + RT_PEER_CHUNK_FLAG(r, j)), seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                            return fail(c, RT_ERR_CUDA, "cuStreamWaitValue32 failed");
                } else if (j == 0) {                            // no stream-ordered waits on this driver: copy after the frame
                    CU(c, cudaStreamWaitEvent(c->copy_stream, c->evk1, 0));
                }
                long long tiles_done = (long long)(j + 1) * c->world * P.chunk_tiles;
                if (tiles_done > total_tiles) tiles_done = total_tiles;
                size_t y_end = (j == F - 1) ? (size_t)P.H : (size_t)(tiles_done / P.tiles_x) * RT_TILE_H;
                if (y_end > (size_t)P.H) y_end = (size_t)P.H;
                if (y_end <= y_prev) continue;
                int rc = copy_rows(y_prev, y_end);
                if (rc != RT_OK) return rc;
                y_prev = y_end;
            }
            CU(c, cudaEventRecord(c->copy_done, c->copy_stream));
            if (pipelined) *pipelined = true;
            // had a wait timed out (a rank died), release the stream-ordered band waits so that nothing hangs
            CU(c, rt_launch_flag_unblock(c->peer_err, c->flags + RT_PEER_CHUNK_FLAG(0, 0), RT_PEER_FLAG_STRIDE, RT_PEER_MAX_RANKS * RT_PEER_MAX_CHUNKS, seq, c->stream));
            ++launches;
            CU(c, cudaStreamWaitEvent(c->stream, c->copy_done, 0));
        }
    } else if (into && c->world == 1 && !dbg_shard && P.tiles_y >= 2 && persistent) {
        // single GPU: one persistent launch; each band is published by the kernel as it completes and copied out by this thread
        int rc = ensure_band_resources(c);
        if (rc == RT_OK) rc = ensure_host_flags(c);
        if (rc != RT_OK) return rc;
        const unsigned seq = ++c->hseq;
        P.num_chunks = plan_bands(P.local_tiles, (P.local_tiles + 1) / 2, true, P.band_end);     // halves, the second one halved three times: 1/2, 3/4, 7/8, 15/16, 1
        P.seq = seq;
        P.flags = c->hflags_dev + RT_PEER_CHUNK_FLAG(0, 0);
        P.host_counters = c->pin_dev + 8;
        CU(c, cudaEventRecord(c->evk0, c->stream));
        CU(c, rt_launch_render(P, fr->kernel_variant, c->stream, &launches));
        CU(c, cudaEventRecord(c->evk1, c->stream));
        CU(c, cudaStreamWaitEvent(c->copy_stream, c->evk0, 0));
        rc = pump_bands(seq);
        if (rc != RT_OK) return rc;
        // end of the frame without a stream synchronisation: a stream memory operation writes "copies done" into page-locked host
        // memory behind the last copy; rt_render_into polls that word and takes the ray counts from the words the kernel's last
        // block wrote (measured: the event + small device->host reads + cudaStreamSynchronize it replaces cost 70-170 us per frame)
        c->fast_finish = false;
        if (persistent && !c->inproc) {
            StreamWriteValue32Fn w32 = stream_write_value32();
            if (w32 && w32((CUstream)c->copy_stream, (CUdeviceptr)(c->pin_dev + 12), seq, 0) == CUDA_SUCCESS) c->fast_finish = true;
        }
        CU(c, cudaEventRecord(c->copy_done, c->copy_stream));
        if (!c->fast_finish) CU(c, cudaStreamWaitEvent(c->stream, c->copy_done, 0));
        if (pipelined) *pipelined = true;
    } else if (into && c->world == 1 && !dbg_shard && P.tiles_y >= 2) {
        // kernels without in-kernel completion flags: one launch per band on a few streams, copies chained with events
        int rc = ensure_band_resources(c);
        if (rc != RT_OK) return rc;
        const int B = P.tiles_y < RT_BANDS ? P.tiles_y : RT_BANDS;
        // Band boundaries in tile rows.  Only the LAST band's copy cannot overlap the rendering, so the bands shrink towards the end
        // of the frame (the last one is 2 % of it: 0.5 MB of a 4K 8-bit frame instead of 3.1 MB with equal bands).
        auto band_row = [&](int k) -> int {
            static const int kCut[RT_BANDS + 1] = {0, 18, 36, 54, 70, 83, 92, 98, 100};     // per cent of the frame
            if (B != RT_BANDS) return (int)((long long)P.tiles_y * k / B);
            return (int)((long long)P.tiles_y * kCut[k] / 100);
        };
        CU(c, cudaEventRecord(c->evk0, c->stream));
        for (int k = 0; k < B; ++k) {                 // all band kernels first: a pageable destination makes the copies host-synchronous
            const int row_a = band_row(k), row_b = band_row(k + 1);
            FrameParams Q = P;
            Q.tile_offset = row_a * P.tiles_x; Q.local_tiles = (row_b - row_a) * P.tiles_x;
            cudaStream_t s = c->band_stream[k % RT_BAND_STREAMS];
            CU(c, cudaStreamWaitEvent(s, c->evk0, 0));
            int l = 0;
            if (row_b > row_a) CU(c, rt_launch_render(Q, fr->kernel_variant, s, &l));
            launches += l;
            CU(c, cudaEventRecord(c->band_ev[k], s));
            CU(c, cudaStreamWaitEvent(c->stream, c->band_ev[k], 0));
        }
        CU(c, cudaEventRecord(c->evk1, c->stream));    // every band kernel has finished
        for (int k = 0; k < B; ++k) {
            const int row_a = band_row(k), row_b = band_row(k + 1);
            if (row_b <= row_a) continue;
            const size_t y0 = (size_t)row_a * RT_TILE_H, y1 = (size_t)row_b * RT_TILE_H < (size_t)P.H ? (size_t)row_b * RT_TILE_H : (size_t)P.H;
            CU(c, cudaStreamWaitEvent(c->copy_stream, c->band_ev[k], 0));
            rc = copy_rows(y0, y1);
            if (rc != RT_OK) return rc;
        }
        {   // end of the frame without a stream synchronisation (see rt_render_into): ray counters into page-locked memory and a
            // "copies done" word behind the last copy, both on the copy stream, which has waited for every band kernel by now
            StreamWriteValue32Fn w32 = stream_write_value32();
            const unsigned seq = ++c->hseq;
            if (w32) {
                CU(c, cudaMemcpyAsync(c->pin + 8, c->counters.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->copy_stream));
                if (w32((CUstream)c->copy_stream, (CUdeviceptr)(c->pin_dev + 12), seq, 0) == CUDA_SUCCESS) { c->fast_finish = true; c->pump_seq = seq; c->pin[10] = seq; }
            }
        }
        CU(c, cudaEventRecord(c->copy_done, c->copy_stream));
        if (!c->fast_finish) CU(c, cudaStreamWaitEvent(c->stream, c->copy_done, 0));
        if (pipelined) *pipelined = true;
    } else {
        CU(c, cudaEventRecord(c->evk0, c->stream));
        CU(c, rt_launch_render(P, fr->kernel_variant, c->stream, &launches));
        CU(c, cudaEventRecord(c->evk1, c->stream));
    }

    if (c->world > 1 && !peer && !host_direct) {
        // Tile gather to rank 0: one grouped send/recv per requested plane, then an unpack kernel per source rank.
        const size_t other_pix = (size_t)(c->world - 1) * npix_loc;   // every rank has the same number of tile slots
        if (c->rank == 0) {
            if (outputs & RT_OUT_RGB_F32) { CU(c, c->img_rgb.reserve(12 * npix_full)); CU(c, c->stage_rgb.reserve(12 * other_pix + 16)); }
            if (outputs & RT_OUT_RGB8) { CU(c, c->img_rgb8.reserve(3 * npix_full)); CU(c, c->stage_rgb8.reserve(3 * other_pix + 16)); }
            if (outputs & RT_OUT_TRI_ID) { CU(c, c->img_id.reserve(4 * npix_full)); CU(c, c->stage_id.reserve(4 * other_pix + 16)); }
            if (outputs & RT_OUT_T) { CU(c, c->img_t.reserve(4 * npix_full)); CU(c, c->stage_t.reserve(4 * other_pix + 16)); }
        }
        struct PlaneIo { uint32_t bit; size_t bpp; Plane* loc; Plane* stage; };
        PlaneIo io[4] = {{RT_OUT_RGB_F32, 12, &c->loc_rgb, &c->stage_rgb}, {RT_OUT_RGB8, 3, &c->loc_rgb8, &c->stage_rgb8},
                         {RT_OUT_TRI_ID, 4, &c->loc_id, &c->stage_id}, {RT_OUT_T, 4, &c->loc_t, &c->stage_t}};
        NC(c, GroupStart());
        for (const PlaneIo& q : io) {
            if (!(outputs & q.bit)) continue;
            if (c->rank == 0) {
                size_t off = 0;
                for (int r = 1; r < c->world; ++r) {
                    size_t bytes = npix_loc * q.bpp;
                    if (bytes) NC(c, Recv((char*)q.stage->p + off, bytes, ncclUint8, r, c->comm, c->stream));
                    off += bytes;
                }
            } else {
                size_t bytes = npix_loc * q.bpp;
                if (bytes) NC(c, Send(q.loc->p, bytes, ncclUint8, 0, c->comm, c->stream));
            }
        }
        NC(c, GroupEnd());
        if (c->rank == 0) {
            size_t off_pix = 0;
            for (int r = 0; r < c->world; ++r) {
                const bool self = r == 0;
                const float* s_rgb = (outputs & RT_OUT_RGB_F32) ? (self ? (const float*)c->loc_rgb.p : (const float*)((char*)c->stage_rgb.p + 12 * off_pix)) : nullptr;
                const uint8_t* s_rgb8 = (outputs & RT_OUT_RGB8) ? (self ? (const uint8_t*)c->loc_rgb8.p : (const uint8_t*)c->stage_rgb8.p + 3 * off_pix) : nullptr;
                const int32_t* s_id = (outputs & RT_OUT_TRI_ID) ? (self ? (const int32_t*)c->loc_id.p : (const int32_t*)c->stage_id.p + off_pix) : nullptr;
                const float* s_t = (outputs & RT_OUT_T) ? (self ? (const float*)c->loc_t.p : (const float*)c->stage_t.p + off_pix) : nullptr;
                CU(c, rt_launch_unpack(P, r, s_rgb, s_rgb8, s_id, s_t, (float*)c->img_rgb.p, (uint8_t*)c->img_rgb8.p,
                                       (int32_t*)c->img_id.p, (float*)c->img_t.p, c->stream));
                ++launches;
                if (!self) off_pix += npix_loc;
            }
        }
    }
    CU(c, cudaEventRecord(c->ev1, c->stream));
    c->outputs = host_direct ? 0u : outputs;          // shared-host frames leave nothing gathered on rank 0's device
    c->launches = launches;
    c->frame_valid = true;
    return RT_OK;
}
} // namespace

extern "C" {

int rt_render(rt_ctx* c, const rt_frame* fr) {
    int rc = render_impl(c, fr, nullptr, nullptr);
    if (c && c->inproc && c->group_root == c)                   // single-process group: the same frame on every rank, from this thread
        for (rt_ctx* m : c->members) {
            if (rc != RT_OK) break;
            rc = render_impl(m, fr, nullptr, nullptr);
            if (rc != RT_OK) fail(c, rc, "rt_render: rank %d: %s", m->rank, m->err.c_str());
        }
    if (c) cudaSetDevice(c->device);
    return rc;
}

int rt_render_into(rt_ctx* c, const rt_frame* fr, rt_image* img) {
    if (!img) return fail(c, RT_ERR_ARG, "rt_render_into: NULL image");
    bool pipelined = false;
    static const bool timing = getenv("RT_TIMING") != nullptr;        // host-side phase times on stderr (diagnostic)
    const auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (timing && c) fprintf(stderr, "[rt_render_into r%d] %-28s %8.1f us\n", c->rank, what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tp0).count());
    };
    if (c && c->inproc && c->group_root == c) {
        // single-process group: every rank renders its bands and copies them straight into the caller's buffers (one host thread
        // enqueues all ranks' frames first, then pumps each rank's band copies); no gather on rank 0's device
        std::vector<rt_ctx*> all{c};
        all.insert(all.end(), c->members.begin(), c->members.end());
        int rc = RT_OK;
        uint64_t prim = 0, shad = 0;
        for (rt_ctx* m : all) {
            rc = render_impl(m, fr, img, &pipelined);
            if (rc != RT_OK) { if (m != c) fail(c, rc, "rt_render_into: rank %d: %s", m->rank, m->err.c_str()); break; }
        }
        // pump every rank's band copies from this thread until all bands of all ranks are on their way
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned spins = 0; rc == RT_OK; ++spins) {
            bool pending = false;
            for (rt_ctx* m : all) {
                if (!m->last_pumped || m->pump_next >= m->fp.num_chunks) continue;
                rc = band_pump(m, false);
                if (rc != RT_OK) { if (m != c) fail(c, rc, "rt_render_into: rank %d: %s", m->rank, m->err.c_str()); break; }
                pending = pending || m->pump_next < m->fp.num_chunks;
            }
            if (!pending || rc != RT_OK) break;
            if ((spins & 0xfffu) == 0xfffu) {
                for (rt_ctx* m : all) {
                    cudaSetDevice(m->device);
                    const cudaError_t q = cudaStreamQuery(m->stream);
                    if (q != cudaSuccess && q != cudaErrorNotReady) { rc = fail(c, RT_ERR_CUDA, "rt_render_into: rank %d: %s", m->rank, cudaGetErrorString(q)); break; }
                }
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e9 > (double)c->peer_timeout_ns)
                    rc = fail(c, RT_ERR_STATE, "rt_render_into: bands not finished after %llu s", c->peer_timeout_ns / 1000000000ull);
            }
        }
        for (rt_ctx* m : all) {
            if (rc != RT_OK || !m->last_pumped) continue;
            cudaSetDevice(m->device);
            if (cudaStreamSynchronize(m->copy_stream) != cudaSuccess) rc = fail(c, RT_ERR_CUDA, "rt_render_into: rank %d: copy stream", m->rank);
        }
        for (rt_ctx* m : all) {
            if (rc != RT_OK) break;
            rt_image meta{};
            rc = rt_download_image(m, &meta);
            if (rc != RT_OK) { if (m != c) fail(c, rc, "rt_render_into: rank %d: %s", m->rank, m->err.c_str()); break; }
            prim += meta.rays_primary; shad += meta.rays_shadow;
            if (m == c) { img->width = meta.width; img->height = meta.height; img->gpu_ms = meta.gpu_ms; }
        }
        img->rays_primary = prim; img->rays_shadow = shad;
        cudaSetDevice(c->device);
        return rc;
    }
    int rc = render_impl(c, fr, img, &pipelined);
    lap("frame + copies enqueued");
    if (rc != RT_OK) return rc;
    if (!pipelined) return rt_download_image(c, img);           // multi-GPU contexts: gather first, then one copy
    rt_image meta{};                                            // planes are already on their way: fetch counters and times only
    if (c->fast_finish) {
        const unsigned seq = c->pump_seq;
        volatile unsigned long long* pin = c->pin;
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned spins = 0; (unsigned)pin[12] != seq || (unsigned)pin[10] != seq; ++spins) {
            if ((spins & 0x3fffu) == 0x3fffu) {
                const cudaError_t q = cudaStreamQuery(c->copy_stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) return fail(c, RT_ERR_CUDA, "rt_render_into: %s", cudaGetErrorString(q));
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e9 > (double)c->peer_timeout_ns)
                    return fail(c, RT_ERR_STATE, "rt_render_into: copies not finished after %llu s", c->peer_timeout_ns / 1000000000ull);
            }
        }
        __sync_synchronize();
        meta.width = c->fp.W; meta.height = c->fp.H; meta.rays_primary = pin[8]; meta.rays_shadow = pin[9];
        float ms = 0.f;
        CU(c, cudaEventSynchronize(c->ev1));                    // (recorded behind the kernel, which ended before its last band's copy)
        CU(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        meta.gpu_ms = ms;
        rc = RT_OK;
    } else rc = rt_download_image(c, &meta);
    lap("stream drained, counters read");
    if (rc != RT_OK) return rc;
    if (c->last_host_direct && c->rank == 0) {                  // shared host image: every rank's bands must have landed
        volatile unsigned* hdr = (volatile unsigned*)c->shm_base;
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < c->world; ++r)
            while ((int)(hdr[16 * (1 + r)] - c->hseq) < 0) {
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e9 > (double)c->peer_timeout_ns)
                    return fail(c, RT_ERR_STATE, "rt_render_into: rank 0 waited more than %llu s for rank %d's bands", c->peer_timeout_ns / 1000000000ull, r);
            }
        __sync_synchronize();
        lap("every rank's bands landed");
    }
    img->width = meta.width; img->height = meta.height; img->rays_primary = meta.rays_primary; img->rays_shadow = meta.rays_shadow; img->gpu_ms = meta.gpu_ms;
    return RT_OK;
}

int rt_sync(rt_ctx* c, float* gpu_ms) {
    if (!c) return fail(nullptr, RT_ERR_ARG, "rt_sync: NULL ctx");
    CU(c, cudaSetDevice(c->device));
    if (!c->frame_valid) return fail(c, RT_ERR_STATE, "rt_sync: no frame rendered");
    c->pin[4] = 0ull;
    if (c->peer && c->peer_err) CU(c, cudaMemcpyAsync(&c->pin[4], c->peer_err, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    const unsigned perr = (unsigned)c->pin[4];
    if (perr) return peer_timeout(c, perr);
    if (gpu_ms) CU(c, cudaEventElapsedTime(gpu_ms, c->ev0, c->ev1));
    return RT_OK;
}

int rt_frame_times(rt_ctx* c, float* total_ms, float* kernel_ms) {
    int rc = rt_sync(c, total_ms);
    if (rc != RT_OK) return rc;
    if (kernel_ms) CU(c, cudaEventElapsedTime(kernel_ms, c->evk0, c->evk1));
    return RT_OK;
}

int rt_device_of(const rt_ctx* c, int* device) {
    if (!c || !device) return RT_ERR_ARG;
    *device = c->device;
    return RT_OK;
}

int rt_stream_handle(const rt_ctx* c, void** cuda_stream) {
    if (!c || !cuda_stream) return fail(nullptr, RT_ERR_ARG, "rt_stream_handle: NULL");
    *cuda_stream = (void*)c->stream;
    return RT_OK;
}

int rt_frame_stats(rt_ctx* c, uint64_t* node_visits, uint64_t* tri_tests, uint64_t* node_lines, uint64_t* tri_blocks) {
    if (!c) return fail(nullptr, RT_ERR_ARG, "rt_frame_stats: NULL ctx");
    CU(c, cudaSetDevice(c->device));
    if (!c->frame_valid) return fail(c, RT_ERR_STATE, "rt_frame_stats: no frame rendered");
    unsigned long long cnt[6] = {0, 0, 0, 0, 0, 0};
    CU(c, cudaMemcpyAsync(cnt, c->counters.p, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (node_visits) *node_visits = cnt[2];
    if (tri_tests) *tri_tests = cnt[3];
    if (node_lines) *node_lines = cnt[4];
    if (tri_blocks) *tri_blocks = cnt[5];
    return RT_OK;
}

int rt_download_image(rt_ctx* c, rt_image* img) {
    if (!c || !img) return fail(c, RT_ERR_ARG, "rt_download_image: NULL argument");
    CU(c, cudaSetDevice(c->device));
    if (!c->frame_valid) return fail(c, RT_ERR_STATE, "rt_download_image: no frame rendered");
    if (c->world == 1 && c->dbg_world > 1 && (img->rgb || img->rgb8 || img->tri_id || img->t))
        return fail(c, RT_ERR_STATE, "rt_download_image: rt_debug_set_shard frames are tile-packed and for timing only");
    const FrameParams& P = c->fp;
    const size_t npix = (size_t)P.W * P.H;
    const bool gathered = c->world > 1;
    if (!gathered || c->rank == 0) {
        struct Out { void* host; uint32_t bit; size_t bpp; const Plane* single; const Plane* multi; const char* name; };
        Out outs[4] = {{img->rgb, RT_OUT_RGB_F32, 12, &c->loc_rgb, &c->img_rgb, "rgb"}, {img->rgb8, RT_OUT_RGB8, 3, &c->loc_rgb8, &c->img_rgb8, "rgb8"},
                       {img->tri_id, RT_OUT_TRI_ID, 4, &c->loc_id, &c->img_id, "tri_id"}, {img->t, RT_OUT_T, 4, &c->loc_t, &c->img_t, "t"}};
        for (const Out& o : outs) {
            if (!o.host) continue;
            if (!(c->outputs & o.bit)) return fail(c, RT_ERR_STATE, "rt_download_image: plane '%s' was not requested in rt_frame.outputs", o.name);
            const Plane* src = gathered ? o.multi : o.single;
            CU(c, cudaMemcpyAsync(o.host, src->p, o.bpp * npix, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    unsigned long long* cnt = c->pin;                          // page-locked: truly asynchronous small reads
    c->pin[4] = 0ull;
    CU(c, cudaMemcpyAsync(cnt, c->counters.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    if (c->peer && c->peer_err) CU(c, cudaMemcpyAsync(&c->pin[4], c->peer_err, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    const unsigned perr = (unsigned)c->pin[4];
    if (perr) return peer_timeout(c, perr);
    img->width = P.W; img->height = P.H;
    img->rays_primary = cnt[0]; img->rays_shadow = cnt[1];
    float ms = 0.f;
    CU(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    img->gpu_ms = ms;
    if (c->inproc && c->group_root == c && (img->rgb || img->rgb8 || img->tri_id || img->t || !c->last_pumped)) {
        // single-process group, called by the user on the group: ray counts of every rank (rank 0's kernel ends after the others')
        for (rt_ctx* m : c->members) {
            rt_image meta{};
            int rc = rt_download_image(m, &meta);
            if (rc != RT_OK) return fail(c, rc, "rt_download_image: rank %d: %s", m->rank, m->err.c_str());
            img->rays_primary += meta.rays_primary; img->rays_shadow += meta.rays_shadow;
        }
        cudaSetDevice(c->device);
    }
    return RT_OK;
}

int rt_debug_set_shard(rt_ctx* c, int rank, int world) {
    if (!c || world < 0 || (world > 0 && (rank < 0 || rank >= world))) return fail(c, RT_ERR_ARG, "rt_debug_set_shard: bad arguments");
    c->dbg_rank = rank; c->dbg_world = world;
    c->frame_valid = false;
    return RT_OK;
}

int rt_debug_download_bvh(rt_ctx* c, void* nodes64, void* tri_blocks48, int32_t* tri_ids) {
    if (!c) return fail(nullptr, RT_ERR_ARG, "rt_debug_download_bvh: NULL ctx");
    CU(c, cudaSetDevice(c->device));
    if (!c->has_scene) return fail(c, RT_ERR_STATE, "rt_debug_download_bvh: no scene uploaded");
    CU(c, cudaStreamSynchronize(c->stream));
    if (nodes64 && c->num_nodes) CU(c, cudaMemcpy(nodes64, c->nodes, sizeof(BvhNode) * (size_t)c->num_nodes, cudaMemcpyDeviceToHost));
    if (tri_blocks48) CU(c, cudaMemcpy(tri_blocks48, c->geom, sizeof(TriBlock) * (size_t)c->num_tris, cudaMemcpyDeviceToHost));
    if (tri_ids) {
        std::vector<TriBlock> tmp(c->num_tris);
        CU(c, cudaMemcpy(tmp.data(), c->geom, sizeof(TriBlock) * (size_t)c->num_tris, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < c->num_tris; ++i) memcpy(&tri_ids[i], &tmp[i].g[3], 4);
    }
    return RT_OK;
}

} // extern "C"
