// rt_obj_device.cu — OBJ ingest on the device.
//
// Replaces LoadOBJ_ToMesh (HW2/HW2/GPUandCPU/include/MeshOBJ.h:260-427; same stream as HW1/src/MeshOBJ.cpp:143-281) for
// callers that want the mesh to exist only in HBM: the file's BYTES are copied to the device and parsed there; the
// arrays the reference loader would have produced on the host — vertices de-duplicated by their (v, vt, vn) reference in
// order of first use, quads split (0,1,2),(0,2,3), relative indices, per-object ids from o/g tags — come out as device
// arrays that rt_upload_scene takes as they are.  Not a port of the loader's loop (which is inherently sequential: a
// hash map filled line by line): lines are found with a flag + select, classified one thread per line, every
// "so far" quantity of the sequential loader (vertex counts for relative indices, tags seen, triangles emitted) is an
// exclusive scan over the lines, and first-use de-duplication is two stable radix sorts over the face corners plus a
// sort of the groups by their first occurrence.
//
// Numbers: strtof semantics.  A decimal literal with <= 19 significant digits whose value is m x 10^e with m < 2^53 and
// |e| <= 22 converts exactly in one fp64 operation; narrowing that to fp32 is correctly rounded unless the fp64 value sits
// exactly on an fp32 rounding boundary.  Those (and longer literals, huge exponents, sub-normal results) are "hard": the device
// records the token and the host converts just those tokens with strtof itself (the caller's buffer is host memory).  inf /
// nan / hexadecimal literals and lines of 1024 bytes or more (the loader's fgets buffer would split them) are refused.
// The host restatement of the same loader (host/mesh_ingest.cpp, checked against the reference's loader output in
// tests/golden/*_mesh.npz) is the oracle: tests/test_gpu_ingest.py compares every array bit for bit.
#include "../../include/rt_api.h"
#include "rt_obj_core.h"

#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

struct rt_dmesh {
    int device = 0;
    float* positions = nullptr;      // 3 per vertex
    float* normals = nullptr;        // 3 per vertex or nullptr
    uint32_t* indices = nullptr;     // 3 per triangle
    int32_t* tri_obj_ids = nullptr;  // 1 per triangle
    uint64_t nv = 0, nt = 0;
    float parse_ms = 0.f;
    uint64_t hard_numbers = 0, lines = 0;
};

namespace {

thread_local std::string g_err;

enum : uint8_t { LT_NONE = 0, LT_TAG, LT_V, LT_VT, LT_VN, LT_F };
enum : uint32_t { E_BAD_V = 1, E_BAD_VT, E_BAD_VN, E_FACE_LT3, E_MISSING_VERTEX, E_LONG_LINE, E_EXOTIC_NUMBER };
const char* const kErrText[] = {"", "bad 'v' line", "bad 'vt' line", "bad 'vn' line", "face with fewer than 3 vertices",
                                "face references a missing vertex", "line of 1024 bytes or more (not supported by the device parser)",
                                "inf / nan / hexadecimal number (not supported by the device parser)"};

struct LineCounts { uint32_t v, vt, vn, tag, tri, call; };
struct AddCounts {
    __host__ __device__ LineCounts operator()(const LineCounts& a, const LineCounts& b) const {
        return LineCounts{a.v + b.v, a.vt + b.vt, a.vn + b.vn, a.tag + b.tag, a.tri + b.tri, a.call + b.call};
    }
};
struct HardToken { uint32_t start, len, array, index; };       // array: 0 = raw positions, 1 = raw normals
struct Globals {
    unsigned long long err;          // min over (line << 8 | kind)
    uint32_t first_tag_line, first_face_line, first_n_event_line, first_t_event_line;
    uint32_t num_hard;
};

__global__ void k_mark_lines(const char* __restrict__ text, uint32_t n, uint8_t* __restrict__ flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || text[i - 1] == '\n') ? 1 : 0;
}

__device__ inline void report(Globals* g, uint32_t line, uint32_t kind) { atomicMin(&g->err, ((unsigned long long)line << 8) | kind); }

// One thread per line: what the loader's if-chain would do with it, and how much it adds to every running count.
__global__ void k_classify(const char* __restrict__ text, uint32_t n, const uint32_t* __restrict__ line_start, uint32_t nlines,
                           uint8_t* __restrict__ types, LineCounts* __restrict__ counts, Globals* g) {
    const uint32_t L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= nlines) return;
    const uint32_t s = line_start[L], e = L + 1 < nlines ? line_start[L + 1] : n;
    LineCounts lc{0, 0, 0, 0, 0, 0};
    uint8_t ty = LT_NONE;
    if (e - s >= 1024u) report(g, L, E_LONG_LINE);
    DCur c{text, s, e};
    c.ws();
    if (!(c.eol() || c.c() == '#')) {
        const char a = c.c(), b = c.at(c.i + 1), d = c.at(c.i + 2);
        if (a == 'o' || a == 'g') { ty = LT_TAG; lc.tag = 1; atomicMin(&g->first_tag_line, L); }
        else if (a == 'v' && (b == ' ' || b == '\t')) { ty = LT_V; lc.v = 1; }
        else if (a == 'v' && b == 't' && (d == ' ' || d == '\t')) { ty = LT_VT; lc.vt = 1; }
        else if (a == 'v' && b == 'n' && (d == ' ' || d == '\t')) { ty = LT_VN; lc.vn = 1; }
        else if (a == 'f' && (b == ' ' || b == '\t')) {
            ty = LT_F;
            c.i += 1;
            Corner k[4];
            const int nc = parse_face(c, k, 0u, 0u, 0u);        // the cursor movements do not depend on the counts
            if (nc < 3) report(g, L, E_FACE_LT3);
            else { lc.tri = nc == 4 ? 2u : 1u; lc.call = (uint32_t)nc; atomicMin(&g->first_face_line, L); }
        }
    }
    types[L] = ty;
    counts[L] = lc;
}

struct ParseOut {
    float* rp; float* rn;                      // raw v / vn values
    int* cv; int* ct; int* cn; uint32_t* cnn;  // per vertex() call: the key and the number of vn lines seen at that point
    uint32_t* tri_call; uint8_t* tri_kind; uint32_t* tri_tags;
    HardToken* hard; uint32_t hard_cap;
};

__device__ inline bool store_real(DCur& c, float* dst, uint32_t array, uint32_t index, Globals* g, const ParseOut& o, uint32_t L, uint32_t errkind) {
    float v = 0.f; uint32_t tok0 = 0;
    const int r = parse_real(c, v, tok0);
    if (r == 0) { report(g, L, errkind); return false; }
    if (r == 3) { report(g, L, E_EXOTIC_NUMBER); return false; }
    if (r == 2) {
        const uint32_t h = atomicAdd(&g->num_hard, 1u);
        if (h < o.hard_cap) o.hard[h] = HardToken{tok0, c.i - tok0, array, index};
    }
    if (dst) dst[index] = v;
    return true;
}

// One thread per line, second pass: values and face corners, placed by the exclusive counts of the lines before.
__global__ void k_parse(const char* __restrict__ text, uint32_t n, const uint32_t* __restrict__ line_start, uint32_t nlines,
                        const uint8_t* __restrict__ types, const LineCounts* __restrict__ pref, ParseOut o, Globals* g) {
    const uint32_t L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= nlines) return;
    const uint8_t ty = types[L];
    if (ty == LT_NONE || ty == LT_TAG) return;
    const uint32_t s = line_start[L], e = L + 1 < nlines ? line_start[L + 1] : n;
    const LineCounts pc = pref[L];
    DCur c{text, s, e};
    c.ws();
    if (ty == LT_V) {
        c.i += 1;
        for (uint32_t k = 0; k < 3; ++k) if (!store_real(c, o.rp, 0u, 3u * pc.v + k, g, o, L, E_BAD_V)) return;
    } else if (ty == LT_VN) {
        c.i += 2;
        atomicMin(&g->first_n_event_line, L);
        for (uint32_t k = 0; k < 3; ++k) if (!store_real(c, o.rn, 1u, 3u * pc.vn + k, g, o, L, E_BAD_VN)) return;
    } else if (ty == LT_VT) {
        c.i += 2;
        atomicMin(&g->first_t_event_line, L);
        for (uint32_t k = 0; k < 2; ++k) if (!store_real(c, nullptr, 2u, 0u, g, o, L, E_BAD_VT)) return;   // (texture coordinates are validated, not kept)
    } else {
        c.i += 1;
        Corner k[4];
        const int nc = parse_face(c, k, pc.v, pc.vt, pc.vn);
        if (nc < 3) return;                                   // (reported by k_classify)
        bool evn = false, evt = false;
        for (int j = 0; j < nc; ++j) {
            evn = evn || k[j].n >= 0; evt = evt || k[j].t >= 0;
            if (k[j].v < 0 || (uint32_t)k[j].v >= pc.v) report(g, L, E_MISSING_VERTEX);
            o.cv[pc.call + j] = k[j].v; o.ct[pc.call + j] = k[j].t; o.cn[pc.call + j] = k[j].n; o.cnn[pc.call + j] = pc.vn;
        }
        if (evn) atomicMin(&g->first_n_event_line, L);
        if (evt) atomicMin(&g->first_t_event_line, L);
        o.tri_call[pc.tri] = pc.call; o.tri_kind[pc.tri] = 0; o.tri_tags[pc.tri] = pc.tag;
        if (nc == 4) { o.tri_call[pc.tri + 1] = pc.call; o.tri_kind[pc.tri + 1] = 1; o.tri_tags[pc.tri + 1] = pc.tag; }
    }
}

__global__ void k_tkeys(const int* __restrict__ ct, uint32_t S, uint32_t* __restrict__ key, uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) { key[i] = (uint32_t)(ct[i] + 1); idx[i] = i; }
}
__global__ void k_vnkeys(const int* __restrict__ cv, const int* __restrict__ cn, const uint32_t* __restrict__ idx, uint32_t S, unsigned long long* __restrict__ key) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < S) { const uint32_t i = idx[p]; key[p] = ((unsigned long long)(uint32_t)cv[i] << 32) | (uint32_t)(cn[i] + 1); }
}
__global__ void k_heads(const int* __restrict__ cv, const int* __restrict__ ct, const int* __restrict__ cn, const uint32_t* __restrict__ order, uint32_t S, uint32_t* __restrict__ head) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= S) return;
    if (p == 0) { head[0] = 1u; return; }
    const uint32_t a = order[p], b = order[p - 1];
    head[p] = (cv[a] != cv[b] || ct[a] != ct[b] || cn[a] != cn[b]) ? 1u : 0u;
}
// gid = inclusive scan of head - 1.  The sorts are stable and started from increasing call order, so the head of a group is
// its first use.
__global__ void k_first_use(const uint32_t* __restrict__ order, const uint32_t* __restrict__ head, const uint32_t* __restrict__ gscan, uint32_t S,
                            uint32_t* __restrict__ first_use, uint32_t* __restrict__ giota) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < S && head[p]) { const uint32_t g = gscan[p] - 1u; first_use[g] = order[p]; giota[g] = g; }
}
__global__ void k_vid(const uint32_t* __restrict__ gsorted, uint32_t G, uint32_t* __restrict__ vid) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < G) vid[gsorted[r]] = r;
}
__global__ void k_call_vid(const uint32_t* __restrict__ order, const uint32_t* __restrict__ gscan, const uint32_t* __restrict__ vid, uint32_t S, uint32_t* __restrict__ call_vid) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < S) call_vid[order[p]] = vid[gscan[p] - 1u];
}
// Vertex r = the group whose first use comes r-th: the raw position, and the raw normal if the file had that many vn
// lines WHEN THE VERTEX WAS CREATED (the loader looks it up at the first use), else zero.
__global__ void k_vertices(const uint32_t* __restrict__ first_sorted, uint32_t G, const int* __restrict__ cv, const int* __restrict__ cn, const uint32_t* __restrict__ cnn,
                           const float* __restrict__ rp, const float* __restrict__ rn, float* __restrict__ pos, float* __restrict__ nrm) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= G) return;
    const uint32_t call = first_sorted[r];
    const int v = cv[call], nn = cn[call];
    pos[3 * r] = rp[3 * v]; pos[3 * r + 1] = rp[3 * v + 1]; pos[3 * r + 2] = rp[3 * v + 2];
    if (nrm) {
        const bool ok = nn >= 0 && (uint32_t)nn < cnn[call];
        nrm[3 * r] = ok ? rn[3 * nn] : 0.f; nrm[3 * r + 1] = ok ? rn[3 * nn + 1] : 0.f; nrm[3 * r + 2] = ok ? rn[3 * nn + 2] : 0.f;
    }
}
__global__ void k_triangles(const uint32_t* __restrict__ tri_call, const uint8_t* __restrict__ tri_kind, const uint32_t* __restrict__ tri_tags, uint32_t T,
                            const uint32_t* __restrict__ call_vid, int32_t base_id, uint32_t first_tag_discount, uint32_t* __restrict__ idx, int32_t* __restrict__ obj) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const uint32_t c = tri_call[t];
    const bool second = tri_kind[t] != 0;
    idx[3 * t] = call_vid[c]; idx[3 * t + 1] = call_vid[c + (second ? 2u : 1u)]; idx[3 * t + 2] = call_vid[c + (second ? 3u : 2u)];
    const uint32_t tags = tri_tags[t];
    obj[t] = base_id + (int32_t)(tags == 0u ? 0u : tags - first_tag_discount);
}
__global__ void k_rebase(uint32_t* __restrict__ idx, uint64_t n, uint32_t base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] += base;
}

inline unsigned blocks_for(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

struct Scratch {          // stream-ordered scratch, everything freed on every exit path
    cudaStream_t stream = nullptr;
    std::vector<void*> ptrs;
    template <typename T> cudaError_t alloc(T** p, size_t count) {
        *p = nullptr;
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(p), (count ? count : 1) * sizeof(T), stream);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, stream); }
};

int fail(int code, const std::string& why) { g_err = why; return code; }

} // namespace

extern "C" {

const char* rt_dmesh_last_error(void) { return g_err.c_str(); }

void rt_dmesh_free(rt_dmesh* m) {
    if (!m) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(m->device);
    cudaFree(m->positions); cudaFree(m->normals); cudaFree(m->indices); cudaFree(m->tri_obj_ids);
    cudaSetDevice(cur);
    delete m;
}

int rt_dmesh_parse_obj(rt_ctx* ctx, const char* text, uint64_t nbytes, int32_t* next_object_id, rt_dmesh** out) {
    if (!ctx || !text || !out) return fail(RT_ERR_ARG, "rt_dmesh_parse_obj: NULL argument");
    *out = nullptr;
    if (nbytes == 0) return fail(RT_ERR_ARG, "rt_dmesh_parse_obj: no geometry");
    if (nbytes >= (1ull << 31)) return fail(RT_ERR_UNSUPPORTED, "rt_dmesh_parse_obj: files of 2 GiB or more are not supported");
    int device = 0;
    void* sh = nullptr;
    if (rt_device_of(ctx, &device) != RT_OK || rt_stream_handle(ctx, &sh) != RT_OK) return fail(RT_ERR_ARG, "rt_dmesh_parse_obj: bad context");
    cudaStream_t stream = (cudaStream_t)sh;
    const uint32_t n = (uint32_t)nbytes;
    const unsigned T = 256;
    Scratch sc;
    sc.stream = stream;
    cudaError_t ce = cudaSuccess;
#define CK(x) do { ce = (x); if (ce != cudaSuccess) return fail(RT_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(ce)); } while (0)
    CK(cudaSetDevice(device));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{e0, e1};
    CK(cudaEventRecord(e0, stream));

    // ---- bytes to HBM, line starts
    char* d_text = nullptr; uint8_t* d_flag = nullptr; uint32_t* d_lines = nullptr; uint32_t* d_nlines = nullptr; Globals* d_g = nullptr;
    void* d_tmp = nullptr; size_t tmp_bytes = 0;
    CK(sc.alloc(&d_text, n)); CK(sc.alloc(&d_flag, n)); CK(sc.alloc(&d_lines, n)); CK(sc.alloc(&d_nlines, 1)); CK(sc.alloc(&d_g, 1));
    CK(cudaMemcpyAsync(d_text, text, n, cudaMemcpyHostToDevice, stream));
    Globals g0; memset(&g0, 0, sizeof g0);
    g0.err = ~0ull; g0.first_tag_line = g0.first_face_line = g0.first_n_event_line = g0.first_t_event_line = 0xffffffffu;
    CK(cudaMemcpyAsync(d_g, &g0, sizeof g0, cudaMemcpyHostToDevice, stream));
    k_mark_lines<<<blocks_for(n, T), T, 0, stream>>>(d_text, n, d_flag);
    cub::CountingInputIterator<uint32_t> iota(0u);
    CK(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, d_flag, d_lines, d_nlines, (int)n, stream));
    CK(sc.alloc(reinterpret_cast<char**>(&d_tmp), tmp_bytes));
    CK(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, iota, d_flag, d_lines, d_nlines, (int)n, stream));
    uint32_t nlines = 0;
    CK(cudaMemcpyAsync(&nlines, d_nlines, sizeof nlines, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));

    // ---- classify, running counts
    uint8_t* d_types = nullptr; LineCounts *d_counts = nullptr, *d_pref = nullptr;
    CK(sc.alloc(&d_types, nlines)); CK(sc.alloc(&d_counts, nlines)); CK(sc.alloc(&d_pref, nlines));
    k_classify<<<blocks_for(nlines, T), T, 0, stream>>>(d_text, n, d_lines, nlines, d_types, d_counts, d_g);
    void* d_tmp2 = nullptr; size_t tmp2 = 0;
    CK(cub::DeviceScan::ExclusiveScan(nullptr, tmp2, d_counts, d_pref, AddCounts(), LineCounts{0, 0, 0, 0, 0, 0}, (int)nlines, stream));
    CK(sc.alloc(reinterpret_cast<char**>(&d_tmp2), tmp2));
    CK(cub::DeviceScan::ExclusiveScan(d_tmp2, tmp2, d_counts, d_pref, AddCounts(), LineCounts{0, 0, 0, 0, 0, 0}, (int)nlines, stream));
    LineCounts lastp{}, lastc{};
    Globals g1;
    CK(cudaMemcpyAsync(&lastp, d_pref + (nlines - 1), sizeof lastp, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&lastc, d_counts + (nlines - 1), sizeof lastc, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&g1, d_g, sizeof g1, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const LineCounts tot = AddCounts()(lastp, lastc);
    uint32_t tris_before_first_tag = 0;
    if (g1.first_tag_line != 0xffffffffu) {
        LineCounts p{};
        CK(cudaMemcpyAsync(&p, d_pref + g1.first_tag_line, sizeof p, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        tris_before_first_tag = p.tri;
    }

    // ---- values and face corners
    const uint32_t S = tot.call, NT = tot.tri;
    ParseOut po{};
    const uint32_t hard_cap = 1u << 16;
    CK(sc.alloc(&po.rp, 3 * (size_t)tot.v)); CK(sc.alloc(&po.rn, 3 * (size_t)tot.vn));
    CK(sc.alloc(&po.cv, S)); CK(sc.alloc(&po.ct, S)); CK(sc.alloc(&po.cn, S)); CK(sc.alloc(&po.cnn, S));
    CK(sc.alloc(&po.tri_call, NT)); CK(sc.alloc(&po.tri_kind, NT)); CK(sc.alloc(&po.tri_tags, NT));
    CK(sc.alloc(&po.hard, hard_cap));
    po.hard_cap = hard_cap;
    k_parse<<<blocks_for(nlines, T), T, 0, stream>>>(d_text, n, d_lines, nlines, d_types, d_pref, po, d_g);
    CK(cudaMemcpyAsync(&g1, d_g, sizeof g1, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (g1.err != ~0ull) {
        const uint32_t kind = (uint32_t)(g1.err & 0xffu);
        return fail(kind == E_LONG_LINE || kind == E_EXOTIC_NUMBER ? RT_ERR_UNSUPPORTED : RT_ERR_ARG,
                    std::string("rt_dmesh_parse_obj: line ") + std::to_string((unsigned long long)(g1.err >> 8) + 1ull) + ": " + kErrText[kind]);
    }
    if (NT == 0 || S == 0) return fail(RT_ERR_ARG, "rt_dmesh_parse_obj: no geometry");
    if (g1.num_hard > hard_cap) return fail(RT_ERR_UNSUPPORTED, "rt_dmesh_parse_obj: more than 65536 numbers need the host's strtof");
    if (g1.num_hard) {               // the literals fp64 cannot decide: the host's strtof on the caller's own buffer
        std::vector<HardToken> hard(g1.num_hard);
        CK(cudaMemcpyAsync(hard.data(), po.hard, sizeof(HardToken) * g1.num_hard, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (const HardToken& h : hard) {
            if (h.array > 1u) continue;
            const std::string tok(text + h.start, text + h.start + h.len);
            const float v = strtof(tok.c_str(), nullptr);
            CK(cudaMemcpyAsync((h.array == 0u ? po.rp : po.rn) + h.index, &v, sizeof v, cudaMemcpyHostToDevice, stream));
            CK(cudaStreamSynchronize(stream));
        }
    }
    // streams the loader keeps aligned with the vertices: a stream that appears after the first vertex was created is an error
    const bool has_nrm = g1.first_n_event_line != 0xffffffffu, has_uv = g1.first_t_event_line != 0xffffffffu;
    if (has_uv && g1.first_face_line < g1.first_t_event_line) return fail(RT_ERR_ARG, "rt_dmesh_parse_obj: uv stream misaligned");
    if (has_nrm && g1.first_face_line < g1.first_n_event_line) return fail(RT_ERR_ARG, "rt_dmesh_parse_obj: normal stream misaligned");

    // ---- first-use de-duplication of the (v, vt, vn) keys
    uint32_t *d_k32 = nullptr, *d_k32b = nullptr, *d_idx = nullptr, *d_idx1 = nullptr, *d_order = nullptr, *d_head = nullptr, *d_gscan = nullptr;
    unsigned long long *d_k64 = nullptr, *d_k64b = nullptr;
    uint32_t *d_first = nullptr, *d_giota = nullptr, *d_first_sorted = nullptr, *d_gsorted = nullptr, *d_vid = nullptr, *d_call_vid = nullptr;
    CK(sc.alloc(&d_k32, S)); CK(sc.alloc(&d_k32b, S)); CK(sc.alloc(&d_idx, S)); CK(sc.alloc(&d_idx1, S)); CK(sc.alloc(&d_order, S));
    CK(sc.alloc(&d_head, S)); CK(sc.alloc(&d_gscan, S)); CK(sc.alloc(&d_k64, S)); CK(sc.alloc(&d_k64b, S));
    CK(sc.alloc(&d_first, S)); CK(sc.alloc(&d_giota, S)); CK(sc.alloc(&d_first_sorted, S)); CK(sc.alloc(&d_gsorted, S)); CK(sc.alloc(&d_vid, S)); CK(sc.alloc(&d_call_vid, S));
    size_t t1 = 0, t2 = 0, t3 = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, t1, d_k32, d_k32b, d_idx, d_idx1, (int)S, 0, 32, stream));
    CK(cub::DeviceRadixSort::SortPairs(nullptr, t2, d_k64, d_k64b, d_idx1, d_order, (int)S, 0, 64, stream));
    CK(cub::DeviceScan::InclusiveSum(nullptr, t3, d_head, d_gscan, (int)S, stream));
    size_t tmax = t1 > t2 ? t1 : t2; if (t3 > tmax) tmax = t3;
    void* d_tmp3 = nullptr;
    CK(sc.alloc(reinterpret_cast<char**>(&d_tmp3), tmax));
    k_tkeys<<<blocks_for(S, T), T, 0, stream>>>(po.ct, S, d_k32, d_idx);
    CK(cub::DeviceRadixSort::SortPairs(d_tmp3, tmax, d_k32, d_k32b, d_idx, d_idx1, (int)S, 0, 32, stream));
    k_vnkeys<<<blocks_for(S, T), T, 0, stream>>>(po.cv, po.cn, d_idx1, S, d_k64);
    CK(cub::DeviceRadixSort::SortPairs(d_tmp3, tmax, d_k64, d_k64b, d_idx1, d_order, (int)S, 0, 64, stream));
    k_heads<<<blocks_for(S, T), T, 0, stream>>>(po.cv, po.ct, po.cn, d_order, S, d_head);
    CK(cub::DeviceScan::InclusiveSum(d_tmp3, tmax, d_head, d_gscan, (int)S, stream));
    uint32_t G = 0;
    CK(cudaMemcpyAsync(&G, d_gscan + (S - 1), sizeof G, cudaMemcpyDeviceToHost, stream));
    k_first_use<<<blocks_for(S, T), T, 0, stream>>>(d_order, d_head, d_gscan, S, d_first, d_giota);
    CK(cudaStreamSynchronize(stream));
    CK(cub::DeviceRadixSort::SortPairs(d_tmp3, tmax, d_first, d_first_sorted, d_giota, d_gsorted, (int)G, 0, 32, stream));
    k_vid<<<blocks_for(G, T), T, 0, stream>>>(d_gsorted, G, d_vid);
    k_call_vid<<<blocks_for(S, T), T, 0, stream>>>(d_order, d_gscan, d_vid, S, d_call_vid);

    // ---- the mesh
    rt_dmesh* m = new (std::nothrow) rt_dmesh;
    if (!m) return fail(RT_ERR_NOMEM, "rt_dmesh_parse_obj: out of memory");
    m->device = device; m->nv = G; m->nt = NT; m->lines = nlines; m->hard_numbers = g1.num_hard;
#define CKM(x) do { ce = (x); if (ce != cudaSuccess) { rt_dmesh_free(m); return fail(RT_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(ce)); } } while (0)
    CKM(cudaMalloc(&m->positions, sizeof(float) * 3 * (size_t)G));
    if (has_nrm) CKM(cudaMalloc(&m->normals, sizeof(float) * 3 * (size_t)G));
    CKM(cudaMalloc(&m->indices, sizeof(uint32_t) * 3 * (size_t)NT));
    CKM(cudaMalloc(&m->tri_obj_ids, sizeof(int32_t) * (size_t)NT));
    const int32_t base_id = next_object_id ? *next_object_id : 0;
    // every tag after the first starts a new object; the first one does too if faces were already emitted under the implicit one
    const uint32_t discount = tris_before_first_tag == 0u ? 1u : 0u;
    k_vertices<<<blocks_for(G, T), T, 0, stream>>>(d_first_sorted, G, po.cv, po.cn, po.cnn, po.rp, po.rn, m->positions, m->normals);
    k_triangles<<<blocks_for(NT, T), T, 0, stream>>>(po.tri_call, po.tri_kind, po.tri_tags, NT, d_call_vid, base_id, discount, m->indices, m->tri_obj_ids);
    CKM(cudaGetLastError());
    CKM(cudaEventRecord(e1, stream));
    CKM(cudaStreamSynchronize(stream));
    cudaEventElapsedTime(&m->parse_ms, e0, e1);
    if (next_object_id) *next_object_id = base_id + (int32_t)(tot.tag == 0u ? 0u : tot.tag - discount) + 1;
#undef CKM
#undef CK
    *out = m;
    return RT_OK;
}

int rt_dmesh_counts(const rt_dmesh* m, uint64_t* num_vertices, uint64_t* num_normals, uint64_t* num_triangles) {
    if (!m) return RT_ERR_ARG;
    if (num_vertices) *num_vertices = m->nv;
    if (num_normals) *num_normals = m->normals ? m->nv : 0;
    if (num_triangles) *num_triangles = m->nt;
    return RT_OK;
}

int rt_dmesh_arrays(const rt_dmesh* m, const float** positions, const float** normals, const uint32_t** indices, const int32_t** tri_obj_ids) {
    if (!m) return RT_ERR_ARG;
    if (positions) *positions = m->positions;
    if (normals) *normals = m->normals;
    if (indices) *indices = m->indices;
    if (tri_obj_ids) *tri_obj_ids = m->tri_obj_ids;
    return RT_OK;
}

int rt_dmesh_copy(const rt_dmesh* m, float* positions, float* normals, uint32_t* indices, int32_t* tri_obj_ids) {
    if (!m) return RT_ERR_ARG;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(m->device);
    cudaError_t e = cudaSuccess;
    if (positions && e == cudaSuccess) e = cudaMemcpy(positions, m->positions, sizeof(float) * 3 * m->nv, cudaMemcpyDeviceToHost);
    if (normals && m->normals && e == cudaSuccess) e = cudaMemcpy(normals, m->normals, sizeof(float) * 3 * m->nv, cudaMemcpyDeviceToHost);
    if (indices && e == cudaSuccess) e = cudaMemcpy(indices, m->indices, sizeof(uint32_t) * 3 * m->nt, cudaMemcpyDeviceToHost);
    if (tri_obj_ids && e == cudaSuccess) e = cudaMemcpy(tri_obj_ids, m->tri_obj_ids, sizeof(int32_t) * m->nt, cudaMemcpyDeviceToHost);
    cudaSetDevice(cur);
    return e == cudaSuccess ? RT_OK : fail(RT_ERR_CUDA, std::string("rt_dmesh_copy: ") + cudaGetErrorString(e));
}

int rt_dmesh_stats(const rt_dmesh* m, float* parse_ms, uint64_t* lines, uint64_t* host_converted_numbers) {
    if (!m) return RT_ERR_ARG;
    if (parse_ms) *parse_ms = m->parse_ms;
    if (lines) *lines = m->lines;
    if (host_converted_numbers) *host_converted_numbers = m->hard_numbers;
    return RT_OK;
}

// AppendMesh (MeshOBJ.h:429-466) on the device: indices re-based, a missing normal stream zero-filled.
int rt_dmesh_append(rt_dmesh* dst, const rt_dmesh* src) {
    if (!dst || !src || dst == src) return fail(RT_ERR_ARG, "rt_dmesh_append: bad arguments");
    if (dst->nv && dst->device != src->device) return fail(RT_ERR_ARG, "rt_dmesh_append: meshes live on different devices");
    if (dst->nv + src->nv >= (1ull << 32)) return fail(RT_ERR_ARG, "rt_dmesh_append: more than 2^32 vertices");
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(src->device);
    dst->device = src->device;
    const uint64_t nv = dst->nv + src->nv, nt = dst->nt + src->nt;
    const bool nrm = dst->normals || src->normals;
    float *pos = nullptr, *nr = nullptr; uint32_t* idx = nullptr; int32_t* obj = nullptr;
    cudaError_t e = cudaMalloc(&pos, sizeof(float) * 3 * nv);
    if (e == cudaSuccess && nrm) e = cudaMalloc(&nr, sizeof(float) * 3 * nv);
    if (e == cudaSuccess) e = cudaMalloc(&idx, sizeof(uint32_t) * 3 * nt);
    if (e == cudaSuccess) e = cudaMalloc(&obj, sizeof(int32_t) * nt);
    if (e == cudaSuccess && nrm) e = cudaMemset(nr, 0, sizeof(float) * 3 * nv);
    if (e == cudaSuccess && dst->nv) e = cudaMemcpy(pos, dst->positions, sizeof(float) * 3 * dst->nv, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pos + 3 * dst->nv, src->positions, sizeof(float) * 3 * src->nv, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess && dst->normals) e = cudaMemcpy(nr, dst->normals, sizeof(float) * 3 * dst->nv, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess && src->normals) e = cudaMemcpy(nr + 3 * dst->nv, src->normals, sizeof(float) * 3 * src->nv, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess && dst->nt) e = cudaMemcpy(idx, dst->indices, sizeof(uint32_t) * 3 * dst->nt, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(idx + 3 * dst->nt, src->indices, sizeof(uint32_t) * 3 * src->nt, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess && dst->nt) e = cudaMemcpy(obj, dst->tri_obj_ids, sizeof(int32_t) * dst->nt, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(obj + dst->nt, src->tri_obj_ids, sizeof(int32_t) * src->nt, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess && dst->nv) {
        k_rebase<<<blocks_for(3 * src->nt, 256), 256>>>(idx + 3 * dst->nt, 3 * src->nt, (uint32_t)dst->nv);
        e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) {
        cudaFree(pos); cudaFree(nr); cudaFree(idx); cudaFree(obj);
        cudaSetDevice(cur);
        return fail(RT_ERR_CUDA, std::string("rt_dmesh_append: ") + cudaGetErrorString(e));
    }
    cudaFree(dst->positions); cudaFree(dst->normals); cudaFree(dst->indices); cudaFree(dst->tri_obj_ids);
    dst->positions = pos; dst->normals = nr; dst->indices = idx; dst->tri_obj_ids = obj; dst->nv = nv; dst->nt = nt;
    cudaSetDevice(cur);
    return RT_OK;
}

int rt_dmesh_create(rt_dmesh** out) {
    if (!out) return RT_ERR_ARG;
    *out = new (std::nothrow) rt_dmesh;
    return *out ? RT_OK : RT_ERR_NOMEM;
}

} // extern "C"
