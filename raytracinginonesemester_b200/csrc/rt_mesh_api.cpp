// rt_mesh_api.cpp — C ABI over the host mesh ingest (host/mesh_ingest.cpp): the reference loaders'
// output format (MeshOBJ.h:260-427, main.cu:57-96, MeshOBJ.h:429-466) for callers that do not link C++.
#include "../../include/rt_api.h"
#include "../host/mesh_ingest.hpp"

#include <cstring>
#include <new>
#include <string>

struct rt_mesh { rtb200::HostMesh m; };

namespace { thread_local std::string g_mesh_err; }

extern "C" {

const char* rt_mesh_last_error(void) { return g_mesh_err.c_str(); }

int rt_mesh_load_obj(const char* path, int32_t* next_object_id, rt_mesh** out) {
    if (!path || !out) { g_mesh_err = "rt_mesh_load_obj: NULL argument"; return RT_ERR_ARG; }
    *out = nullptr;
    rt_mesh* h = new (std::nothrow) rt_mesh;
    if (!h) return RT_ERR_NOMEM;
    int nid = next_object_id ? *next_object_id : 0;
    std::string err;
    if (!rtb200::load_obj(path, h->m, nid, &err)) { g_mesh_err = err; delete h; return RT_ERR_ARG; }
    if (next_object_id) *next_object_id = nid;
    *out = h;
    return RT_OK;
}

int rt_mesh_create(rt_mesh** out) {
    if (!out) return RT_ERR_ARG;
    *out = new (std::nothrow) rt_mesh;
    return *out ? RT_OK : RT_ERR_NOMEM;
}

void rt_mesh_free(rt_mesh* m) { delete m; }

int rt_mesh_counts(const rt_mesh* m, uint64_t* num_vertices, uint64_t* num_normals, uint64_t* num_triangles) {
    if (!m) return RT_ERR_ARG;
    if (num_vertices) *num_vertices = m->m.num_vertices();
    if (num_normals) *num_normals = m->m.normals.size() / 3;
    if (num_triangles) *num_triangles = m->m.num_triangles();
    return RT_OK;
}

int rt_mesh_copy(const rt_mesh* m, float* positions, float* normals, uint32_t* indices, int32_t* tri_obj_ids) {
    if (!m) return RT_ERR_ARG;
    if (positions) std::memcpy(positions, m->m.positions.data(), m->m.positions.size() * sizeof(float));
    if (normals) std::memcpy(normals, m->m.normals.data(), m->m.normals.size() * sizeof(float));
    if (indices) std::memcpy(indices, m->m.indices.data(), m->m.indices.size() * sizeof(uint32_t));
    if (tri_obj_ids) std::memcpy(tri_obj_ids, m->m.tri_obj_ids.data(), m->m.tri_obj_ids.size() * sizeof(int32_t));
    return RT_OK;
}

int rt_mesh_transform(rt_mesh* m, const float position[3], const float rotation_deg[3], const float scale[3]) {
    if (!m || !position || !rotation_deg || !scale) return RT_ERR_ARG;
    rtb200::transform_mesh(m->m, position, rotation_deg, scale);
    return RT_OK;
}

int rt_mesh_append(rt_mesh* dst, const rt_mesh* src) {
    if (!dst || !src) return RT_ERR_ARG;
    rtb200::append_mesh(dst->m, src->m);
    return RT_OK;
}

} // extern "C"
