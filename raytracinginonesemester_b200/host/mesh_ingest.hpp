// mesh_ingest.hpp — host-side mesh ingest for the C++ driver and the C ABI (rt_mesh_*).
// Produces exactly the arrays the reference loaders hand to their renderers:
//   LoadOBJ_ToMesh     HW2/HW2/GPUandCPU/include/MeshOBJ.h:260-427  (per-object ids from o/g tags)
//   LoadOBJ_ToMeshSOA  HW1/src/MeshOBJ.cpp:143-281                  (same vertex/index stream)
//   applyObjectTransform / AppendMesh   GPUandCPU/src/main.cu:57-96, MeshOBJ.h:429-466
// i.e. vertices de-duplicated by their (v, vt, vn) reference in order of first use, quads split
// as (0,1,2),(0,2,3), negative (relative) indices, strtof number parsing.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace rtb200 {

struct HostMesh {
    std::vector<float> positions;     // 3 per vertex
    std::vector<float> normals;       // 3 per vertex, or empty when the file has none
    std::vector<float> uvs;           // 2 per vertex, or empty
    std::vector<uint32_t> indices;    // 3 per triangle
    std::vector<int32_t> tri_obj_ids; // 1 per triangle
    size_t num_vertices() const { return positions.size() / 3; }
    size_t num_triangles() const { return indices.size() / 3; }
};

// next_object_id: in = id of the first object in the file, out = first unused id.
bool load_obj(const std::string& path, HostMesh& out, int& next_object_id, std::string* err);

// scale -> rotate X, Y, Z (degrees) -> translate; normals by inverse scale, re-normalised.
void transform_mesh(HostMesh& m, const float position[3], const float rotation_deg[3], const float scale[3]);

// Appends src to dst, re-basing indices and zero-filling missing normal/uv streams.
void append_mesh(HostMesh& dst, const HostMesh& src);

} // namespace rtb200
