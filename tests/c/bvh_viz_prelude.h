// Stand-ins for what HW2/HW2/GPUandCPU/src/main.cu has in scope where INTEGRATION.md §2 splices the C ABI in:
// the types are laid out like the reference's (Vec3 12 B, Material 13 floats material.h:6-20, Light 28 B scene.h:21-25,
// Mesh MeshOBJ.h:69-93 with `indices` holding 3 entries per triangle).  TEST INFRASTRUCTURE: lets the documented binding
// be compiled and run verbatim (tests/test_integration_snippet.py); nothing here is reference code.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

struct Vec3 { float x, y, z; };
struct Material { float albedo[3], kd, specular_color[3], ks, shininess, kr, emission[3]; };
struct Light { Vec3 position; Vec3 color; int intensity; };
struct Mesh {
    std::vector<Vec3> positions, normals;
    std::vector<uint32_t> indices;          // 3 per triangle
    std::vector<int32_t> triangleObjIds;    // 1 per triangle
};
struct CamCfg { int pixel_width, pixel_height; };

// scene file written by the test: nv, nt, W, H, spp (int32) | positions | indices | cam pos, look_at, up (9 f32) | focal, sensor (2 f64)
// | light (7 x 4 B) | miss (3 f32) | material (13 f32)
struct Loaded {
    Mesh globalMesh; std::vector<Material> objectMaterials; std::vector<Light> lights;
    CamCfg cam; Vec3 camPos, camLookAt, camUp, missColor; double focal_mm, sensor_mm; int spp;
};
inline Loaded load_scene_file(const char* path) {
    Loaded L;
    FILE* f = fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    int32_t h[5];
    if (fread(h, 4, 5, f) != 5) throw std::runtime_error("short scene file");
    L.globalMesh.positions.resize(h[0]); L.globalMesh.indices.resize(3 * (size_t)h[1]); L.globalMesh.triangleObjIds.assign(h[1], 0);
    L.cam.pixel_width = h[2]; L.cam.pixel_height = h[3]; L.spp = h[4];
    bool ok = fread(L.globalMesh.positions.data(), 12, h[0], f) == (size_t)h[0] && fread(L.globalMesh.indices.data(), 4, 3 * (size_t)h[1], f) == 3 * (size_t)h[1];
    float c[9]; double d[2]; Light li; float miss[3]; Material m;
    ok = ok && fread(c, 4, 9, f) == 9 && fread(d, 8, 2, f) == 2 && fread(&li, 28, 1, f) == 1 && fread(miss, 4, 3, f) == 3 && fread(&m, 52, 1, f) == 1;
    fclose(f);
    if (!ok) throw std::runtime_error("short scene file");
    L.camPos = {c[0], c[1], c[2]}; L.camLookAt = {c[3], c[4], c[5]}; L.camUp = {c[6], c[7], c[8]};
    L.focal_mm = d[0]; L.sensor_mm = d[1]; L.lights.push_back(li); L.missColor = {miss[0], miss[1], miss[2]}; L.objectMaterials.push_back(m);
    return L;
}
