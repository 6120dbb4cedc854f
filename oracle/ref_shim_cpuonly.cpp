// ref_shim_cpuonly.cpp — C entry points over the UNMODIFIED HW2/HW2/CPUOnly reference sources, compiled where they
// lie under /root/reference (oracle/build.py -> oracle/_ref/libref_cpuonly.so, git-ignored).
// TEST INFRASTRUCTURE ONLY: validates oracle/rt_oracle.c's RT_MODE_HW2_CPU restatement and produces the fixtures
// of tools/make_golden_cpuonly.py.  No reference source is copied: the headers are #included in place.
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#define private public          // camera keeps pixel00_loc / pixel_delta_u / pixel_delta_v private
#include "camera.h"
#undef private
#include "raytracer.h"
#include "transform.h"

extern "C" {

void* ref_cpu_load_obj(const char* path) {
    MeshSOA* m = new MeshSOA;
    if (!LoadOBJ_ToMeshSOA(path, *m)) { delete m; return nullptr; }
    return m;
}
void ref_cpu_mesh_counts(void* h, uint64_t* nv, uint64_t* nn, uint64_t* nt) {
    MeshSOA* m = (MeshSOA*)h;
    *nv = m->positions.size(); *nn = m->normals.size(); *nt = m->indices.size() / 3;
}
void ref_cpu_mesh_copy(void* h, float* pos, float* nrm, uint32_t* idx) {
    MeshSOA* m = (MeshSOA*)h;
    std::memcpy(pos, m->positions.data(), m->positions.size() * sizeof(Vec3));
    if (nrm && !m->normals.empty()) std::memcpy(nrm, m->normals.data(), m->normals.size() * sizeof(Vec3));
    std::memcpy(idx, m->indices.data(), m->indices.size() * sizeof(uint32_t));
}
// ApplyTransformToMeshSOA (transform.h:76-85) on the loaded mesh
void ref_cpu_mesh_transform(void* h, const float* position, const float* rotation_deg, const float* scale) {
    Transform t;
    t.position = make_vec3(position[0], position[1], position[2]);
    t.rotation_deg = make_vec3(rotation_deg[0], rotation_deg[1], rotation_deg[2]);
    t.scale = make_vec3(scale[0], scale[1], scale[2]);
    ApplyTransformToMeshSOA(*(MeshSOA*)h, t);
}
void ref_cpu_mesh_free(void* h) { delete (MeshSOA*)h; }

// camera::initialize (camera.h:64-104): out12 = center, pixel00_loc, pixel_delta_u, pixel_delta_v.  Returns 1 when it throws.
int ref_cpu_camera(const float* pos, const float* look, const float* up, double focal_mm, double sensor_h_mm,
                   double sensor_w_mm, int W, int H, float* out12) {
    try {
        camera cam(make_vec3(pos[0], pos[1], pos[2]), make_vec3(look[0], look[1], look[2]), make_vec3(up[0], up[1], up[2]),
                   focal_mm, sensor_h_mm, sensor_w_mm, W, H);
        const Vec3 v[4] = {cam.center, cam.pixel00_loc, cam.pixel_delta_u, cam.pixel_delta_v};
        std::memcpy(out12, v, sizeof v);
        return 0;
    } catch (const std::exception&) { return 1; }
}

struct ref_cpu_light { float position[3]; float color[3]; float intensity; };

// The pixel loop of src/render.cpp:118-139 at samples_per_pixel == 1 (pixel centre +0.5) over triangles built as in
// render.cpp:79-97 (per-object material; face normals when nrm == NULL), TraceRay with diffuse_bounce == false
// (mirror bounces only — the deterministic subset; the unused xi still draws from random_float()).
// tri_id / t: closest hit of the primary ray with IntersectScene's rule (t >= RT_EPS, strict <, first wins).
void ref_cpu_render_rows(const float* pos, const float* nrm, uint64_t nv, const uint32_t* idx, uint64_t nt, const int32_t* obj_ids,
                         const void* materials52, int num_materials, const float* cpos, const float* look, const float* up,
                         double focal_mm, double sensor_h_mm, double sensor_w_mm, int W, int H, const ref_cpu_light* lights_in, int num_lights,
                         int max_bounces, int row_begin, int row_step, float* rgb, int32_t* tri_id, float* tout)
{
    (void)nv;
    static_assert(sizeof(Material) == 52, "Material layout");
    const Vec3* P = (const Vec3*)pos; const Vec3* N = (const Vec3*)nrm;
    const Material* mats = (const Material*)materials52;
    std::vector<Triangle> tris;
    tris.reserve(nt);
    for (uint64_t k = 0; k < nt; ++k) {
        Triangle tri;
        const int o = obj_ids ? obj_ids[k] : 0;
        tri.mat = (o >= 0 && o < num_materials) ? mats[o] : Material{};
        tri.v0 = P[idx[3 * k]]; tri.v1 = P[idx[3 * k + 1]]; tri.v2 = P[idx[3 * k + 2]];
        if (N) { tri.n0 = N[idx[3 * k]]; tri.n1 = N[idx[3 * k + 1]]; tri.n2 = N[idx[3 * k + 2]]; }
        else { Vec3 faceN = unit_vector(cross(tri.v1 - tri.v0, tri.v2 - tri.v0)); tri.n0 = tri.n1 = tri.n2 = faceN; }
        tris.push_back(tri);
    }
    std::vector<Light> lights;
    for (int i = 0; i < num_lights; ++i) {
        Light L;
        L.position = make_vec3(lights_in[i].position[0], lights_in[i].position[1], lights_in[i].position[2]);
        L.color = make_vec3(lights_in[i].color[0], lights_in[i].color[1], lights_in[i].color[2]);
        L.intensity = lights_in[i].intensity; L.radius = 0.0f; L.shadow_samples = 1;
        lights.push_back(L);
    }
    camera cam(make_vec3(cpos[0], cpos[1], cpos[2]), make_vec3(look[0], look[1], look[2]), make_vec3(up[0], up[1], up[2]),
               focal_mm, sensor_h_mm, sensor_w_mm, W, H);
    const Vec3 center = cam.get_center();
    const int maxDepth = std::max(1, max_bounces);          // render.cpp:113
    if (row_step < 1) row_step = 1;
    for (int j = row_begin; j < H; j += row_step)
        for (int i = 0; i < W; ++i) {
            const double u = static_cast<double>(i) + 0.5, v = static_cast<double>(j) + 0.5;
            Vec3 target = cam.get_pixel_position(u, v);
            Ray r(center, target - center);
            const size_t pix = (size_t)j * W + i;
            if (rgb) {
                Vec3 c = make_vec3(0, 0, 0) + TraceRay(r, tris, lights, maxDepth, false);
                c = c / static_cast<float>(1);
                rgb[3 * pix] = c.x; rgb[3 * pix + 1] = c.y; rgb[3 * pix + 2] = c.z;
            }
            if (tri_id || tout) {
                int best = -1; double closest = std::numeric_limits<double>::infinity();
                for (size_t k = 0; k < tris.size(); ++k) {
                    HitRecord h = ray_intersection(r, tris[k]);
                    if (h.hit && h.t >= RT_EPS && h.t < closest) { closest = h.t; best = (int)k; }
                }
                if (tri_id) tri_id[pix] = best;
                if (tout) tout[pix] = best >= 0 ? (float)closest : -1.0f;
            }
        }
}

// Soft shadows (Light::radius / shadow_samples, raytracer.h:37-46,121-168): the same pixel loop run `runs` times with the
// reference's own random_float() (std::mt19937 seeded by std::random_device, :12-16 — every run differs); per pixel and
// channel the mean and the mean square of the `runs` colours.  Single-threaded: the generator is a function-local static.
void ref_cpu_render_area_stats(const float* pos, const float* nrm, uint64_t nv, const uint32_t* idx, uint64_t nt, const int32_t* obj_ids,
                               const void* materials52, int num_materials, const float* cpos, const float* look, const float* up,
                               double focal_mm, double sensor_h_mm, double sensor_w_mm, int W, int H, const ref_cpu_light* lights_in,
                               const float* radius, const int* shadow_samples, int num_lights, int max_bounces, int runs,
                               double* mean, double* meansq)
{
    (void)nv;
    const Vec3* P = (const Vec3*)pos; const Vec3* N = (const Vec3*)nrm;
    const Material* mats = (const Material*)materials52;
    std::vector<Triangle> tris;
    tris.reserve(nt);
    for (uint64_t k = 0; k < nt; ++k) {
        Triangle tri;
        const int o = obj_ids ? obj_ids[k] : 0;
        tri.mat = (o >= 0 && o < num_materials) ? mats[o] : Material{};
        tri.v0 = P[idx[3 * k]]; tri.v1 = P[idx[3 * k + 1]]; tri.v2 = P[idx[3 * k + 2]];
        if (N) { tri.n0 = N[idx[3 * k]]; tri.n1 = N[idx[3 * k + 1]]; tri.n2 = N[idx[3 * k + 2]]; }
        else { Vec3 faceN = unit_vector(cross(tri.v1 - tri.v0, tri.v2 - tri.v0)); tri.n0 = tri.n1 = tri.n2 = faceN; }
        tris.push_back(tri);
    }
    std::vector<Light> lights;
    for (int i = 0; i < num_lights; ++i) {
        Light L;
        L.position = make_vec3(lights_in[i].position[0], lights_in[i].position[1], lights_in[i].position[2]);
        L.color = make_vec3(lights_in[i].color[0], lights_in[i].color[1], lights_in[i].color[2]);
        L.intensity = lights_in[i].intensity; L.radius = radius[i]; L.shadow_samples = shadow_samples[i];
        lights.push_back(L);
    }
    camera cam(make_vec3(cpos[0], cpos[1], cpos[2]), make_vec3(look[0], look[1], look[2]), make_vec3(up[0], up[1], up[2]),
               focal_mm, sensor_h_mm, sensor_w_mm, W, H);
    const Vec3 center = cam.get_center();
    const int maxDepth = std::max(1, max_bounces);
    for (size_t k = 0; k < (size_t)3 * W * H; ++k) { mean[k] = 0.0; meansq[k] = 0.0; }
    for (int run = 0; run < runs; ++run)
        for (int j = 0; j < H; ++j)
            for (int i = 0; i < W; ++i) {
                Vec3 target = cam.get_pixel_position(static_cast<double>(i) + 0.5, static_cast<double>(j) + 0.5);
                Ray r(center, target - center);
                const Vec3 c = make_vec3(0, 0, 0) + TraceRay(r, tris, lights, maxDepth, false);
                const size_t pix = 3 * ((size_t)j * W + i);
                mean[pix] += c.x; mean[pix + 1] += c.y; mean[pix + 2] += c.z;
                meansq[pix] += (double)c.x * c.x; meansq[pix + 1] += (double)c.y * c.y; meansq[pix + 2] += (double)c.z * c.z;
            }
    for (size_t k = 0; k < (size_t)3 * W * H; ++k) { mean[k] /= runs; meansq[k] /= runs; }
}

// The 8-bit conversion of render.cpp:157-163: clamp to [0,1], (uchar)(255.99f*c).
void ref_cpu_quantise(const float* rgb, uint64_t n, uint8_t* out) {
    for (uint64_t k = 0; k < n; ++k) {
        Vec3 c = clamp(make_vec3(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]));
        out[3 * k + 0] = static_cast<unsigned char>(255.99f * c.x);
        out[3 * k + 1] = static_cast<unsigned char>(255.99f * c.y);
        out[3 * k + 2] = static_cast<unsigned char>(255.99f * c.z);
    }
}

} // extern "C"
