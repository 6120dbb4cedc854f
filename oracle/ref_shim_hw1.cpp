// ref_shim_hw1.cpp — C-ABI doorway into the UNMODIFIED HW1 reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/rt_oracle.c header).  This file contains no
// renderer arithmetic of its own: it #includes the reference headers where they lie
// under /root/reference/HW1/include (passed with -I by oracle/build.py) and calls the
// reference's camera, Ray, ray_intersection and shade.  The pixel loop of
// HW1/src/render.cpp:72-116 lives inside main() there, so it is re-driven here
// call for call.  Output goes to oracle/_ref/libref_hw1.so (git-ignored).
#define private public   // camera keeps pixel00_loc / pixel_delta_* private
#include "camera.h"
#undef private
#include "ray.h"
#include "raytracer.h"
#include "antialias.h"
#include "MeshOBJ.h"

#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

extern "C" {

// LoadOBJ_ToMeshSOA (HW1/src/MeshOBJ.cpp:143) ------------------------------------------
struct ref_mesh { MeshSOA m; };
void* ref_hw1_load_obj(const char* path) {
    ref_mesh* r = new ref_mesh;
    if (!LoadOBJ_ToMeshSOA(path, r->m)) { delete r; return nullptr; }
    return r;
}
void ref_hw1_mesh_counts(void* h, uint64_t* nv, uint64_t* nn, uint64_t* nt) {
    ref_mesh* r = (ref_mesh*)h;
    *nv = r->m.positions.size(); *nn = r->m.normals.size(); *nt = r->m.indices.size() / 3;
}
void ref_hw1_mesh_copy(void* h, float* pos, float* nrm, uint32_t* idx) {
    ref_mesh* r = (ref_mesh*)h;
    if (pos) std::memcpy(pos, r->m.positions.data(), r->m.positions.size() * sizeof(Vec3));
    if (nrm) std::memcpy(nrm, r->m.normals.data(), r->m.normals.size() * sizeof(Vec3));
    if (idx) std::memcpy(idx, r->m.indices.data(), r->m.indices.size() * sizeof(uint32_t));
}
void ref_hw1_mesh_free(void* h) { delete (ref_mesh*)h; }

// camera (HW1/include/camera.h) -> center, pixel00_loc, delta_u, delta_v (12 floats)
int ref_hw1_camera(const float* pos, const float* look, const float* up, double focal_mm,
                   double sensor_mm, int W, int H, float* out12) {
    try {
        camera cam(make_vec3(pos[0], pos[1], pos[2]), make_vec3(look[0], look[1], look[2]),
                   make_vec3(up[0], up[1], up[2]), focal_mm, sensor_mm, W, H);
        Vec3 v[4] = {cam.center, cam.pixel00_loc, cam.pixel_delta_u, cam.pixel_delta_v};
        std::memcpy(out12, v, sizeof v);
        return 0;
    } catch (...) { return -1; }
}

void ref_hw1_jitter(int spp, unsigned seed, float* out) {
    auto o = jittered_samples(spp, seed);
    for (int i = 0; i < spp; ++i) { out[2 * i] = o[i].first; out[2 * i + 1] = o[i].second; }
}

// Single-triangle probe for the reference's own unit vectors.
int ref_hw1_ray_triangle(const float* orig, const float* dir, const float* v0, const float* v1,
                         const float* v2, float* t_out) {
    Ray r(make_vec3(orig[0], orig[1], orig[2]), make_vec3(dir[0], dir[1], dir[2]));
    Triangle tri{};
    tri.v0 = make_vec3(v0[0], v0[1], v0[2]); tri.v1 = make_vec3(v1[0], v1[1], v1[2]); tri.v2 = make_vec3(v2[0], v2[1], v2[2]);
    tri.n0 = tri.n1 = tri.n2 = make_vec3(0, 0, 0);
    HitRecord rec = ray_intersection(r, tri);
    if (t_out) *t_out = rec.hit ? (float)rec.t : -1.0f;
    return rec.hit ? 1 : 0;
}

// The loop of HW1/src/render.cpp:72-124 driven over rows row_begin, +row_step, ...
// rgb/rgb8/tri_id/t may be NULL.  Returns the number of ray/triangle tests done.
uint64_t ref_hw1_render(const float* pos, const float* nrm, const uint32_t* idx, uint64_t ntri,
                        const float* cpos, const float* look, const float* up, double focal_mm,
                        double sensor_mm, int W, int H, const float* light_pos, const float* light_col,
                        int spp, unsigned seed, int row_begin, int row_step, int nthreads,
                        float* rgb, uint8_t* rgb8, int32_t* tri_id, float* tout)
{
    camera cam(make_vec3(cpos[0], cpos[1], cpos[2]), make_vec3(look[0], look[1], look[2]),
               make_vec3(up[0], up[1], up[2]), focal_mm, sensor_mm, W, H);
    Light light;
    light.position = make_vec3(light_pos[0], light_pos[1], light_pos[2]);
    light.color = make_vec3(light_col[0], light_col[1], light_col[2]);
    const Vec3* P = (const Vec3*)pos;
    const Vec3* N = (const Vec3*)nrm;
    auto offsets = jittered_samples(spp, seed);
    auto center = cam.get_center();
    const size_t indexCount = ntri * 3;
    if (nthreads < 1) nthreads = 1;
    if (row_step < 1) row_step = 1;
    auto body = [&](int tid) {
        int kk = 0;
        for (int j = row_begin; j < H; j += row_step, ++kk) {
            if (kk % nthreads != tid) continue;
            for (int i = 0; i < W; i++) {
                Vec3 accum_color = make_vec3(0.0f, 0.0f, 0.0f);
                int first_id = -1; float first_t = -1.0f; bool first = true;
                for (const auto& o : offsets) {
                    float px = float(i) + o.first;
                    float py = float(j) + o.second;
                    Ray r = Ray(center, cam.get_pixel_position(px, py) - center);
                    HitRecord prev;
                    prev.hit = false;
                    prev.t = std::numeric_limits<float>::max();
                    auto color = shade(r, prev, light);
                    int best = -1;
                    for (size_t k = 0; k < indexCount; k += 3) {
                        Triangle tri;
                        tri.v0 = P[idx[k]]; tri.v1 = P[idx[k + 1]]; tri.v2 = P[idx[k + 2]];
                        tri.n0 = N[idx[k]]; tri.n1 = N[idx[k + 1]]; tri.n2 = N[idx[k + 2]];
                        HitRecord rec = ray_intersection(r, tri);
                        if (rec.hit && rec.t < prev.t) {
                            color = shade(r, rec, light);
                            prev = rec;
                            best = (int)(k / 3);
                        }
                    }
                    if (first) { first_id = best; first_t = best >= 0 ? (float)prev.t : -1.0f; first = false; }
                    accum_color = accum_color + color;
                }
                Vec3 final_color = accum_color / float(offsets.size());
                size_t pix = (size_t)j * W + i;
                if (rgb) { rgb[3 * pix] = final_color.x; rgb[3 * pix + 1] = final_color.y; rgb[3 * pix + 2] = final_color.z; }
                if (rgb8) {
                    rgb8[3 * pix + 0] = (unsigned char)(255.99f * final_color.x);
                    rgb8[3 * pix + 1] = (unsigned char)(255.99f * final_color.y);
                    rgb8[3 * pix + 2] = (unsigned char)(255.99f * final_color.z);
                }
                if (tri_id) tri_id[pix] = first_id;
                if (tout) tout[pix] = first_t;
            }
        }
    };
    if (nthreads == 1) body(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(body, t);
        for (auto& t : th) t.join();
    }
    uint64_t rows = 0;
    for (int j = row_begin; j < H; j += row_step) ++rows;
    return rows * (uint64_t)W * (uint64_t)spp * ntri;
}

} // extern "C"
