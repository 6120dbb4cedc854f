// rt_host.cpp — pure host helpers of the C ABI: camera set-up and jitter tables.
// The device only ever sees the four vectors Camera::initialize leaves behind
// ("stays on host, device receives the 4 float3"), so the mixed fp64/fp32
// rounding of that routine is reproduced here once per frame.  Built with -ffp-contract=off.
#include "../../include/rt_api.h"

#include <cmath>
#include <random>

namespace {
struct V { float x, y, z; };
inline V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V add(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V scale(V a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V cross(V u, V v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
// Vec3 / double: each component is divided in fp64 and narrowed (vec3.h operator/(Vec3,double)).
inline V div64(V a, double t) { return {(float)(a.x / t), (float)(a.y / t), (float)(a.z / t)}; }
// Camera::unit_vector with its (0,0,1) fallback (GPUandCPU/include/camera.h:64-69).
inline V unit_or_z(V v) {
    float len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    if ((double)len < 1e-12) return {0.0f, 0.0f, 1.0f};
    return div64(v, (double)len);
}
inline void put(float* dst, V v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
} // namespace

// sensor_width_mm <= 0: viewport width from the pixel aspect ratio (GPUandCPU, HW1); > 0: the sensor's own width (CPUOnly).
static int camera_init(rt_camera* out, const float pos[3], const float look_at[3], const float up[3],
                       double focal_length_mm, double sensor_height_mm, double sensor_width_mm, int width, int height) {
    if (!out || !pos || !look_at || !up || width < 1 || height < 1) return RT_ERR_ARG;
    const V center{pos[0], pos[1], pos[2]}, target{look_at[0], look_at[1], look_at[2]}, up_v{up[0], up[1], up[2]};
    const V forward = unit_or_z(sub(target, center));
    const V right = unit_or_z(cross(forward, up_v));
    const V up_ortho = cross(right, forward);
    const double focal_m = focal_length_mm / 1000.0;
    const double vp_h = sensor_height_mm / 1000.0;
    const double vp_w = sensor_width_mm > 0.0 ? sensor_width_mm / 1000.0 : vp_h * (double(width) / double(height));
    // `double * Vec3` binds to operator*(float, Vec3): the scalar is narrowed before the multiply.
    const V vp_u = scale(right, (float)vp_w);
    const V vp_v = scale(up_ortho, (float)(-vp_h));
    const V du = div64(vp_u, double(width));
    const V dv = div64(vp_v, double(height));
    const V vp_center = add(center, scale(forward, (float)focal_m));
    const V upper_left = sub(sub(vp_center, scale(vp_u, 0.5f)), scale(vp_v, 0.5f));
    const V p00 = add(upper_left, scale(add(du, dv), 0.5f));
    put(out->center, center); put(out->pixel00_loc, p00); put(out->pixel_delta_u, du); put(out->pixel_delta_v, dv);
    return RT_OK;
}

extern "C" int rt_camera_init(rt_camera* out, const float pos[3], const float look_at[3], const float up[3],
                              double focal_length_mm, double sensor_height_mm, int width, int height) {
    // GPUandCPU/include/camera.h:72-94; HW1/include/camera.h:55-92 throws on width/height < 1.
    return camera_init(out, pos, look_at, up, focal_length_mm, sensor_height_mm, 0.0, width, height);
}

extern "C" int rt_camera_init_cpuonly(rt_camera* out, const float pos[3], const float look_at[3], const float up[3],
                                      double focal_length_mm, double sensor_height_mm, double sensor_width_mm, int width, int height) {
    // HW2/HW2/CPUOnly/include/camera.h:64-104 (same arithmetic; viewport width = sensor_width_mm / 1000).
    if (!(sensor_width_mm > 0.0)) return RT_ERR_ARG;
    return camera_init(out, pos, look_at, up, focal_length_mm, sensor_height_mm, sensor_width_mm, width, height);
}

extern "C" int rt_jitter_table(float* out, int spp, uint32_t seed, int centered) {
    // jittered_samples: GPUandCPU/include/antialias.h:12-27 (centered), HW1/include/antialias.h:12-27 (not).
    if (!out || spp < 0) return RT_ERR_ARG;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> uni(0.0f, 1.0f);
    for (int s = 0; s < spp; ++s) {
        float dx = uni(rng), dy = uni(rng);
        if (centered) { dx = dx - 0.5f; dy = dy - 0.5f; }
        out[2 * s] = dx; out[2 * s + 1] = dy;
    }
    return RT_OK;
}
