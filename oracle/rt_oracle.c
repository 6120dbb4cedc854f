/* rt_oracle.c — CPU restatement of the reference ray-casting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (raytracinginonesemester_b200/,
 * include/) may link, import or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py (and tests/test_soft_shadows.py for the
 * disk lights) check this file against
 *   - the reference's own code compiled in place (oracle/ref_shim_hw1.cpp, ref_shim_hw2.cpp,
 *     ref_shim_cpuonly.cpp, ref_shim_ppm.cpp -> oracle/_ref/*.so, built by oracle/build.py),
 *   - the reference's committed golden images HW1/frog_output.png and
 *     CPUOnly/output/sphere_point_output.png, and the ray/triangle unit vectors of
 *     HW1/test_ray_tri_inter_STANDALONE,
 *   - golden fixtures under tests/golden/ generated from those shims by tools/make_golden*.py
 *     (make_golden.py, make_golden_bounce.py, make_golden_cpuonly.py, make_golden_area.py,
 *     make_golden_e2e.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -pthread (no -march=native, no
 * -ffast-math): the x86-64 reference build has no FMA contraction and the hit
 * ids depend on that.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  G = HW2/HW2/GPUandCPU/include, H1 = HW1/include.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rt_api.h"

typedef struct { float x, y, z; } v3;

/* ---------------------------------------------------------------- vec3 ---- */
/* G/vec3.h:38-58 and H1/vec3.h:38-56: operator order is part of the contract. */
static inline v3 mk(float x, float y, float z) { v3 v = {x, y, z}; return v; }
static inline v3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
static inline v3 add(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 neg(v3 a) { return mk(-a.x, -a.y, -a.z); }
static inline v3 mulv(v3 a, v3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 muls(v3 v, float t) { return mk(v.x * t, v.y * t, v.z * t); }
static inline v3 divv(v3 a, v3 b) { return mk(a.x / b.x, a.y / b.y, a.z / b.z); }
/* operator/(Vec3, double): fp64 divide then narrow (vec3.h:45 / :44) */
static inline v3 divd(v3 a, double t) { return mk((float)(a.x / t), (float)(a.y / t), (float)(a.z / t)); }
static inline float dot(v3 u, v3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }
static inline v3 cross(v3 u, v3 v) {
    return mk(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
/* global unit_vector, vec3.h:53-56 (three float divides) */
static inline v3 unit_vector(v3 v) {
    float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    return mk(v.x / len, v.y / len, v.z / len);
}
/* G/vec3.h:50-52 */
static inline float length_squared(v3 v) { return dot(v, v); }
static inline v3 normalize(v3 v) { return divd(v, (double)sqrtf(dot(v, v))); }

/* -------------------------------------------------------------- camera ---- */
/* Camera::unit_vector (G/camera.h:64-69, H1/camera.h:48-53) */
static v3 cam_unit_vector(v3 v) {
    float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if ((double)len < 1e-12) return mk(0.0f, 0.0f, 1.0f);
    return divd(v, (double)len);
}

/* Camera::initialize, G/camera.h:72-94 (== H1/camera.h:55-92 for valid sizes).
 * `double * Vec3` resolves to operator*(float, Vec3): the double is narrowed first. */
static int camera_init_common(rt_camera* out, const float pos[3], const float look_at_[3], const float up_[3],
                              double focal_length_mm, double sensor_height_mm, double sensor_width_mm, int width, int height)
{
    if (width < 1 || height < 1) return RT_ERR_ARG; /* H1/camera.h:57-62 throws; G clamps */
    v3 center = ld3(pos), look_at = ld3(look_at_), up = ld3(up_);
    v3 forward = cam_unit_vector(sub(look_at, center));
    v3 right = cam_unit_vector(cross(forward, up));
    v3 up_corrected = cross(right, forward);

    double focal_length_m = focal_length_mm / 1000.0;
    double sensor_height_m = sensor_height_mm / 1000.0;
    double viewport_height = sensor_height_m;
    /* G/camera.h:82: aspect-derived width; CPUOnly/include/camera.h:85-86: the sensor's own width */
    double viewport_width = sensor_width_mm > 0.0 ? sensor_width_mm / 1000.0 : viewport_height * ((double)width / (double)height);

    v3 viewport_u = muls(right, (float)viewport_width);
    v3 viewport_v = muls(up_corrected, (float)(-viewport_height));
    v3 du = divd(viewport_u, (double)width);
    v3 dv = divd(viewport_v, (double)height);

    v3 viewport_center = add(center, muls(forward, (float)focal_length_m));
    v3 upper_left = sub(sub(viewport_center, muls(viewport_u, 0.5f)), muls(viewport_v, 0.5f));
    v3 p00 = add(upper_left, muls(add(du, dv), 0.5f));

    out->center[0] = center.x; out->center[1] = center.y; out->center[2] = center.z;
    out->pixel00_loc[0] = p00.x; out->pixel00_loc[1] = p00.y; out->pixel00_loc[2] = p00.z;
    out->pixel_delta_u[0] = du.x; out->pixel_delta_u[1] = du.y; out->pixel_delta_u[2] = du.z;
    out->pixel_delta_v[0] = dv.x; out->pixel_delta_v[1] = dv.y; out->pixel_delta_v[2] = dv.z;
    return RT_OK;
}

int orc_camera_init(rt_camera* out, const float pos[3], const float look_at_[3], const float up_[3],
                    double focal_length_mm, double sensor_height_mm, int width, int height)
{
    return camera_init_common(out, pos, look_at_, up_, focal_length_mm, sensor_height_mm, 0.0, width, height);
}
/* camera::initialize, HW2/HW2/CPUOnly/include/camera.h:64-104 */
int orc_camera_init_cpuonly(rt_camera* out, const float pos[3], const float look_at_[3], const float up_[3],
                            double focal_length_mm, double sensor_height_mm, double sensor_width_mm, int width, int height)
{
    return camera_init_common(out, pos, look_at_, up_, focal_length_mm, sensor_height_mm, sensor_width_mm, width, height);
}

/* ------------------------------------------------------------- mt19937 ---- */
/* std::mt19937 + libstdc++ generate_canonical<float,24> as used by
 * jittered_samples (G/antialias.h:12-27, H1/antialias.h:12-27). */
typedef struct { uint32_t mt[624]; int idx; } mt19937;
static void mt_seed(mt19937* s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}
static uint32_t mt_next(mt19937* s) {
    if (s->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
            s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
static float mt_uniform01(mt19937* s) {
    float r = (float)mt_next(s) / 4294967296.0f;
    if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
    return r;
}
/* centered=1: GPUandCPU variant ([-0.5,0.5)); centered=0: HW1 variant ([0,1)). */
int orc_jitter_table(float* out, int spp, uint32_t seed, int centered) {
    if (!out || spp < 0) return RT_ERR_ARG;
    mt19937 s; mt_seed(&s, seed);
    for (int i = 0; i < spp; ++i) {
        float dx = mt_uniform01(&s), dy = mt_uniform01(&s);
        if (centered) { dx = dx - 0.5f; dy = dy - 0.5f; }
        out[2 * i] = dx; out[2 * i + 1] = dy;
    }
    return RT_OK;
}

/* ------------------------------------------------------------ triangles ---- */
typedef struct { v3 v0, v1, v2, n0, n1, n2; } tri_t; /* Triangle, G/MeshOBJ.h:42-67 */

static tri_t* gather_triangles(const rt_scene* sc) {
    /* buildTriangles loop, GPUandCPU/src/main.cu:387-403; HW1/src/render.cpp:90-99 */
    size_t P = (size_t)sc->num_triangles;
    tri_t* t = (tri_t*)malloc((P ? P : 1) * sizeof(tri_t));
    for (size_t i = 0; i < P; ++i) {
        uint32_t a = sc->indices[3 * i], b = sc->indices[3 * i + 1], c = sc->indices[3 * i + 2];
        t[i].v0 = ld3(sc->positions + 3 * (size_t)a);
        t[i].v1 = ld3(sc->positions + 3 * (size_t)b);
        t[i].v2 = ld3(sc->positions + 3 * (size_t)c);
        if (sc->normals) {
            t[i].n0 = ld3(sc->normals + 3 * (size_t)a);
            t[i].n1 = ld3(sc->normals + 3 * (size_t)b);
            t[i].n2 = ld3(sc->normals + 3 * (size_t)c);
        } else {
            t[i].n0 = t[i].n1 = t[i].n2 = mk(0, 0, 0);
        }
    }
    return t;
}

typedef struct { v3 orig, dir; } ray_t;

typedef struct {
    int hit; int tri; float t; v3 p; v3 normal; int front_face;
} hit_t;

/* ray_intersection, H1/ray.h:67-117.  Compares promote to double (harmless). */
static int mt_hw1(const ray_t* r, const tri_t* tri, hit_t* rec) {
    const float eps = FLT_EPSILON;
    v3 e1 = sub(tri->v1, tri->v0);
    v3 e2 = sub(tri->v2, tri->v0);
    v3 pvec = cross(r->dir, e2);
    float det = dot(pvec, e1);
    if (fabsf(det) < eps) return 0;
    float invDet = (float)(1.0 / (double)det);
    v3 tvec = sub(r->orig, tri->v0);
    float u = dot(tvec, pvec) * invDet;
    if (u < 0.0f || u > 1.0f) return 0;
    v3 qvec = cross(tvec, e1);
    float v = dot(r->dir, qvec) * invDet;
    if (v < 0.0f || u + v > 1.0f) return 0;
    float t = dot(e2, qvec) * invDet;
    if (t < 0.0f) return 0;
    rec->hit = 1;
    rec->t = t;
    rec->p = add(r->orig, muls(r->dir, t)); /* Ray::at, H1/ray.h:31-33: orig + t*dir */
    /* (1 - u - v)*n0 + u*n1 + v*n2, un-normalised (ray.h:109) */
    rec->normal = add(add(muls(tri->n0, 1 - u - v), muls(tri->n1, u)), muls(tri->n2, v));
    rec->front_face = 0;
    return 1;
}

/* intersectTriangle, G/query.h:72-132 */
static int mt_hw2(const ray_t* r, const tri_t* tri, float tmin, float tmax, hit_t* rec) {
    v3 e1 = sub(tri->v1, tri->v0);
    v3 e2 = sub(tri->v2, tri->v0);
    v3 pvec = cross(r->dir, e2);
    float det = dot(e1, pvec);
    if (fabsf(det) < 1e-8f) return 0;
    float invDet = 1.0f / det;
    v3 tvec = sub(r->orig, tri->v0);
    float u = dot(tvec, pvec) * invDet;
    if (u < 0.0f || u > 1.0f) return 0;
    v3 qvec = cross(tvec, e1);
    float v = dot(r->dir, qvec) * invDet;
    if (v < 0.0f || (u + v) > 1.0f) return 0;
    float t = dot(e2, qvec) * invDet;
    if (t < tmin || t > tmax) return 0;

    rec->hit = 1;
    rec->t = t;
    rec->p = add(r->orig, muls(r->dir, t));
    v3 geomN = normalize(cross(e1, e2));
    rec->front_face = dot(r->dir, geomN) < 0.0f;
    if (!rec->front_face) geomN = neg(geomN);
    v3 shadingN = add(add(muls(tri->n0, 1.0f - u - v), muls(tri->n1, u)), muls(tri->n2, v));
    if (length_squared(shadingN) < 1e-12f) {
        shadingN = geomN;
    } else {
        shadingN = normalize(shadingN);
        if (dot(shadingN, geomN) < 0.0f) shadingN = neg(shadingN);
    }
    rec->normal = shadingN;
    return 1;
}

/* unit_vector of HW2/HW2/CPUOnly/include/vec3.h:53-57 (zero-length guard) */
static v3 unit_vector_c(v3 v) {
    float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (len < 1e-12f) return mk(0, 0, 0);
    return mk(v.x / len, v.y / len, v.z / len);
}
/* ray_intersection of HW2/HW2/CPUOnly/include/ray.h:48-97.  have_normals == 0: the triangle carries its face normal
 * in all three slots (CPUOnly/src/render.cpp:88-96). */
static int mt_cpuonly(const ray_t* r, const tri_t* tri, int have_normals, hit_t* rec) {
    const float eps = FLT_EPSILON;
    v3 e1 = sub(tri->v1, tri->v0);
    v3 e2 = sub(tri->v2, tri->v0);
    v3 pvec = cross(r->dir, e2);
    float det = dot(pvec, e1);
    if (fabsf(det) < eps) return 0;
    float invDet = 1.0f / det;
    v3 tvec = sub(r->orig, tri->v0);
    float u = dot(tvec, pvec) * invDet;
    if (u < 0.0f || u > 1.0f) return 0;
    v3 qvec = cross(tvec, e1);
    float v = dot(r->dir, qvec) * invDet;
    if (v < 0.0f || u + v > 1.0f) return 0;
    float t = dot(e2, qvec) * invDet;
    if (t < 0.0f) return 0;
    rec->hit = 1;
    rec->t = t;
    rec->p = add(r->orig, muls(r->dir, t));               /* Ray::at: orig + float(t)*dir */
    v3 faceN = unit_vector_c(cross(e1, e2));
    rec->front_face = dot(r->dir, faceN) < 0;             /* set_face_normal, ray.h:41-44 */
    v3 n0 = tri->n0, n1 = tri->n1, n2 = tri->n2;
    if (!have_normals) n0 = n1 = n2 = faceN;
    v3 shadeN = add(add(muls(n0, 1.0f - u - v), muls(n1, u)), muls(n2, v));
    shadeN = unit_vector_c(shadeN);
    if (!rec->front_face) shadeN = neg(shadeN);
    rec->normal = shadeN;
    return 1;
}

/* ----------------------------------------------------------------- LBVH ---- */
typedef struct { uint32_t parent, left, right, object; } node_t; /* BVHNode, G/bvh.h:7-13 */
typedef struct { v3 mn, mx; } aabb_t;                              /* AABB, G/bvh.h:28-53    */

typedef struct orc_bvh {
    uint32_t P;
    node_t* nodes;   /* 2P-1: internal 0..P-2, leaves P-1..2P-2 */
    aabb_t* aabbs;
    aabb_t scene;
} orc_bvh;

static aabb_t aabb_merge(aabb_t a, aabb_t b) { /* AABB::merge, G/bvh.h:36-51 */
    aabb_t r;
    r.mn = mk(fminf(a.mn.x, b.mn.x), fminf(a.mn.y, b.mn.y), fminf(a.mn.z, b.mn.z));
    r.mx = mk(fmaxf(a.mx.x, b.mx.x), fmaxf(a.mx.y, b.mx.y), fmaxf(a.mx.z, b.mx.z));
    return r;
}

static uint32_t bit_expansion(uint32_t v) { /* G/bvh.h:131-138 */
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static uint32_t morton_code(v3 p) { /* ComputeMortonCode, G/bvh.h:142-151 */
    const float res = 1024.0f;
    p.x = fminf(fmaxf(p.x * res, 0.0f), res - 1.0f);
    p.y = fminf(fmaxf(p.y * res, 0.0f), res - 1.0f);
    p.z = fminf(fmaxf(p.z * res, 0.0f), res - 1.0f);
    uint32_t xx = bit_expansion((uint32_t)p.x), yy = bit_expansion((uint32_t)p.y), zz = bit_expansion((uint32_t)p.z);
    return xx * 4 + yy * 2 + zz;
}
static int cub64(uint64_t a, uint64_t b) { /* common_upper_bits_cpu, G/bvh.h:297-301 */
    uint64_t d = a ^ b;
    return d == 0 ? 64 : __builtin_clzll(d);
}
/* determine_range_cpu, G/bvh.h:304-358 */
static void determine_range(const uint64_t* code, uint32_t n, uint32_t idx, uint32_t* lo, uint32_t* hi) {
    if (idx == 0) { *lo = 0; *hi = n - 1; return; }
    uint64_t self = code[idx];
    int Ld = cub64(self, code[idx - 1]);
    int Rd = cub64(self, code[idx + 1]);
    int d = (Rd > Ld) ? 1 : -1;
    int dmin = Ld < Rd ? Ld : Rd;
    int lmax = 2, delta = -1;
    long it = (long)idx + (long)d * lmax;
    if (0 <= it && it < (long)n) delta = cub64(self, code[it]);
    while (delta > dmin) {
        lmax <<= 1;
        it = (long)idx + (long)d * lmax;
        delta = -1;
        if (0 <= it && it < (long)n) delta = cub64(self, code[it]);
    }
    int l = 0;
    for (int t = lmax >> 1; t > 0; t >>= 1) {
        it = (long)idx + (long)(l + t) * d;
        delta = -1;
        if (0 <= it && it < (long)n) delta = cub64(self, code[it]);
        if (delta > dmin) l += t;
    }
    uint32_t j = (uint32_t)((long)idx + (long)l * d);
    if (d < 0) { *lo = j; *hi = idx; } else { *lo = idx; *hi = j; }
}
/* find_split_cpu, G/bvh.h:361-392 */
static uint32_t find_split(const uint64_t* code, uint32_t first, uint32_t last) {
    uint64_t fc = code[first], lc = code[last];
    if (fc == lc) return (first + last) >> 1;
    int dn = cub64(fc, lc);
    int split = (int)first, stride = (int)(last - first);
    do {
        stride = (stride + 1) >> 1;
        int mid = split + stride;
        if (mid < (int)last) {
            int delta = cub64(fc, code[mid]);
            if (delta > dn) split = mid;
        }
    } while (stride > 1);
    return (uint32_t)split;
}
static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* calculateAABBs (G/bvh.cu:60-90) + scene bounds (GPUandCPU/src/main.cu:296-302)
 * + buildBVH CPU (G/bvh.cu:209-317).  refit_cpu (G/bvh.h:394-404) is done as an
 * explicit post-order so 10M-triangle inputs cannot blow the C stack; the merge
 * per node is the same fminf/fmaxf pair. */
orc_bvh* orc_bvh_build(const rt_scene* sc) {
    uint32_t P = (uint32_t)sc->num_triangles;
    if (P == 0) return NULL;
    orc_bvh* b = (orc_bvh*)calloc(1, sizeof(orc_bvh));
    b->P = P;
    size_t total = 2 * (size_t)P - 1;
    b->nodes = (node_t*)malloc(total * sizeof(node_t));
    b->aabbs = (aabb_t*)malloc(total * sizeof(aabb_t));
    memset(b->nodes, 0xFF, total * sizeof(node_t));
    aabb_t* leaf = (aabb_t*)malloc((size_t)P * sizeof(aabb_t));
    aabb_t scene;
    scene.mn = mk(INFINITY, INFINITY, INFINITY);
    scene.mx = mk(-INFINITY, -INFINITY, -INFINITY);
    for (uint32_t i = 0; i < P; ++i) {
        v3 a = ld3(sc->positions + 3 * (size_t)sc->indices[3 * (size_t)i]);
        v3 c = ld3(sc->positions + 3 * (size_t)sc->indices[3 * (size_t)i + 1]);
        v3 d = ld3(sc->positions + 3 * (size_t)sc->indices[3 * (size_t)i + 2]);
        aabb_t box; /* aabb_of_triangle, G/bvh.h:55-71 with eps = 0 */
        box.mn = mk(fminf(a.x, fminf(c.x, d.x)), fminf(a.y, fminf(c.y, d.y)), fminf(a.z, fminf(c.z, d.z)));
        box.mx = mk(fmaxf(a.x, fmaxf(c.x, d.x)), fmaxf(a.y, fmaxf(c.y, d.y)), fmaxf(a.z, fmaxf(c.z, d.z)));
        box.mn.x -= 0.0f; box.mn.y -= 0.0f; box.mn.z -= 0.0f;
        box.mx.x += 0.0f; box.mx.y += 0.0f; box.mx.z += 0.0f;
        leaf[i] = box;
        scene = aabb_merge(scene, box);
    }
    b->scene = scene;
    uint64_t* keys = (uint64_t*)malloc((size_t)P * sizeof(uint64_t));
    for (uint32_t i = 0; i < P; ++i) {
        v3 centroid = muls(add(leaf[i].mn, leaf[i].mx), 0.5f);
        v3 nrm = divv(sub(centroid, scene.mn), sub(scene.mx, scene.mn));
        keys[i] = ((uint64_t)morton_code(nrm) << 32) | (uint64_t)i;
    }
    qsort(keys, P, sizeof(uint64_t), cmp_u64); /* keys are unique: any sort gives std::sort's order */
    for (uint32_t i = 0; i < P; ++i) {
        uint32_t tri = (uint32_t)(keys[i] & 0xFFFFFFFFu);
        b->aabbs[(size_t)(P - 1) + i] = leaf[tri];
        b->nodes[(size_t)(P - 1) + i].object = tri;
    }
    for (uint32_t idx = 0; idx + 1 < P; ++idx) {
        uint32_t lo, hi;
        b->nodes[idx].object = 0xFFFFFFFFu;
        determine_range(keys, P, idx, &lo, &hi);
        uint32_t gamma = find_split(keys, lo, hi);
        uint32_t L = gamma, R = gamma + 1;
        if (lo == gamma) L += P - 1;
        if (hi == gamma + 1) R += P - 1;
        b->nodes[idx].left = L;
        b->nodes[idx].right = R;
        b->nodes[L].parent = idx;
        b->nodes[R].parent = idx;
    }
    if (P > 1) {
        /* post-order refit without recursion */
        uint32_t* stack = (uint32_t*)malloc(2 * (size_t)P * sizeof(uint32_t));
        uint8_t* seen = (uint8_t*)calloc(P, 1);
        size_t sp = 0;
        stack[sp++] = 0;
        while (sp) {
            uint32_t n = stack[sp - 1];
            if (n >= P - 1) { --sp; continue; }
            if (!seen[n]) {
                seen[n] = 1;
                stack[sp++] = b->nodes[n].left;
                stack[sp++] = b->nodes[n].right;
            } else {
                b->aabbs[n] = aabb_merge(b->aabbs[b->nodes[n].left], b->aabbs[b->nodes[n].right]);
                --sp;
            }
        }
        free(stack); free(seen);
    }
    free(keys); free(leaf);
    return b;
}
void orc_bvh_free(orc_bvh* b) { if (b) { free(b->nodes); free(b->aabbs); free(b); } }
uint32_t orc_bvh_num_triangles(const orc_bvh* b) { return b ? b->P : 0; }
/* nodes: 4 x u32 each; aabbs: 6 floats each (2P-1 of both) */
void orc_bvh_export(const orc_bvh* b, uint32_t* nodes, float* aabbs) {
    size_t total = 2 * (size_t)b->P - 1;
    if (nodes) memcpy(nodes, b->nodes, total * sizeof(node_t));
    if (aabbs) memcpy(aabbs, b->aabbs, total * sizeof(aabb_t));
}

/* intersectAABB, G/bvh.h:81-129 (fp64 slabs, axis-parallel branch) */
static int intersect_aabb(const ray_t* r, const aabb_t* bx, double tmin, double tmax) {
    const float eps = 1e-8f;
    double t0 = tmin, t1 = tmax;
    const float o[3] = {r->orig.x, r->orig.y, r->orig.z};
    const float d[3] = {r->dir.x, r->dir.y, r->dir.z};
    const float mn[3] = {bx->mn.x, bx->mn.y, bx->mn.z};
    const float mx[3] = {bx->mx.x, bx->mx.y, bx->mx.z};
    for (int a = 0; a < 3; ++a) {
        if (fabsf(d[a]) < eps) {
            if (o[a] < mn[a] || o[a] > mx[a]) return 0;
        } else {
            const double inv = 1.0 / (double)d[a];
            double tn = ((double)mn[a] - (double)o[a]) * inv;
            double tf = ((double)mx[a] - (double)o[a]) * inv;
            if (tn > tf) { double tmp = tn; tn = tf; tf = tmp; }
            if (tn > t0) t0 = tn;
            if (tf < t1) t1 = tf;
            if (t0 > t1) return 0;
        }
    }
    return 1;
}

typedef struct orc_counters {
    uint64_t rays_primary, rays_shadow, node_pops, box_tests, tri_tests, max_stack;
} orc_counters;

/* SearchBVH, G/query.h:224-311.  Last accepted hit wins on equal t (t > tmax rejects). */
static void search_bvh(const orc_bvh* b, const tri_t* tris, const ray_t* ray, hit_t* out, orc_counters* c) {
    const float tmin = 1e-4f;
    float bestT = FLT_MAX;
    hit_t best; memset(&best, 0, sizeof best); best.tri = -1; best.t = -1.0f;
    uint32_t stack[512];
    int sp = 0, overflow = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        uint32_t ni = stack[--sp];
        if (c) { c->node_pops++; c->box_tests++; }
        if (!intersect_aabb(ray, &b->aabbs[ni], tmin, bestT)) continue;
        node_t node = b->nodes[ni];
        if (node.object != 0xFFFFFFFFu) {
            if (node.object < b->P) {
                hit_t rec;
                if (c) c->tri_tests++;
                if (mt_hw2(ray, &tris[node.object], tmin, bestT, &rec)) {
                    rec.tri = (int)node.object;
                    bestT = rec.t;
                    best = rec;
                }
            }
            continue;
        }
        if (node.left != 0xFFFFFFFFu) {
            if (c) c->box_tests++;
            if (intersect_aabb(ray, &b->aabbs[node.left], tmin, bestT)) {
                if (sp < 512) stack[sp++] = node.left; else overflow = 1;
            }
        }
        if (node.right != 0xFFFFFFFFu) {
            if (c) c->box_tests++;
            if (intersect_aabb(ray, &b->aabbs[node.right], tmin, bestT)) {
                if (sp < 512) stack[sp++] = node.right; else overflow = 1;
            }
        }
        if (c && (uint64_t)sp > c->max_stack) c->max_stack = (uint64_t)sp;
    }
    if (overflow) {
        for (uint32_t i = 0; i < b->P; ++i) {
            hit_t rec;
            if (mt_hw2(ray, &tris[i], tmin, bestT, &rec)) { rec.tri = (int)i; bestT = rec.t; best = rec; }
        }
    }
    *out = best;
}

/* Canonical closest hit: min t, then min triangle id, over intersectTriangle on
 * every triangle (SURVEY §7 H2, §8c last row).  Independent of any BVH topology. */
static void brute_hw2(const tri_t* tris, uint32_t P, const ray_t* ray, hit_t* out, orc_counters* c) {
    const float tmin = 1e-4f;
    float bestT = FLT_MAX;
    hit_t best; memset(&best, 0, sizeof best); best.tri = -1; best.t = -1.0f;
    for (uint32_t i = 0; i < P; ++i) {
        hit_t rec;
        if (c) c->tri_tests++;
        if (mt_hw2(ray, &tris[i], tmin, FLT_MAX, &rec) && (!best.hit || rec.t < bestT)) {
            rec.tri = (int)i; bestT = rec.t; best = rec;
        }
    }
    *out = best;
}

/* ---------------------------------------------------------------- shading ---- */
/* shade(), H1/raytracer.h:21-48 with the constant METAL material of H1/ray.h:111-114 */
static v3 shade_hw1(const ray_t* r, const hit_t* rec, const rt_light* light) {
    if (!rec->hit) {
        v3 ud = unit_vector(r->dir);
        float t = 0.5f * (ud.z + 1.0f);
        return add(muls(mk(1.0f, 1.0f, 1.0f), 1.0f - t), muls(mk(0.5f, 0.7f, 1.0f), t));
    }
    const v3 albedo = mk(0.8f, 0.2f, 0.2f);
    const float shininess = 64.0f;
    v3 lpos = ld3(light->position), lcol = ld3(light->color);
    v3 ambient = muls(albedo, 0.1f);
    v3 lightDir = unit_vector(sub(lpos, rec->p));
    float diff = fmaxf(dot(rec->normal, lightDir), 0.0f);
    v3 diffuse = muls(mulv(albedo, lcol), diff);
    v3 viewDir = unit_vector(sub(r->orig, rec->p));
    v3 halfDir = unit_vector(add(lightDir, viewDir));
    float spec = powf(fmaxf(dot(rec->normal, halfDir), 0.0f), shininess);
    v3 specular = muls(lcol, spec);
    v3 c = add(add(ambient, diffuse), specular);
    if (c.x > 1.0) c.x = 1.0f; /* clamp, H1/raytracer.h:13-19 (upper only) */
    if (c.y > 1.0) c.y = 1.0f;
    if (c.z > 1.0) c.z = 1.0f;
    return c;
}

static const rt_material kDefaultMaterial = { /* Material(), G/material.h:6-20 */
    {0.8f, 0.8f, 0.8f}, 1.0f, {0.04f, 0.04f, 0.04f}, 0.0f, 32.0f, 0.0f, {0.0f, 0.0f, 0.0f}};

/* EvaluateBRDF, G/brdf.h:12-39 */
static v3 evaluate_brdf(const rt_material* m, v3 N, v3 V, v3 L) {
    float NdotL = fmaxf(dot(N, L), 0.0f);
    float NdotV = fmaxf(dot(N, V), 0.0f);
    if (NdotL <= 0.f || NdotV <= 0.f) return mk(0, 0, 0);
    const float invPi = 0.31830988618f;
    v3 fd = muls(ld3(m->albedo), m->kd * invPi);
    v3 H = unit_vector(add(L, V));
    float NdotH = fmaxf(dot(N, H), 0.0f);
    const float inv2Pi = 0.15915494309f;
    float specNorm = (m->shininess + 2.0f) * inv2Pi;
    float specLobe = specNorm * powf(NdotH, m->shininess);
    v3 fs = muls(muls(ld3(m->specular_color), m->ks), specLobe);
    return add(fd, fs);
}

typedef struct {
    const rt_scene* sc; const rt_frame* fr; const tri_t* tris; const orc_bvh* bvh;
    uint32_t P;
} world_t;

static void closest_hw2(const world_t* w, const ray_t* ray, hit_t* out, orc_counters* c) {
    if (w->fr->accel == RT_ACCEL_BVH && w->bvh) search_bvh(w->bvh, w->tris, ray, out, c);
    else brute_hw2(w->tris, w->P, ray, out, c);
}

/* ShadeDirect + IsInShadow, G/shader.h:44-110; RT_EPS = 1e-3 (shader.h:22) */
static v3 shade_direct_hw2(const world_t* w, const ray_t* r, const hit_t* rec, const rt_material* mat, orc_counters* c) {
    const float RT_EPS = 1e-3f;
    v3 N = unit_vector(rec->normal);
    v3 V = unit_vector(sub(r->orig, rec->p));
    v3 Lo = mk(0, 0, 0);
    Lo = add(Lo, muls(ld3(mat->albedo), 0.05f));
    Lo = add(Lo, ld3(mat->emission));
    for (int i = 0; i < w->fr->num_lights; ++i) {
        const rt_light* light = &w->fr->lights[i];
        v3 lpos = ld3(light->position);
        v3 L = unit_vector(sub(lpos, rec->p));
        float NdotL = fmaxf(dot(N, L), 0.0f);
        if (NdotL <= 0.0f) continue;
        if (w->fr->shadows) {
            v3 toL = sub(lpos, rec->p);
            float distToL = sqrtf(dot(toL, toL));
            if (!(distToL <= 0.0f)) {
                ray_t sray;
                sray.dir = divd(toL, (double)distToL);
                sray.orig = add(rec->p, muls(N, RT_EPS));
                hit_t sh;
                if (c) c->rays_shadow++;
                closest_hw2(w, &sray, &sh, c);
                if (sh.hit && (double)sh.t < (double)distToL) continue;
            }
        }
        v3 f = evaluate_brdf(mat, rec->normal, V, L);
        v3 radiance = muls(ld3(light->color), (float)light->intensity);
        v3 direct = muls(mulv(radiance, f), NdotL);
        Lo = add(Lo, direct);
    }
    return Lo;
}

static v3 clamp01v(v3 c) { /* clamp, G/shader.h:24-32 */
    if (c.x > 1.0f) c.x = 1.0f; if (c.y > 1.0f) c.y = 1.0f; if (c.z > 1.0f) c.z = 1.0f;
    if (c.x < 0.0f) c.x = 0.0f; if (c.y < 0.0f) c.y = 0.0f; if (c.z < 0.0f) c.z = 0.0f;
    return c;
}

/* rng_next, G/query.h:32-42 (LCG step + Wang-hash style mix, float(h)/float(0xFFFFFFFF)) */
static float rng_next(uint32_t* state) {
    *state = *state * 1664525u + 1013904223u;
    uint32_t h = *state;
    h = (h ^ 61u) ^ (h >> 16u);
    h *= 9u;
    h ^= h >> 4u;
    h *= 0x27d4eb2du;
    h ^= h >> 15u;
    return (float)h / (float)0xFFFFFFFFu;
}
/* make_rng_seed, G/query.h:44-48 */
static uint32_t make_rng_seed(int x, int y, int sample) {
    return (uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)sample * 83492791u;
}
/* random_unit_vector / random_on_hemisphere, G/query.h:50-70 */
static v3 random_unit_vector(uint32_t* state) {
    for (;;) {
        float x = 2.0f * rng_next(state) - 1.0f;
        float y = 2.0f * rng_next(state) - 1.0f;
        float z = 2.0f * rng_next(state) - 1.0f;
        float lensq = x * x + y * y + z * z;
        if (lensq > 1e-10f && lensq <= 1.0f) {
            float inv = 1.0f / sqrtf(lensq);
            return mk(x * inv, y * inv, z * inv);
        }
    }
}
static v3 random_on_hemisphere(v3 normal, uint32_t* state) {
    v3 u = random_unit_vector(state);
    if (dot(u, normal) > 0.0f) return u;
    return mk(-u.x, -u.y, -u.z);
}

/* TraceRayIterative, G/query.h:156-220: closest hit, direct light, then a mirror or (hash-RNG) diffuse bounce
 * until maxDepth, a miss, a non-reflecting material or a throughput below 1e-4.  `first` = depth-0 hit. */
static v3 trace_hw2(const world_t* w, const ray_t* primary, hit_t* first, uint32_t rng_state, orc_counters* c) {
    const float RT_EPS = 1e-3f;   /* G/shader.h:22 */
    v3 radiance = mk(0, 0, 0), throughput = mk(1, 1, 1);
    memset(first, 0, sizeof *first); first->tri = -1;
    if (w->fr->max_depth <= 0) return radiance;
    ray_t ray = *primary;
    for (int depth = 0; depth < w->fr->max_depth; ++depth) {
        hit_t rec;
        if (c) c->rays_primary++;
        closest_hw2(w, &ray, &rec, c);
        if (depth == 0) *first = rec;
        if (!rec.hit) {
            radiance = add(radiance, mulv(throughput, ld3(w->fr->miss_color)));
            break;
        }
        /* assignMaterialToHit, G/query.h:134-153 */
        rt_material mat = kDefaultMaterial;
        if (w->sc->tri_obj_ids && w->sc->materials && rec.tri >= 0 && (uint32_t)rec.tri < w->P) {
            int obj = w->sc->tri_obj_ids[rec.tri];
            if (obj >= 0 && obj < w->sc->num_materials) mat = w->sc->materials[obj];
        }
        v3 direct = shade_direct_hw2(w, &ray, &rec, &mat, c);
        radiance = add(radiance, mulv(throughput, direct));
        const float kd = mat.kd, kr = mat.kr, total = kd + kr;
        if (total <= 0.0f) break;
        const v3 N = normalize(rec.normal);
        const float xi = rng_next(&rng_state);
        if (w->fr->diffuse_bounce && xi < kd / total) {
            v3 diffuse_dir = random_on_hemisphere(N, &rng_state);
            ray.orig = add(rec.p, muls(N, RT_EPS));
            ray.dir = diffuse_dir;
            float NdotL = fmaxf(dot(N, diffuse_dir), 0.0f);
            throughput = mulv(throughput, muls(ld3(mat.albedo), 2.0f * NdotL));
        } else {
            v3 I = unit_vector(ray.dir);
            v3 reflDir = sub(I, muls(N, 2.0f * dot(I, N)));   /* reflect_dir, G/shader.h:38-42 */
            ray.orig = add(rec.p, muls(N, RT_EPS));
            ray.dir = reflDir;
            throughput = mulv(throughput, muls(ld3(mat.specular_color), kr));
        }
        if (throughput.x < 1e-4f && throughput.y < 1e-4f && throughput.z < 1e-4f) break;
    }
    return clamp01v(radiance);
}

/* ------------------------------------------------- HW2/CPUOnly renderer ---- */
/* C = HW2/HW2/CPUOnly/include.  Deterministic subset: point lights, diffuse_bounce == false. */
/* EvaluateBRDF, C/brdf.h:12-37 (fs = specularColor * (ks * specLobe)) */
static v3 evaluate_brdf_c(const rt_material* m, v3 N, v3 V, v3 L) {
    float NdotL = fmaxf(dot(N, L), 0.0f);
    float NdotV = fmaxf(dot(N, V), 0.0f);
    if (NdotL <= 0.f || NdotV <= 0.f) return mk(0, 0, 0);
    const float invPi = 0.31830988618f;
    v3 fd = muls(ld3(m->albedo), m->kd * invPi);
    v3 H = unit_vector_c(add(L, V));
    float NdotH = fmaxf(dot(N, H), 0.0f);
    const float inv2Pi = 0.15915494309f;
    float specNorm = (m->shininess + 2.0f) * inv2Pi;
    float specLobe = specNorm * powf(NdotH, m->shininess);
    v3 fs = muls(ld3(m->specular_color), m->ks * specLobe);
    return add(fd, fs);
}
/* IntersectScene, C/raytracer.h:96-118: t >= t_min and t < closest in double, first triangle wins on equal t. */
static int intersect_scene_c(const world_t* w, const ray_t* r, double t_min, double t_max, hit_t* out, orc_counters* c) {
    int hit_anything = 0;
    double closest = t_max;
    const int have_normals = w->sc->normals != NULL;
    for (uint32_t i = 0; i < w->P; ++i) {
        hit_t tmp;
        if (c) c->tri_tests++;
        if (!mt_cpuonly(r, &w->tris[i], have_normals, &tmp)) continue;
        if ((double)tmp.t >= t_min && (double)tmp.t < closest) {
            hit_anything = 1; closest = (double)tmp.t; tmp.tri = (int)i; *out = tmp;
        }
    }
    return hit_anything;
}
static rt_material material_of(const world_t* w, int tri) {   /* tri.mat, C/src/render.cpp:79-97 */
    rt_material mat = kDefaultMaterial;
    if (w->sc->tri_obj_ids && w->sc->materials) {
        int obj = w->sc->tri_obj_ids[tri];
        if (obj >= 0 && obj < w->sc->num_materials) mat = w->sc->materials[obj];
    } else if (w->sc->materials && w->sc->num_materials > 0) mat = w->sc->materials[0];
    return mat;
}
/* random_in_unit_disk, C/raytracer.h:76-85.  The reference draws from a process-wide std::mt19937 seeded by
   std::random_device (:12-16), consumed in pixel order on one thread: not reproducible.  The restatement takes its
   numbers from the reference's other generator, the per-pixel hash RNG of G/query.h:32-48 (rng_next above), seeded by
   the caller with (x, y, sample) and rt_frame.rng_seed — the same choice as the device (include/rt_api.h, rt_frame). */
static void random_in_unit_disk(uint32_t* state, float* dx, float* dy) {
    for (;;) {
        float x = 2.0f * rng_next(state) - 1.0f;
        float y = 2.0f * rng_next(state) - 1.0f;
        float r2 = x * x + y * y;
        if (r2 > 1e-10f && r2 <= 1.0f) { *dx = x; *dy = y; return; }
    }
}
/* ShadeDirect + ShadowVisibility (point light: one shadow ray; disk light: shadow_samples rays), C/raytracer.h:121-211;
   make_basis :88-93; RT_EPS = 1e-4 (:49) */
static v3 shade_direct_c(const world_t* w, const ray_t* r, const hit_t* rec, const rt_material* mat, uint32_t* rng, orc_counters* c) {
    const float RT_EPS = 1e-4f;
    v3 N = unit_vector_c(rec->normal);
    v3 V = unit_vector_c(sub(r->orig, rec->p));
    v3 Lo = mk(0, 0, 0);
    Lo = add(Lo, muls(ld3(mat->albedo), 0.05f));
    Lo = add(Lo, ld3(mat->emission));
    for (int i = 0; i < w->fr->num_lights; ++i) {
        const rt_light* light = &w->fr->lights[i];
        v3 lpos = ld3(light->position);
        v3 toL = sub(lpos, rec->p);
        float dist = sqrtf(dot(toL, toL));
        if (dist <= 0.0f) continue;
        v3 L = divd(toL, (double)dist);
        float NdotL = fmaxf(dot(N, L), 0.0f);
        if (NdotL <= 0.0f) continue;
        float vis = 1.0f;                                  /* ShadowVisibility, :121-168 */
        if (w->fr->shadows) {
            const float radius = w->fr->light_radius ? w->fr->light_radius[i] : 0.0f;
            int S = (radius > 0.0f && w->fr->light_shadow_samples) ? w->fr->light_shadow_samples[i] : 1;
            if (S < 1) S = 1;
            float unoccluded = 0.0f;
            v3 toC = sub(lpos, rec->p);
            float distC = sqrtf(dot(toC, toC));
            if (distC <= 0.0f) vis = 1.0f;
            else {
                v3 Wd = divd(sub(rec->p, lpos), (double)distC);
                v3 T = mk(0, 0, 0), B = mk(0, 0, 0);
                if (radius > 0.0f) {                       /* make_basis */
                    v3 a = fabsf(Wd.x) > 0.9f ? mk(0, 1, 0) : mk(1, 0, 0);
                    T = unit_vector_c(cross(a, Wd));
                    B = cross(Wd, T);
                }
                for (int k = 0; k < S; ++k) {
                    v3 lightPos = lpos;
                    if (radius > 0.0f) {
                        float dx, dy;
                        random_in_unit_disk(rng, &dx, &dy);
                        lightPos = add(add(lpos, muls(T, dx * radius)), muls(B, dy * radius));
                    }
                    v3 toS = sub(lightPos, rec->p);
                    float distToL = sqrtf(dot(toS, toS));
                    if (distToL <= 0.0f) { unoccluded += 1.0f; continue; }
                    v3 Ldir = divd(toS, (double)distToL);
                    ray_t sray;
                    sray.orig = add(rec->p, muls(N, RT_EPS));
                    sray.dir = unit_vector_c(Ldir);        /* the Ray constructor normalises again, C/ray.h:13-14 */
                    hit_t sh;
                    if (c) c->rays_shadow++;
                    if (!intersect_scene_c(w, &sray, (double)RT_EPS, (double)distToL - (double)RT_EPS, &sh, c)) unoccluded += 1.0f;
                }
                vis = unoccluded / (float)S;
            }
        }
        if (vis <= 0.0f) continue;
        v3 f = evaluate_brdf_c(mat, N, V, L);
        v3 radiance = muls(ld3(light->color), light->intensity_f);
        Lo = add(Lo, muls(mulv(radiance, f), NdotL * vis));
    }
    return Lo;
}
/* TraceRay, C/raytracer.h:215-260, with diffuse_bounce == false: direct light + perfect-mirror recursion. */
static v3 trace_cpuonly(const world_t* w, const ray_t* r, int depth, hit_t* first, uint32_t* rng, orc_counters* c) {
    const float RT_EPS = 1e-4f;
    if (first) { memset(first, 0, sizeof *first); first->tri = -1; }
    if (depth <= 0) return mk(0, 0, 0);
    hit_t rec;
    if (c) c->rays_primary++;
    if (!intersect_scene_c(w, r, (double)RT_EPS, INFINITY, &rec, c)) {
        v3 unit_dir = unit_vector_c(r->dir);
        float t = 0.5f * (unit_dir.z + 1.0f);
        return add(muls(mk(1.0f, 1.0f, 1.0f), 1.0f - t), muls(mk(0.5f, 0.7f, 1.0f), t));
    }
    if (first) *first = rec;
    v3 N = unit_vector_c(rec.normal);
    rt_material mat = material_of(w, rec.tri);
    v3 Lo = shade_direct_c(w, r, &rec, &mat, rng, c);
    float kd = mat.kd, kr = mat.kr, total = kd + kr;
    if (total > 0.0f) {
        /* xi = random_float() is drawn but unused when diffuse_bounce is false */
        if (kr > 0.0f) {
            v3 I = unit_vector_c(r->dir);
            v3 refl = sub(I, muls(N, 2.0f * dot(I, N)));      /* reflect_dir, C/raytracer.h:70-74 */
            ray_t rr;
            rr.orig = add(rec.p, muls(N, RT_EPS));
            rr.dir = unit_vector_c(refl);                      /* Ray constructor */
            v3 bounced = trace_cpuonly(w, &rr, depth - 1, NULL, rng, c);
            v3 tint = ld3(mat.specular_color);
            Lo = add(Lo, muls(mulv(tint, bounced), mat.kr));   /* (diffuse_bounce ? total : kr) * (tint * bounced) */
        }
    }
    return Lo;
}

/* ----------------------------------------------------------- quantisers ---- */
static uint8_t quantise(float c, int q) {
    switch (q) {
    case RT_QUANT_HW1_TRUNC: return (uint8_t)(255.99f * c);                       /* HW1/src/render.cpp:121-123 */
    case RT_QUANT_HW2_TRUNC: return (uint8_t)(255.0f * (c < 1.0f ? c : 1.0f));    /* GPUandCPU/src/main.cu:428-430 */
    case RT_QUANT_CPU_TRUNC: { float x = c > 1.0f ? 1.0f : c; if (x < 0.0f) x = 0.0f; return (uint8_t)(255.99f * x); } /* CPUOnly/src/render.cpp:157-163 */
    default: {                                                                    /* ppm_p6.cpp:137-155 */
        double x = (double)c;
        if (q == RT_QUANT_PPM_GAMMA2) { if (x < 0.0) x = 0.0; x = sqrt(x); }
        if (x < 0.0) x = 0.0; if (x > 1.0) x = 1.0;
        long r = lround(x * 255.0);
        if (r < 0) r = 0; if (r > 255) r = 255;
        return (uint8_t)r;
    }
    }
}

/* ------------------------------------------------------------ pixel loop ---- */
typedef struct {
    world_t w; rt_image* img; int row_begin, row_step, tid, nthreads; orc_counters cnt;
} job_t;

static void render_pixel(const world_t* w, int x, int y, rt_image* img, orc_counters* c) {
    const rt_frame* fr = w->fr;
    const int W = fr->width;
    v3 center = ld3(fr->cam.center), p00 = ld3(fr->cam.pixel00_loc);
    v3 du = ld3(fr->cam.pixel_delta_u), dv = ld3(fr->cam.pixel_delta_v);
    v3 accum = mk(0, 0, 0);
    hit_t first; memset(&first, 0, sizeof first); first.tri = -1;
    for (int s = 0; s < fr->spp; ++s) {
        float jx = fr->jitter ? fr->jitter[2 * s] : 0.0f, jy = fr->jitter ? fr->jitter[2 * s + 1] : 0.0f;
        float px = (float)x + jx, py = (float)y + jy;
        ray_t ray; hit_t rec; v3 color;
        if (fr->mode == RT_MODE_HW1) {
            /* HW1/src/render.cpp:79-107.  get_pixel_position(int,int) truncates the jitter (quirk Q1);
             * Ray ctor normalises (H1/ray.h:25). */
            int ix = (int)px, iy = (int)py;
            v3 pp = add(add(p00, muls(du, (float)(double)ix)), muls(dv, (float)(double)iy));
            ray.orig = center;
            ray.dir = unit_vector(sub(pp, center));
            memset(&rec, 0, sizeof rec); rec.tri = -1;
            float prev_t = FLT_MAX;
            if (c) c->rays_primary++;
            for (uint32_t k = 0; k < w->P; ++k) {
                hit_t h;
                if (c) c->tri_tests++;
                if (mt_hw1(&ray, &w->tris[k], &h) && (double)h.t < (double)prev_t) { h.tri = (int)k; rec = h; prev_t = h.t; }
            }
            color = shade_hw1(&ray, &rec, &fr->lights[0]);
        } else if (fr->mode == RT_MODE_HW2_CPU) {
            /* CPUOnly/src/render.cpp:124-135: u = double(i) + du; get_pixel_position(double,double) narrows each to float
             * (C/camera.h:41-43); Ray constructor normalises (C/ray.h:13-14). */
            float fu = (float)((double)x + (double)jx), fv = (float)((double)y + (double)jy);
            v3 pp = add(add(p00, muls(du, fu)), muls(dv, fv));
            ray.orig = center;
            ray.dir = unit_vector_c(sub(pp, center));
            int depth = fr->max_depth > 1 ? fr->max_depth : 1;    /* std::max(1, max_bounces), render.cpp:113 */
            uint32_t rng = make_rng_seed(x, y, s) ^ fr->rng_seed;     /* hash RNG for the disk-light samples (see random_in_unit_disk) */
            color = trace_cpuonly(w, &ray, depth, &rec, &rng, c);
        } else {
            /* Camera::get_ray(float,float), G/camera.h:49-53; CPU loop G/query.cu:136-165 */
            v3 pp = add(add(p00, muls(du, px)), muls(dv, py));
            ray.orig = center;
            ray.dir = cam_unit_vector(sub(pp, center));
            color = trace_hw2(w, &ray, &rec, make_rng_seed(x, y, s), c);   /* rng seed: G/query.cu:151 */
        }
        if (s == 0) first = rec;
        accum = add(accum, color);
    }
    v3 fin = divd(accum, (double)(float)fr->spp); /* col/float(spp): render.cpp:110, query.cu:163 */
    size_t pix = (size_t)y * W + x;
    if (img->rgb) { img->rgb[3 * pix] = fin.x; img->rgb[3 * pix + 1] = fin.y; img->rgb[3 * pix + 2] = fin.z; }
    if (img->rgb8) {
        img->rgb8[3 * pix] = quantise(fin.x, fr->quantiser);
        img->rgb8[3 * pix + 1] = quantise(fin.y, fr->quantiser);
        img->rgb8[3 * pix + 2] = quantise(fin.z, fr->quantiser);
    }
    if (img->tri_id) img->tri_id[pix] = first.hit ? first.tri : -1;
    if (img->t) img->t[pix] = first.hit ? first.t : -1.0f;
}

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    const rt_frame* fr = j->w.fr;
    int k = 0;
    for (int y = j->row_begin; y < fr->height; y += j->row_step, ++k) {
        if (k % j->nthreads != j->tid) continue;
        for (int x = 0; x < fr->width; ++x) render_pixel(&j->w, x, y, j->img, &j->cnt);
    }
    return NULL;
}

/* Renders rows row_begin, row_begin+row_step, ... of the frame (other rows of the
 * output planes are left untouched) on `nthreads` host threads.  bvh may be NULL
 * (brute force).  Returns RT_OK. */
int orc_render(const rt_scene* sc, const rt_frame* fr, const orc_bvh* bvh, rt_image* img,
               int nthreads, int row_begin, int row_step, orc_counters* counters)
{
    if (!sc || !fr || !img || fr->width < 1 || fr->height < 1 || fr->spp < 1) return RT_ERR_ARG;
    if (fr->mode == RT_MODE_HW1 && fr->num_lights < 1) return RT_ERR_ARG;
    if (nthreads < 1) nthreads = 1;
    if (row_step < 1) row_step = 1;
    tri_t* tris = gather_triangles(sc);
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; ++i) {
        jobs[i].w.sc = sc; jobs[i].w.fr = fr; jobs[i].w.tris = tris; jobs[i].w.bvh = bvh;
        jobs[i].w.P = (uint32_t)sc->num_triangles;
        jobs[i].img = img; jobs[i].row_begin = row_begin; jobs[i].row_step = row_step;
        jobs[i].tid = i; jobs[i].nthreads = nthreads;
        if (nthreads > 1) pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    if (nthreads == 1) worker(&jobs[0]);
    else for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    img->width = fr->width; img->height = fr->height;
    orc_counters tot; memset(&tot, 0, sizeof tot);
    for (int i = 0; i < nthreads; ++i) {
        tot.rays_primary += jobs[i].cnt.rays_primary; tot.rays_shadow += jobs[i].cnt.rays_shadow;
        tot.node_pops += jobs[i].cnt.node_pops; tot.box_tests += jobs[i].cnt.box_tests;
        tot.tri_tests += jobs[i].cnt.tri_tests;
        if (jobs[i].cnt.max_stack > tot.max_stack) tot.max_stack = jobs[i].cnt.max_stack;
    }
    img->rays_primary = tot.rays_primary; img->rays_shadow = tot.rays_shadow;
    if (counters) *counters = tot;
    free(jobs); free(th); free(tris);
    return RT_OK;
}

/* Single ray / single triangle probes for the unit vectors of
 * HW1/test_ray_tri_inter_STANDALONE/test_ray_triangle_inter.cpp:17-126.
 * contract: 0 = HW1 ray_intersection, 1 = HW2 intersectTriangle(tmin=1e-4,tmax=FLT_MAX),
 * 2 = CPUOnly ray_intersection.  dir is normalised with the Ray ctor rule when normalise != 0. */
int orc_ray_triangle(int contract, const float orig[3], const float dir[3], int normalise,
                     const float v0[3], const float v1[3], const float v2[3], float* t_out)
{
    ray_t r; r.orig = ld3(orig); r.dir = ld3(dir);
    if (normalise) r.dir = unit_vector(r.dir);
    tri_t tri; memset(&tri, 0, sizeof tri);
    tri.v0 = ld3(v0); tri.v1 = ld3(v1); tri.v2 = ld3(v2);
    hit_t h; memset(&h, 0, sizeof h);
    int hit = contract == 0 ? mt_hw1(&r, &tri, &h) : contract == 1 ? mt_hw2(&r, &tri, 1e-4f, FLT_MAX, &h) : mt_cpuonly(&r, &tri, 1, &h);
    if (t_out) *t_out = hit ? h.t : -1.0f;
    return hit;
}

uint8_t orc_quantise(float c, int q) { return quantise(c, q); }
