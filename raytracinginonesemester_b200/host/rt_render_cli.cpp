// rt_render_cli — C++ host driver over the C ABI: a drop-in for the reference's three command lines.
//
//   rt_render_cli [--hw1] [mesh.obj ...]        HW1 renderer   (HW1/src/render.cpp:15-136)
//   rt_render_cli scene.json | mesh.obj ...     BVH renderer   (HW2/HW2/GPUandCPU/src/main.cu:98-436)
//
// Same inputs (OBJ path(s) or a scene JSON/.scene), same defaults (HW1: camera (0,-1,1)->(0,.15,0),
// 255 mm, 320x180, light (-3,0,1) magenta; BVH: Camera(), fallback light (-3,0,1) x1), same timing prints;
// the image is written as binary PPM P6 through the reference's own ppm_p6_lib when the build links it
// (-DPPM_P6_DIR=<reference>/HW1/ppm_p6_lib), else by a minimal built-in writer of the device-quantised
// bytes (identical bytes for maxval 255: tests/test_gpu_parity.py::test_ppm_bytes_match_reference_writer).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/rt_api.h"
#include "mesh_ingest.hpp"
#include "scene_json.hpp"

#ifdef RT_HAVE_PPM_P6
#include "ppm_p6.hpp"
#endif

using namespace rtb200;

namespace {

bool ends_with(const std::string& s, const char* suf) {
    size_t n = std::strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
std::string dirname_of(const std::string& p) {
    size_t k = p.find_last_of("/\\");
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
}
bool file_exists(const std::string& p) { std::ifstream f(p); return (bool)f; }

int die(rt_ctx* ctx, const char* what) {
    std::fprintf(stderr, "%s: %s\n", what, rt_last_error(ctx));
    return 1;
}

bool write_ppm(const std::string& path, int W, int H, const std::vector<float>& rgb, const std::vector<uint8_t>& rgb8, bool gamma2) {
#ifdef RT_HAVE_PPM_P6
    (void)rgb8;
    ppm_p6::Image img(W, H);
    auto& px = img.pixels();
    for (size_t i = 0; i < (size_t)W * H; ++i) { px[i].r = rgb[3 * i]; px[i].g = rgb[3 * i + 1]; px[i].b = rgb[3 * i + 2]; }
    ppm_p6::WriteOptions opt;
    opt.maxval = 255; opt.clamp = true; opt.gamma2 = gamma2; opt.flip_y = false;
    std::string err;
    if (!ppm_p6::write_p6(path, img, opt, &err)) { std::fprintf(stderr, "write_p6: %s\n", err.c_str()); return false; }
    return true;
#else
    (void)rgb; (void)gamma2;
    std::ofstream out(path, std::ios::binary | std::ios::trunc);
    if (!out) return false;
    out << "P6\n" << W << " " << H << "\n255\n";
    out.write(reinterpret_cast<const char*>(rgb8.data()), (std::streamsize)rgb8.size());
    return (bool)out;
#endif
}

} // namespace

int main(int argc, char** argv) {
    bool hw1 = false, gamma2 = false, brute = false, host_transform = false, device_ingest = false;
    int device = 0, width = 0, height = 0, spp_override = 0, depth_override = 0;
    std::string out_path;
    std::vector<std::string> inputs;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--hw1") hw1 = true;
        else if (a == "--gamma2") gamma2 = true;
        else if (a == "--brute") brute = true;
        else if (a == "--host-transform") host_transform = true;     // bake object transforms on the host instead of the device
        else if (a == "--device-ingest") device_ingest = true;       // parse the OBJ files on the GPU (rt_dmesh_parse_obj): no host copy of the mesh
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "--width" && i + 1 < argc) width = std::atoi(argv[++i]);
        else if (a == "--height" && i + 1 < argc) height = std::atoi(argv[++i]);
        else if (a == "--spp" && i + 1 < argc) spp_override = std::atoi(argv[++i]);
        else if (a == "--depth" && i + 1 < argc) depth_override = std::atoi(argv[++i]);
        else if ((a == "-o" || a == "--out") && i + 1 < argc) out_path = argv[++i];
        else if (a == "-h" || a == "--help") {
            std::printf("usage: rt_render_cli [--hw1] [--brute] [--host-transform | --device-ingest] [--width W --height H] [--spp N] [--depth D] [--gamma2] [--device D] [-o out.ppm] [scene.json | mesh.obj ...]\n");
            return 0;
        } else inputs.push_back(a);
    }

    SceneDesc sd;
    bool has_scene = false;
    std::vector<SceneObjectDesc> objects;
    if (!inputs.empty() && (ends_with(inputs[0], ".json") || ends_with(inputs[0], ".scene"))) {
        std::string err;
        if (!load_scene_file(inputs[0], sd, &err)) { std::fprintf(stderr, "Failed to load scene: %s\n", err.c_str()); return 1; }
        has_scene = true;
        const std::string base = dirname_of(inputs[0]), project = dirname_of(dirname_of(base));
        for (auto o : sd.objects) {                       // path resolution order of main.cu:126-150
            if (!o.type.empty() && o.type != "mesh") continue;
            if (!o.path.empty() && o.path[0] != '/') {
                std::string rel = o.path.rfind("./", 0) == 0 ? o.path.substr(2) : o.path;
                std::string a = base + "/" + o.path, c = project + "/" + rel;
                if (file_exists(a)) o.path = a; else if (file_exists(o.path)) {} else if (file_exists(c)) o.path = c; else o.path = a;
            }
            objects.push_back(o);
        }
    } else {
        if (inputs.empty()) inputs.push_back(hw1 ? "../assets/meshes/sphere.obj" : "../assets/meshes/frog.obj");
        for (auto& p : inputs) { SceneObjectDesc o; o.path = p; o.material = default_material(); objects.push_back(o); }
    }

    rt_ctx* ctx = nullptr;
    if (rt_create(&ctx, device) != RT_OK) return die(nullptr, "rt_create");
    HostMesh mesh;
    rt_dmesh* dmesh = nullptr;                                // --device-ingest: the mesh lives here instead
    uint64_t dnv = 0, dnt = 0;
    if (device_ingest && rt_dmesh_create(&dmesh) != RT_OK) return die(ctx, "rt_dmesh_create");
    std::vector<rt_material> materials;
    std::vector<rt_object_transform> transforms;             // applyObjectTransform runs on the device at upload
    int next_id = 0;
    for (auto& o : objects) {
        std::printf("Loading OBJ: %s\n", o.path.c_str());
        HostMesh part;
        const int first = next_id;
        std::string err;
        if (device_ingest) {
            std::string bytes;
            if (FILE* f = std::fopen(o.path.c_str(), "rb")) {
                char buf[1 << 16];
                size_t k;
                while ((k = std::fread(buf, 1, sizeof buf, f)) > 0) bytes.append(buf, k);
                std::fclose(f);
            } else { std::fprintf(stderr, "Failed to load OBJ: %s (cannot open)\n", o.path.c_str()); continue; }
            rt_dmesh* part_d = nullptr;
            int32_t nid = next_id;
            if (rt_dmesh_parse_obj(ctx, bytes.data(), bytes.size(), &nid, &part_d) != RT_OK) {
                std::fprintf(stderr, "Failed to load OBJ: %s (%s)\n", o.path.c_str(), rt_dmesh_last_error());
                continue;
            }
            next_id = nid;
            uint64_t pv = 0, pn = 0, pt = 0;
            rt_dmesh_counts(part_d, &pv, &pn, &pt);
            rt_object_transform t{};
            t.first_vertex = dnv; t.num_vertices = pv;
            std::memcpy(t.position, o.position, sizeof t.position); std::memcpy(t.rotation_deg, o.rotation, sizeof t.rotation_deg);
            std::memcpy(t.scale, o.scale, sizeof t.scale);
            transforms.push_back(t);
            if (rt_dmesh_append(dmesh, part_d) != RT_OK) { std::fprintf(stderr, "rt_dmesh_append: %s\n", rt_dmesh_last_error()); return 1; }
            rt_dmesh_free(part_d);
            dnv += pv; dnt += pt;
            materials.resize((size_t)next_id, default_material());
            for (int id = first; id < next_id; ++id) materials[(size_t)id] = o.material;
            std::printf("  -> Loaded %llu triangles.\n", (unsigned long long)pt);
            continue;
        }
        if (!load_obj(o.path, part, next_id, &err)) { std::fprintf(stderr, "Failed to load OBJ: %s (%s)\n", o.path.c_str(), err.c_str()); continue; }
        if (host_transform) transform_mesh(part, o.position, o.rotation, o.scale);
        else {
            rt_object_transform t{};
            t.first_vertex = mesh.num_vertices(); t.num_vertices = part.num_vertices();
            std::memcpy(t.position, o.position, sizeof t.position); std::memcpy(t.rotation_deg, o.rotation, sizeof t.rotation_deg);
            std::memcpy(t.scale, o.scale, sizeof t.scale);
            transforms.push_back(t);
        }
        materials.resize((size_t)next_id, default_material());
        for (int id = first; id < next_id; ++id) materials[(size_t)id] = o.material;
        std::printf("  -> Loaded %zu triangles.\n", part.num_triangles());
        append_mesh(mesh, part);
    }
    if (device_ingest ? dnt == 0 : mesh.positions.empty()) { std::fprintf(stderr, "No valid geometry loaded.\n"); return 1; }
    if (hw1 && !device_ingest && mesh.normals.empty()) mesh.normals.assign(mesh.positions.size(), 0.f);

    rt_scene sc{};
    if (device_ingest) {                                       // device pointers: rt_upload_scene copies device to device
        rt_dmesh_arrays(dmesh, &sc.positions, &sc.normals, &sc.indices, &sc.tri_obj_ids);
        sc.num_vertices = dnv; sc.num_triangles = dnt;
    } else {
        sc.positions = mesh.positions.data(); sc.normals = mesh.normals.empty() ? nullptr : mesh.normals.data();
        sc.num_vertices = mesh.num_vertices(); sc.indices = mesh.indices.data(); sc.num_triangles = mesh.num_triangles();
        sc.tri_obj_ids = mesh.tri_obj_ids.data();
    }
    sc.materials = materials.data(); sc.num_materials = (int)materials.size();
    sc.build_flags = (hw1 && brute) ? RT_BUILD_NO_BVH : RT_BUILD_DEFAULT;
    sc.transforms = transforms.empty() ? nullptr : transforms.data(); sc.num_transforms = (int)transforms.size();
    if (rt_upload_scene(ctx, &sc) != RT_OK) return die(ctx, "rt_upload_scene");
    if (dmesh) rt_dmesh_free(dmesh);
    rt_build_info bi{};
    rt_build_info_get(ctx, &bi);
    std::printf("GPU LBVH Build Time: %.3f ms (%llu triangles, %llu nodes)\n", bi.build_ms, (unsigned long long)bi.num_triangles, (unsigned long long)bi.num_nodes);

    rt_frame fr{};
    std::vector<rt_light> lights;
    std::vector<float> jitter;
    if (hw1) {
        const float pos[3] = {0.0f, -1.0f, 1.0f}, look[3] = {0.0f, 0.15f, 0.0f}, up[3] = {0, 0, 1};
        fr.width = width ? width : 320; fr.height = height ? height : 180;
        if (rt_camera_init(&fr.cam, pos, look, up, 255.0, 24.0, fr.width, fr.height) != RT_OK) { std::fprintf(stderr, "Error: pixel_width/pixel_height must be >= 1\n"); return 1; }
        rt_light l{}; l.position[0] = -3.0f; l.position[2] = 1.0f; l.color[0] = 1.0f; l.color[2] = 1.0f; l.intensity = 1;
        lights.push_back(l);
        fr.mode = RT_MODE_HW1; fr.accel = brute ? RT_ACCEL_BRUTE : RT_ACCEL_BVH;
        fr.spp = 1; jitter.resize(2); rt_jitter_table(jitter.data(), 1, 42u, 0);
        fr.quantiser = RT_QUANT_HW1_TRUNC; fr.max_depth = 1;
    } else {
        fr.width = width ? width : sd.pixel_width; fr.height = height ? height : sd.pixel_height;
        if (fr.width < 1) fr.width = 1;                   // Camera clamps (camera.h:73-74)
        if (fr.height < 1) fr.height = 1;
        rt_camera_init(&fr.cam, sd.cam_pos, sd.cam_look_at, sd.cam_up, sd.focal_length_mm, sd.sensor_height_mm, fr.width, fr.height);
        lights = sd.lights;
        if (lights.empty()) {
            rt_light l{}; l.position[0] = -3.0f; l.position[2] = 1.0f; l.color[0] = l.color[1] = l.color[2] = 1.0f; l.intensity = 1;
            lights.push_back(l);
            std::printf("No lights in scene, using fallback light.\n");
        }
        fr.mode = RT_MODE_HW2_BVH; fr.accel = brute ? RT_ACCEL_BRUTE : RT_ACCEL_BVH;
        fr.spp = spp_override ? spp_override : (has_scene ? sd.spp : 1);
        jitter.resize(2 * (size_t)fr.spp); rt_jitter_table(jitter.data(), fr.spp, 42u, 1);
        std::memcpy(fr.miss_color, sd.miss_color, sizeof fr.miss_color);
        fr.max_depth = depth_override > 0 ? depth_override : (has_scene ? sd.max_depth : 1);   // main.cu:322-324
        fr.diffuse_bounce = has_scene ? (sd.diffuse_bounce ? 1 : 0) : 1;
        fr.shadows = 1; fr.quantiser = RT_QUANT_HW2_TRUNC;
    }
    fr.lights = lights.data(); fr.num_lights = (int)lights.size(); fr.jitter = jitter.data();
#ifdef RT_HAVE_PPM_P6
    fr.outputs = RT_OUT_RGB_F32;
#else
    fr.outputs = RT_OUT_RGB8; fr.quantiser = gamma2 ? RT_QUANT_PPM_GAMMA2 : RT_QUANT_PPM_LROUND;
#endif
    std::vector<float> rgb;
    std::vector<uint8_t> rgb8;
    rt_image img{};
    if (fr.outputs & RT_OUT_RGB_F32) { rgb.resize(3 * (size_t)fr.width * fr.height); img.rgb = rgb.data(); }
    if (fr.outputs & RT_OUT_RGB8) { rgb8.resize(3 * (size_t)fr.width * fr.height); img.rgb8 = rgb8.data(); }

    auto t0 = std::chrono::high_resolution_clock::now();
    if (rt_render(ctx, &fr) != RT_OK) return die(ctx, "rt_render");
    if (rt_download_image(ctx, &img) != RT_OK) return die(ctx, "rt_download_image");
    auto t1 = std::chrono::high_resolution_clock::now();
    std::printf("GPU Render Time: %.3f ms (device %.3f ms, %llu primary + %llu shadow rays)\n",
                std::chrono::duration<double, std::milli>(t1 - t0).count(), img.gpu_ms,
                (unsigned long long)img.rays_primary, (unsigned long long)img.rays_shadow);

    if (out_path.empty()) out_path = hw1 ? "output.ppm" : "render.ppm";
    if (!write_ppm(out_path, fr.width, fr.height, rgb, rgb8, gamma2)) { std::fprintf(stderr, "cannot write %s\n", out_path.c_str()); return 1; }
    std::printf("Image saved to %s\n", out_path.c_str());
    rt_destroy(ctx);
    return 0;
}
