// scene_json.hpp — the scene-file dialect of the reference's BVH renderer
// (HW2/HW2/GPUandCPU/include/scene.h:242-380; keys listed in SURVEY §5 "Config / flags"):
//   settings{max_bounces, spp, diffuse_bounce}, miss_color, camera{focal_length_mm, sensor_height_mm,
//   pixel_width, pixel_height, position, look_at, up}, lights[] | light{position,color,intensity},
//   scene[]{name,type,path,transform{position,rotation,scale},material{albedo,kd,ks,shininess,
//   specular_color,kr,emission}}.
// Own recursive-descent JSON reader (the reference hand-writes one too, scene.h:47-217).
#pragma once

#include <string>
#include <vector>

#include "../../include/rt_api.h"

namespace rtb200 {

struct SceneObjectDesc {
    std::string name, type, path;
    float position[3] = {0, 0, 0}, rotation[3] = {0, 0, 0}, scale[3] = {1, 1, 1};
    rt_material material;
};

struct SceneDesc {
    int max_depth = 1, spp = 1;
    bool diffuse_bounce = true;
    float miss_color[3] = {0, 0, 0};
    // Camera() defaults of GPUandCPU/include/camera.h:13-19
    float cam_pos[3] = {0, 0, 0}, cam_look_at[3] = {0, 1, 0}, cam_up[3] = {0, 0, 1};
    double focal_length_mm = 50.0, sensor_height_mm = 24.0;
    int pixel_width = 100, pixel_height = 100;
    std::vector<rt_light> lights;
    std::vector<SceneObjectDesc> objects;
};

rt_material default_material();   // Material(), GPUandCPU/include/material.h:6-20
bool load_scene_file(const std::string& path, SceneDesc& out, std::string* err);
bool parse_scene_text(const std::string& text, SceneDesc& out, std::string* err);

} // namespace rtb200
