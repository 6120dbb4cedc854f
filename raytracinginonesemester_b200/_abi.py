"""ctypes mirror of include/rt_api.h (struct layouts and constants)."""
import ctypes as C

RT_API_VERSION = 5          # include/rt_api.h
RT_OK, RT_ERR_ARG, RT_ERR_CUDA, RT_ERR_STATE, RT_ERR_NCCL, RT_ERR_NOMEM, RT_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
RT_MODE_HW1, RT_MODE_HW2_BVH, RT_MODE_HW2_CPU = 0, 1, 2
RT_ACCEL_BRUTE, RT_ACCEL_BVH = 0, 1
RT_OUT_RGB_F32, RT_OUT_RGB8, RT_OUT_TRI_ID, RT_OUT_T = 1, 2, 4, 8
RT_QUANT_PPM_LROUND, RT_QUANT_PPM_GAMMA2, RT_QUANT_HW1_TRUNC, RT_QUANT_HW2_TRUNC, RT_QUANT_CPU_TRUNC = 0, 1, 2, 3, 4
RT_BUILD_DEFAULT, RT_BUILD_NO_BVH = 0, 1
RT_VARIANT_DEFAULT, RT_VARIANT_PACKET_OCC6, RT_VARIANT_PACKET_OCC10, RT_VARIANT_PACKET_EXACT_SLAB, RT_VARIANT_PER_RAY = 0, 1, 2, 3, 10
RT_VARIANT_PACKET_PREFETCH, RT_VARIANT_PACKET_PIXEL_MAJOR = 4, 5
RT_VARIANT_STATS, RT_VARIANT_PER_RAY_STATS = 100, 110
RT_VARIANT_FRUSTUM, RT_VARIANT_FRUSTUM_STATS, RT_VARIANT_PACKET, RT_VARIANT_PACKET_STATS = 6, 106, 7, 107
RT_VARIANT_PERSIST, RT_VARIANT_PERSIST_EXACT_MT, RT_VARIANT_PERSIST_OCC8, RT_VARIANT_PERSIST_OCC10 = 8, 9, 11, 12
RT_GATHER_AUTO, RT_GATHER_NCCL, RT_GATHER_PEER = 0, 1, 2


def RT_BUILD_LEAF_MAX(n):
    return (int(n) & 0xF) << 8


f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


class rt_material(C.Structure):
    _fields_ = [("albedo", C.c_float * 3), ("kd", C.c_float), ("specular_color", C.c_float * 3),
                ("ks", C.c_float), ("shininess", C.c_float), ("kr", C.c_float), ("emission", C.c_float * 3)]


class _rt_intensity(C.Union):
    _fields_ = [("intensity", C.c_int32), ("intensity_f", C.c_float)]


class rt_light(C.Structure):
    _anonymous_ = ("u",)
    _fields_ = [("position", C.c_float * 3), ("color", C.c_float * 3), ("u", _rt_intensity)]


class rt_object_transform(C.Structure):
    _fields_ = [("first_vertex", C.c_uint64), ("num_vertices", C.c_uint64), ("position", C.c_float * 3),
                ("rotation_deg", C.c_float * 3), ("scale", C.c_float * 3)]


class rt_scene(C.Structure):
    _fields_ = [("positions", f32p), ("normals", f32p), ("num_vertices", C.c_uint64),
                ("indices", u32p), ("num_triangles", C.c_uint64), ("tri_obj_ids", i32p),
                ("materials", C.POINTER(rt_material)), ("num_materials", C.c_int32), ("build_flags", C.c_uint32),
                ("transforms", C.POINTER(rt_object_transform)), ("num_transforms", C.c_int32)]


class rt_camera(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("pixel00_loc", C.c_float * 3),
                ("pixel_delta_u", C.c_float * 3), ("pixel_delta_v", C.c_float * 3)]


class rt_frame(C.Structure):
    _fields_ = [("mode", C.c_int32), ("accel", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("cam", rt_camera), ("lights", C.POINTER(rt_light)), ("num_lights", C.c_int32),
                ("miss_color", C.c_float * 3), ("spp", C.c_int32), ("jitter", f32p),
                ("max_depth", C.c_int32), ("shadows", C.c_int32), ("outputs", C.c_uint32),
                ("quantiser", C.c_int32), ("kernel_variant", C.c_int32), ("diffuse_bounce", C.c_int32),
                ("light_radius", f32p), ("light_shadow_samples", i32p), ("rng_seed", C.c_uint32)]


class rt_image(C.Structure):
    _fields_ = [("rgb", f32p), ("rgb8", u8p), ("tri_id", i32p), ("t", f32p),
                ("width", C.c_int32), ("height", C.c_int32),
                ("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("gpu_ms", C.c_float)]


class rt_build_info(C.Structure):
    _fields_ = [("num_triangles", C.c_uint64), ("num_nodes", C.c_uint64), ("num_leaves", C.c_uint64),
                ("arena_bytes", C.c_uint64), ("build_ms", C.c_float), ("upload_ms", C.c_float),
                ("scene_min", C.c_float * 3), ("scene_max", C.c_float * 3)]


assert C.sizeof(rt_material) == 52 and C.sizeof(rt_light) == 28 and C.sizeof(rt_camera) == 48

# name -> (restype, argtypes): every symbol include/rt_api.h declares
EXPORTS = {
    "rt_api_version": (C.c_int, []),
    "rt_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "rt_create_multi": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]),
    "rt_destroy": (C.c_int, [C.c_void_p]),
    "rt_last_error": (C.c_char_p, [C.c_void_p]),
    "rt_comm_unique_id": (C.c_int, [C.c_void_p]),
    "rt_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rt_comm_rank": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rt_comm_set_gather": (C.c_int, [C.c_void_p, C.c_int]),
    "rt_comm_set_sharding": (C.c_int, [C.c_void_p, C.c_int]),
    "rt_comm_gather_mode": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "rt_comm_set_timeout": (C.c_int, [C.c_void_p, C.c_double]),
    "rt_frame_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "rt_stream_handle": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "rt_upload_scene": (C.c_int, [C.c_void_p, C.POINTER(rt_scene)]),
    "rt_build_info_get": (C.c_int, [C.c_void_p, C.POINTER(rt_build_info)]),
    "rt_render": (C.c_int, [C.c_void_p, C.POINTER(rt_frame)]),
    "rt_render_into": (C.c_int, [C.c_void_p, C.POINTER(rt_frame), C.POINTER(rt_image)]),
    "rt_download_image": (C.c_int, [C.c_void_p, C.POINTER(rt_image)]),
    "rt_host_image_create": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "rt_host_image_destroy": (C.c_int, [C.c_void_p]),
    "rt_sync": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "rt_frame_stats": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4),
    "rt_camera_init": (C.c_int, [C.POINTER(rt_camera), f32p, f32p, f32p, C.c_double, C.c_double, C.c_int, C.c_int]),
    "rt_camera_init_cpuonly": (C.c_int, [C.POINTER(rt_camera), f32p, f32p, f32p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]),
    "rt_jitter_table": (C.c_int, [f32p, C.c_int, C.c_uint32, C.c_int]),
    "rt_mesh_load_obj": (C.c_int, [C.c_char_p, i32p, C.POINTER(C.c_void_p)]),
    "rt_mesh_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "rt_mesh_free": (None, [C.c_void_p]),
    "rt_mesh_counts": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3),
    "rt_mesh_copy": (C.c_int, [C.c_void_p, f32p, f32p, u32p, i32p]),
    "rt_mesh_transform": (C.c_int, [C.c_void_p, f32p, f32p, f32p]),
    "rt_mesh_append": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_mesh_last_error": (C.c_char_p, []),
    "rt_device_of": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "rt_dmesh_parse_obj": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, i32p, C.POINTER(C.c_void_p)]),
    "rt_dmesh_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "rt_dmesh_free": (None, [C.c_void_p]),
    "rt_dmesh_counts": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3),
    "rt_dmesh_arrays": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "rt_dmesh_copy": (C.c_int, [C.c_void_p, f32p, f32p, u32p, i32p]),
    "rt_dmesh_append": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_dmesh_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "rt_dmesh_last_error": (C.c_char_p, []),
    "rt_debug_set_shard": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rt_debug_download_bvh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, i32p]),
}


def bind(lib):
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
