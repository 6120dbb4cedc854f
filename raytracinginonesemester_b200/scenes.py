"""Synthetic scenes and the named benchmark configurations.

terrain(nx, ny, seed): vertices on a regular grid x = -1 + 2i/nx, y = -0.5625 + 1.125 j/ny (z-up),
z = 0.05 sin(9x) cos(7y) + 0.02 sin(31x + 17y) + 0.004 h(i,j) with h a Wang hash of the cell,
evaluated in double and stored as float32; two triangles per cell (v00,v10,v11), (v00,v11,v01);
no vertex normals (geometric-normal fallback, query.h:119-121); one material.
C4 = terrain(1000, 500, 42) = exactly 1 000 000 triangles; C5 = terrain(2500, 2000, 42) = 10 000 000.
"""
import numpy as np

from . import _abi as A
from .api import Frame, Scene, camera_init, camera_init_cpuonly, jitter_table, make_light, make_light_f, make_material


def wang_hash(seed):
    """Integer hash used by the reference (GPUandCPU/include/antialias.h:30-37, query.h:32-42)."""
    s = np.asarray(seed, dtype=np.uint32).copy()
    s = (s ^ np.uint32(61)) ^ (s >> np.uint32(16))
    s = (s * np.uint32(9)).astype(np.uint32)
    s = s ^ (s >> np.uint32(4))
    s = (s * np.uint32(0x27D4EB2D)).astype(np.uint32)
    s = s ^ (s >> np.uint32(15))
    return s


def terrain(nx, ny, seed=42):
    i = np.arange(nx + 1, dtype=np.uint32)[None, :]
    j = np.arange(ny + 1, dtype=np.uint32)[:, None]
    with np.errstate(over="ignore"):
        key = (i * np.uint32(73856093)) ^ (j * np.uint32(19349663)) ^ np.uint32((seed * 83492791) & 0xFFFFFFFF)
    h = wang_hash(key).astype(np.float64) / 4294967295.0
    x = -1.0 + 2.0 * np.arange(nx + 1, dtype=np.float64)[None, :] / nx
    y = -0.5625 + 1.125 * np.arange(ny + 1, dtype=np.float64)[:, None] / ny
    z = 0.05 * np.sin(9.0 * x) * np.cos(7.0 * y) + 0.02 * np.sin(31.0 * x + 17.0 * y) + 0.004 * h
    pos = np.empty((ny + 1, nx + 1, 3), np.float32)
    pos[..., 0] = np.broadcast_to(x, z.shape)
    pos[..., 1] = np.broadcast_to(y, z.shape)
    pos[..., 2] = z
    jj, ii = np.meshgrid(np.arange(ny, dtype=np.uint32), np.arange(nx, dtype=np.uint32), indexing="ij")
    v00 = jj * np.uint32(nx + 1) + ii
    v10, v01, v11 = v00 + 1, v00 + np.uint32(nx + 1), v00 + np.uint32(nx + 2)
    tris = np.stack([np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)], axis=2)  # (ny, nx, 2, 3)
    return pos.reshape(-1, 3), tris.reshape(-1, 3).astype(np.uint32)


TERRAIN_MATERIAL = dict(albedo=(0.6, 0.55, 0.5), kd=1.0, ks=0.3, specular_color=(0.04, 0.04, 0.04), shininess=32.0, kr=0.0)


def terrain_scene(nx, ny, seed=42, build_flags=0):
    pos, idx = terrain(nx, ny, seed)
    return Scene(pos, idx, normals=None, tri_obj_ids=np.zeros(idx.shape[0], np.int32),
                 materials=[make_material(**TERRAIN_MATERIAL)], build_flags=build_flags)


def terrain_frame(width, height, spp=1, shadows=True, outputs=A.RT_OUT_RGB8, accel=A.RT_ACCEL_BVH,
                  quantiser=A.RT_QUANT_PPM_LROUND, kernel_variant=0):
    """C4/C5 camera: straight down from (0,0,1), focal 24 mm, sensor 24 mm; light (-2,-1,1.5) white x5;
    jitter = jittered_samples(spp, 42) (antialias.h:12-27)."""
    cam = camera_init((0, 0, 1), (0, 0, 0), (0, 1, 0), 24.0, 24.0, width, height)
    return Frame(cam, width, height, mode=A.RT_MODE_HW2_BVH, accel=accel,
                 lights=[make_light((-2.0, -1.0, 1.5), (1, 1, 1), 5)], miss_color=(0.5, 0.7, 1.0), spp=spp,
                 jitter=jitter_table(spp, 42, True), max_depth=1, shadows=shadows, outputs=outputs,
                 quantiser=quantiser, kernel_variant=kernel_variant)


def load_mesh_npz(path):
    """Mesh fixture written by tools/make_golden.py: positions, normals, indices (+ obj ids)."""
    d = np.load(path)
    return d["positions"], d["normals"] if "normals" in d.files and d["normals"].size else None, d["indices"], \
        d["tri_obj_ids"] if "tri_obj_ids" in d.files else None


def hw1_frame(width=320, height=180, light_color=(1.0, 0.0, 1.0), accel=A.RT_ACCEL_BRUTE, outputs=A.RT_OUT_RGB_F32,
              quantiser=A.RT_QUANT_HW1_TRUNC):
    """HW1/src/render.cpp:43-58: camera (0,-1,1) -> (0,0.15,0), up z, 255 mm / 24 mm; light (-3,0,1)."""
    cam = camera_init((0.0, -1.0, 1.0), (0.0, 0.15, 0.0), (0.0, 0.0, 1.0), 255.0, 24.0, width, height)
    return Frame(cam, width, height, mode=A.RT_MODE_HW1, accel=accel,
                 lights=[make_light((-3.0, 0.0, 1.0), light_color, 1)], spp=1,
                 jitter=jitter_table(1, 42, False), outputs=outputs, quantiser=quantiser)


FROG_MATERIAL = dict(albedo=(0.8, 0.2, 0.2), kd=1.0, ks=0.5, specular_color=(0.04, 0.04, 0.04), shininess=32.0, kr=0.0)


def frog_frame(width=1920, height=1080, filling=False, outputs=A.RT_OUT_RGB_F32, shadows=True,
               quantiser=A.RT_QUANT_HW2_TRUNC, accel=A.RT_ACCEL_BVH):
    """GPUandCPU/assets/json_files/frog.json at depth 1 (camera (0,-.2,.2) -> (0,.1,0), 45 mm; light
    (-3,0,1) yellow x5, black miss colour); filling=True is the frame-filling variant of SURVEY §8d
    (focal 170 mm, look_at (0,.095,.03))."""
    if filling:
        cam = camera_init((0.0, -0.2, 0.2), (0.0, 0.095, 0.03), (0, 0, 1), 170.0, 24.0, width, height)
    else:
        cam = camera_init((0.0, -0.2, 0.2), (0.0, 0.1, 0.0), (0, 0, 1), 45.0, 24.0, width, height)
    return Frame(cam, width, height, mode=A.RT_MODE_HW2_BVH, accel=accel,
                 lights=[make_light((-3.0, 0.0, 1.0), (1.0, 1.0, 0.0), 5)], miss_color=(0, 0, 0), spp=1,
                 jitter=jitter_table(1, 42, True), max_depth=1, shadows=shadows, outputs=outputs, quantiser=quantiser)


def cornell_bounce_scene(mesh_npz):
    """Bounce-loop test scene: camera inside cornellbox.obj, every object partly mirror-like, one
    pure mirror and one absorber.  Returns (Scene, (pos, look_at, up, focal_mm, sensor_mm), lights, miss_color)."""
    d = np.load(mesh_npz)
    nobj = int(d["tri_obj_ids"].max()) + 1
    mats = [make_material(albedo=(0.7, 0.7, 0.7), kd=0.8, ks=0.1, kr=0.2, specular_color=(0.6, 0.6, 0.6)) for _ in range(nobj)]
    mats[1] = make_material(albedo=(0.8, 0.1, 0.1), kd=0.6, kr=0.4, specular_color=(0.9, 0.9, 0.9))
    mats[2] = make_material(albedo=(0.1, 0.8, 0.1), kd=0.0, kr=0.9, specular_color=(0.8, 0.8, 0.9))
    mats[3] = make_material(albedo=(0.1, 0.1, 0.1), kd=0.0, kr=0.0)
    sc = Scene(d["positions"], d["indices"], normals=d["normals"] if d["normals"].size else None,
               tri_obj_ids=d["tri_obj_ids"], materials=mats)
    cam = ((278.0, 273.0, -800.0), (278.0, 273.0, 0.0), (0.0, 1.0, 0.0), 35.0, 25.0)
    lights = [make_light((278.0, 500.0, 279.5), (1, 1, 1), 2), make_light((100.0, 300.0, 100.0), (1.0, 0.5, 0.2), 1)]
    return sc, cam, lights, (0.2, 0.3, 0.4)


def cornell_bounce_frame(cam_args, lights, miss, width, height, spp, max_depth, diffuse_bounce, outputs=A.RT_OUT_RGB_F32):
    cam = camera_init(cam_args[0], cam_args[1], cam_args[2], cam_args[3], cam_args[4], width, height)
    return Frame(cam, width, height, mode=A.RT_MODE_HW2_BVH, accel=A.RT_ACCEL_BVH, lights=lights, miss_color=miss, spp=spp,
                 jitter=jitter_table(spp, 42, True), max_depth=max_depth, shadows=True, outputs=outputs,
                 quantiser=A.RT_QUANT_HW2_TRUNC, diffuse_bounce=bool(diffuse_bounce))


def cpuonly_case(g, name, width=None, height=None, outputs=A.RT_OUT_RGB_F32, accel=A.RT_ACCEL_BVH):
    """Scene + frame of an RT_MODE_HW2_CPU fixture (tests/golden/cpuonly_scenes.npz, tools/make_golden_cpuonly.py):
    baked mesh, per-object materials, the CPUOnly camera (explicit sensor width), one point light with float intensity,
    one sample at the pixel centre (+0.5, CPUOnly/src/render.cpp:127-131), max_bounces mirror recursion."""
    nrm = g[name + "_normals"]
    mats = []
    for m in g[name + "_materials"]:
        mats.append(make_material(albedo=m[0:3], kd=m[3], specular_color=m[4:7], ks=m[7], shininess=m[8], kr=m[9], emission=m[10:13]))
    sc = Scene(g[name + "_positions"], g[name + "_indices"], normals=nrm if nrm.size else None, tri_obj_ids=g[name + "_tri_obj_ids"], materials=mats)
    c, li = g[name + "_camera"], g[name + "_light"]
    W, H, depth = (int(v) for v in g[name + "_frame"])
    W, H = width or W, height or H
    cam = camera_init_cpuonly(c[0:3], c[3:6], c[6:9], c[9], c[10], c[11], W, H)
    fr = Frame(cam, W, H, mode=A.RT_MODE_HW2_CPU, accel=accel, lights=[make_light_f(li[0:3], li[3:6], li[6])], spp=1,
               jitter=np.array([[0.5, 0.5]], np.float32), max_depth=depth, shadows=True, outputs=outputs, quantiser=A.RT_QUANT_CPU_TRUNC)
    return sc, fr


def cpuonly_area_case(g, outputs=A.RT_OUT_RGB_F32, rng_seed=0, accel=A.RT_ACCEL_BVH):
    """Scene + frame of the soft-shadow fixture (tests/golden/cpuonly_area.npz, tools/make_golden_area.py): the reference's
    config/sphere_area.json — one disk light (radius, shadow_samples), RT_MODE_HW2_CPU, one sample at the pixel centre."""
    mats = [make_material(albedo=m[0:3], kd=m[3], specular_color=m[4:7], ks=m[7], shininess=m[8], kr=m[9], emission=m[10:13]) for m in g["materials"]]
    nrm = g["normals"]
    sc = Scene(g["positions"], g["indices"], normals=nrm if nrm.size else None, tri_obj_ids=g["tri_obj_ids"], materials=mats)
    c, li = g["camera"], g["light"]
    W, H, depth = (int(v) for v in g["frame"][:3])
    cam = camera_init_cpuonly(c[0:3], c[3:6], c[6:9], c[9], c[10], c[11], W, H)
    fr = Frame(cam, W, H, mode=A.RT_MODE_HW2_CPU, accel=accel, lights=[make_light_f(li[0:3], li[3:6], li[6])], spp=1,
               jitter=np.array([[0.5, 0.5]], np.float32), max_depth=depth, shadows=True, outputs=outputs, quantiser=A.RT_QUANT_CPU_TRUNC,
               light_radius=[li[7]], light_shadow_samples=[int(li[8])], rng_seed=rng_seed)
    return sc, fr
