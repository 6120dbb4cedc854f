"""Host-side mirror of the reference's operator interface over the C ABI (include/rt_api.h).

The reference's seam is `render(numTriangles, W, H, cam, missColor, max_depth, spp, nodes, aabbs,
triangles, triObjectIds, objectMaterials, ..., lights, ..., output)` plus
`BVH::calculateAABBs / buildBVH` (HW2/HW2/GPUandCPU/include/query.h:13-29, bvh.h:412-433) and the
pixel loop inlined in HW1/src/render.cpp:60-124.  `Renderer.upload_scene` / `render` / `download`
wrap rt_upload_scene / rt_render / rt_download_image with the same argument meaning; errors raise
RtError carrying rt_last_error() (the reference throws std::runtime_error, imports.h:40-47).

There is no CPU path: importing works anywhere, but creating a Renderer needs the CUDA library
(librt_b200.so, built in-tree by build.py) and an sm_100 device, and fails loudly otherwise.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A

_LIB = None
# RT_B200_LIB: another build of the same library (kernel experiments are compiled side by side and A/B-timed in one GPU session)
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "librt_b200.so")


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rt error %d: %s" % (code, msg))
        self.code = code


def load_library():
    """Loads librt_b200.so (raises if it has not been built: there is no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -m raytracinginonesemester_b200.build` "
                               "(the CUDA extension is the product; there is no CPU fallback)" % LIB_PATH)
        lib = A.bind(C.CDLL(LIB_PATH))
        if lib.rt_api_version() != A.RT_API_VERSION:
            raise RuntimeError("%s has API version %d, this package expects %d: rebuild it (python -m raytracinginonesemester_b200.build)"
                               % (LIB_PATH, lib.rt_api_version(), A.RT_API_VERSION))
        _LIB = lib
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


def camera_init(pos, look_at, up, focal_length_mm, sensor_height_mm, width, height):
    """Camera::initialize (GPUandCPU/include/camera.h:72-94) -> rt_camera.  Raises on W/H < 1 like
    HW1's camera (HW1/include/camera.h:57-62)."""
    lib = load_library()
    cam = A.rt_camera()
    p, l, u = _f32(pos), _f32(look_at), _f32(up)
    rc = lib.rt_camera_init(C.byref(cam), _ptr(p, A.f32p), _ptr(l, A.f32p), _ptr(u, A.f32p),
                            float(focal_length_mm), float(sensor_height_mm), int(width), int(height))
    if rc != A.RT_OK:
        raise RtError(rc, "pixel_width and pixel_height must be >= 1")
    cam.params = dict(pos=tuple(float(x) for x in p), look_at=tuple(float(x) for x in l), up=tuple(float(x) for x in u),
                      focal_mm=float(focal_length_mm), sensor_mm=float(sensor_height_mm))   # for callers that re-create it
    return cam


def jitter_table(spp, seed=42, centered=True):
    """jittered_samples(spp, seed) of GPUandCPU/include/antialias.h:12-27 (centered) or HW1's."""
    lib = load_library()
    out = np.zeros((spp, 2), np.float32)
    rc = lib.rt_jitter_table(_ptr(out, A.f32p), int(spp), int(seed), 1 if centered else 0)
    if rc != A.RT_OK:
        raise RtError(rc, "rt_jitter_table")
    return out


def make_material(albedo=(0.8, 0.8, 0.8), kd=1.0, specular_color=(0.04, 0.04, 0.04), ks=0.0, shininess=32.0,
                  kr=0.0, emission=(0.0, 0.0, 0.0)):
    """Material with the defaults of GPUandCPU/include/material.h:6-20."""
    m = A.rt_material()
    m.albedo[:] = albedo
    m.kd = kd
    m.specular_color[:] = specular_color
    m.ks = ks
    m.shininess = shininess
    m.kr = kr
    m.emission[:] = emission
    return m


def make_light(position, color=(1.0, 1.0, 1.0), intensity=1):
    """Light of the HW1 / HW2-BVH renderers: integer intensity (GPUandCPU/include/scene.h:21-25)."""
    l = A.rt_light()
    l.position[:] = position
    l.color[:] = color
    l.intensity = int(intensity)
    return l


def make_light_f(position, color=(1.0, 1.0, 1.0), intensity=1.0):
    """Point light of the CPUOnly renderer (RT_MODE_HW2_CPU): float intensity (CPUOnly/include/raytracer.h:37-46)."""
    l = A.rt_light()
    l.position[:] = position
    l.color[:] = color
    l.intensity_f = float(intensity)
    return l


def camera_init_cpuonly(pos, look_at, up, focal_length_mm, sensor_height_mm, sensor_width_mm, width, height):
    """camera::initialize of the CPUOnly renderer (CPUOnly/include/camera.h:64-104) -> rt_camera."""
    lib = load_library()
    cam = A.rt_camera()
    p, l, u = _f32(pos), _f32(look_at), _f32(up)
    rc = lib.rt_camera_init_cpuonly(C.byref(cam), _ptr(p, A.f32p), _ptr(l, A.f32p), _ptr(u, A.f32p), float(focal_length_mm),
                                    float(sensor_height_mm), float(sensor_width_mm), int(width), int(height))
    if rc != A.RT_OK:
        raise RtError(rc, "pixel_width and pixel_height must be >= 1 and sensor_width_mm > 0")
    cam.params = dict(pos=tuple(float(x) for x in p), look_at=tuple(float(x) for x in l), up=tuple(float(x) for x in u),
                      focal_mm=float(focal_length_mm), sensor_mm=float(sensor_height_mm), sensor_w_mm=float(sensor_width_mm))
    return cam


def load_obj(path, next_object_id=0, position=None, rotation=None, scale=None):
    """LoadOBJ_ToMesh (+ applyObjectTransform when a transform is given) -> (positions, normals | None,
    indices, tri_obj_ids, next_object_id).  Pure host code in the C library (rt_mesh_*)."""
    lib = load_library()
    h = C.c_void_p()
    nid = C.c_int32(next_object_id)
    rc = lib.rt_mesh_load_obj(os.fsencode(path), C.byref(nid), C.byref(h))
    if rc != A.RT_OK:
        raise RtError(rc, (lib.rt_mesh_last_error() or b"").decode())
    try:
        if position is not None or rotation is not None or scale is not None:
            p, r, s = _f32(position or (0, 0, 0)), _f32(rotation or (0, 0, 0)), _f32(scale or (1, 1, 1))
            lib.rt_mesh_transform(h, _ptr(p, A.f32p), _ptr(r, A.f32p), _ptr(s, A.f32p))
        nv, nn, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.rt_mesh_counts(h, C.byref(nv), C.byref(nn), C.byref(nt))
        pos = np.zeros((nv.value, 3), np.float32)
        nrm = np.zeros((nn.value, 3), np.float32)
        idx = np.zeros((nt.value, 3), np.uint32)
        obj = np.zeros(nt.value, np.int32)
        lib.rt_mesh_copy(h, _ptr(pos, A.f32p), _ptr(nrm, A.f32p) if nn.value else A.f32p(), _ptr(idx, A.u32p), _ptr(obj, A.i32p))
    finally:
        lib.rt_mesh_free(h)
    return pos, (nrm if nn.value else None), idx, obj, int(nid.value)


class Scene:
    """Indexed triangle mesh + per-object materials, as the reference loaders produce them."""

    def __init__(self, positions, indices, normals=None, tri_obj_ids=None, materials=None, build_flags=0, transforms=None):
        # transforms: [(first_vertex, num_vertices, position, rotation_deg, scale), ...] baked on the device at upload
        self.transforms = list(transforms) if transforms else []
        self._xf_arr = (A.rt_object_transform * max(1, len(self.transforms)))()
        for k, (first, count, pos, rot, scl) in enumerate(self.transforms):
            t = self._xf_arr[k]
            t.first_vertex, t.num_vertices = int(first), int(count)
            t.position[:] = [float(v) for v in pos]
            t.rotation_deg[:] = [float(v) for v in rot]
            t.scale[:] = [float(v) for v in scl]
        self.positions = _f32(positions).reshape(-1, 3)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        self.normals = _f32(normals).reshape(-1, 3) if normals is not None and len(normals) else None
        self.tri_obj_ids = np.ascontiguousarray(tri_obj_ids, dtype=np.int32) if tri_obj_ids is not None else None
        self.materials = list(materials) if materials else []
        self.build_flags = build_flags
        self._mat_arr = (A.rt_material * max(1, len(self.materials)))(*self.materials)

    def c_struct(self):
        s = A.rt_scene()
        s.positions = _ptr(self.positions, A.f32p)
        s.normals = _ptr(self.normals, A.f32p)
        s.num_vertices = self.positions.shape[0]
        s.indices = _ptr(self.indices, A.u32p)
        s.num_triangles = self.indices.shape[0]
        s.tri_obj_ids = _ptr(self.tri_obj_ids, A.i32p)
        s.materials = C.cast(self._mat_arr, C.POINTER(A.rt_material)) if self.materials else C.POINTER(A.rt_material)()
        s.num_materials = len(self.materials)
        s.build_flags = self.build_flags
        s.transforms = C.cast(self._xf_arr, C.POINTER(A.rt_object_transform)) if self.transforms else C.POINTER(A.rt_object_transform)()
        s.num_transforms = len(self.transforms)
        return s


class DeviceMesh:
    """A mesh that exists only in device memory: OBJ text parsed on the GPU (rt_dmesh_parse_obj), same arrays as load_obj."""

    def __init__(self, renderer, handle=None):
        self.lib = renderer.lib
        self.h = handle if handle is not None else C.c_void_p()
        if handle is None:
            rc = self.lib.rt_dmesh_create(C.byref(self.h))
            if rc != A.RT_OK:
                raise RtError(rc, "rt_dmesh_create")

    @classmethod
    def parse_obj(cls, renderer, text, next_object_id=0):
        """text: the OBJ file's bytes.  Returns (DeviceMesh, next_object_id)."""
        data = bytes(text)
        h = C.c_void_p()
        nid = C.c_int32(next_object_id)
        rc = renderer.lib.rt_dmesh_parse_obj(renderer.ctx, data, len(data), C.byref(nid), C.byref(h))
        if rc != A.RT_OK:
            raise RtError(rc, (renderer.lib.rt_dmesh_last_error() or b"").decode())
        return cls(renderer, h), int(nid.value)

    def counts(self):
        nv, nn, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.rt_dmesh_counts(self.h, C.byref(nv), C.byref(nn), C.byref(nt))
        return int(nv.value), int(nn.value), int(nt.value)

    def stats(self):
        """(device ms of the whole parse incl. the H2D copy of the text, lines, numbers converted by the host's strtof)"""
        ms, ln, hard = C.c_float(), C.c_uint64(), C.c_uint64()
        self.lib.rt_dmesh_stats(self.h, C.byref(ms), C.byref(ln), C.byref(hard))
        return float(ms.value), int(ln.value), int(hard.value)

    def append(self, other):
        rc = self.lib.rt_dmesh_append(self.h, other.h)
        if rc != A.RT_OK:
            raise RtError(rc, (self.lib.rt_dmesh_last_error() or b"").decode())

    def download(self):
        """(positions, normals | None, indices, tri_obj_ids) copied to the host — for tests."""
        nv, nn, nt = self.counts()
        pos = np.zeros((nv, 3), np.float32)
        nrm = np.zeros((nn, 3), np.float32)
        idx = np.zeros((nt, 3), np.uint32)
        obj = np.zeros(nt, np.int32)
        rc = self.lib.rt_dmesh_copy(self.h, _ptr(pos, A.f32p), _ptr(nrm, A.f32p) if nn else A.f32p(), _ptr(idx, A.u32p), _ptr(obj, A.i32p))
        if rc != A.RT_OK:
            raise RtError(rc, (self.lib.rt_dmesh_last_error() or b"").decode())
        return pos, (nrm if nn else None), idx, obj

    def close(self):
        if self.h:
            self.lib.rt_dmesh_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceScene(Scene):
    """rt_scene over a DeviceMesh: the pointers are device pointers, which rt_upload_scene accepts as they are."""

    def __init__(self, dmesh, materials=None, build_flags=0, transforms=None):
        self.dmesh = dmesh
        nv, nn, nt = dmesh.counts()
        self._counts = (nv, nn, nt)
        super().__init__(np.zeros((1, 3), np.float32), np.zeros((1, 3), np.uint32), materials=materials, build_flags=build_flags, transforms=transforms)

    def c_struct(self):
        s = super().c_struct()
        p, n, i, o = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.dmesh.lib.rt_dmesh_arrays(self.dmesh.h, C.byref(p), C.byref(n), C.byref(i), C.byref(o))
        s.positions = C.cast(p, A.f32p)
        s.normals = C.cast(n, A.f32p) if n.value else A.f32p()
        s.indices = C.cast(i, A.u32p)
        s.tri_obj_ids = C.cast(o, A.i32p)
        s.num_vertices, s.num_triangles = self._counts[0], self._counts[2]
        return s


class Frame:
    """Per-frame arguments of render() (query.h:13-29): camera, lights, miss colour, spp, depth."""

    def __init__(self, cam, width, height, mode=A.RT_MODE_HW2_BVH, accel=A.RT_ACCEL_BVH, lights=(), miss_color=(0, 0, 0),
                 spp=1, jitter=None, max_depth=1, shadows=True, outputs=A.RT_OUT_RGB_F32, quantiser=A.RT_QUANT_PPM_LROUND,
                 kernel_variant=0, diffuse_bounce=False, light_radius=None, light_shadow_samples=None, rng_seed=0):
        self.diffuse_bounce = bool(diffuse_bounce)
        # RT_MODE_HW2_CPU soft shadows (CPUOnly/include/raytracer.h:37-46): per-light disk radius and shadow sample count
        self.light_radius = _f32(light_radius) if light_radius is not None else None
        self.light_shadow_samples = np.ascontiguousarray(light_shadow_samples, dtype=np.int32) if light_shadow_samples is not None else None
        self.rng_seed = int(rng_seed)
        self.cam, self.width, self.height, self.mode, self.accel = cam, int(width), int(height), mode, accel
        self.lights = list(lights)
        self._light_arr = (A.rt_light * max(1, len(self.lights)))(*self.lights)
        self.miss_color = tuple(miss_color)
        self.spp = int(spp)
        self.jitter = _f32(jitter).reshape(-1, 2) if jitter is not None else None
        self.max_depth, self.shadows, self.outputs, self.quantiser = int(max_depth), bool(shadows), int(outputs), int(quantiser)
        self.kernel_variant = int(kernel_variant)

    def c_struct(self):
        f = A.rt_frame()
        f.mode, f.accel, f.width, f.height = self.mode, self.accel, self.width, self.height
        f.cam = self.cam
        f.lights = C.cast(self._light_arr, C.POINTER(A.rt_light)) if self.lights else C.POINTER(A.rt_light)()
        f.num_lights = len(self.lights)
        f.miss_color[:] = self.miss_color
        f.spp = self.spp
        f.jitter = _ptr(self.jitter, A.f32p)
        f.max_depth, f.shadows, f.outputs, f.quantiser = self.max_depth, int(self.shadows), self.outputs, self.quantiser
        f.kernel_variant = self.kernel_variant
        f.diffuse_bounce = int(self.diffuse_bounce)
        f.light_radius = _ptr(self.light_radius, A.f32p)
        f.light_shadow_samples = _ptr(self.light_shadow_samples, A.i32p)
        f.rng_seed = self.rng_seed & 0xFFFFFFFF
        return f


class Renderer:
    """One context per GPU (one process per GPU; screen-space tiles are sharded across ranks)."""

    def __init__(self, device=0, rank=0, world=1, nccl_id=None, devices=None):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        if devices is not None:          # rt_create_multi: one process, one thread, several GPUs (the caller sees one renderer)
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.rt_create_multi(C.byref(self.ctx), arr, len(devices))
        else:
            rc = self.lib.rt_create(C.byref(self.ctx), int(device))
        if rc != A.RT_OK:
            raise RtError(rc, (self.lib.rt_last_error(None) or b"").decode())
        self.rank, self.world = rank, world
        if world > 1:
            assert nccl_id is not None and len(nccl_id) == 128
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            self._check(self.lib.rt_comm_init(self.ctx, rank, world, buf))
        self._frame = None
        self._scene = None

    @staticmethod
    def nccl_unique_id():
        lib = load_library()
        buf = (C.c_char * 128)()
        rc = lib.rt_comm_unique_id(buf)
        if rc != A.RT_OK:
            raise RtError(rc, (lib.rt_last_error(None) or b"").decode())
        return bytes(buf)

    def _check(self, rc):
        if rc != A.RT_OK:
            raise RtError(rc, (self.lib.rt_last_error(self.ctx) or b"").decode())

    def close(self):
        if self.ctx:
            self.lib.rt_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gather(self, mode):
        """Collective: how tiles reach rank 0 (A.RT_GATHER_AUTO / _NCCL / _PEER); returns the mode in effect."""
        self._check(self.lib.rt_comm_set_gather(self.ctx, int(mode)))
        return self.gather_mode()

    def set_sharding(self, chunks_per_rank):
        """Tile ownership bands per rank (0 = default); every rank must use the same value."""
        self._check(self.lib.rt_comm_set_sharding(self.ctx, int(chunks_per_rank)))

    def gather_mode(self):
        m = C.c_int()
        self._check(self.lib.rt_comm_gather_mode(self.ctx, C.byref(m)))
        return int(m.value)

    def upload_scene(self, scene):
        """calculateAABBs + buildBVH + triangle packing on the device; returns rt_build_info."""
        self._scene = scene
        if scene is None:
            self._check(self.lib.rt_upload_scene(self.ctx, None))
        else:
            s = scene.c_struct()
            self._check(self.lib.rt_upload_scene(self.ctx, C.byref(s)))
        info = A.rt_build_info()
        self._check(self.lib.rt_build_info_get(self.ctx, C.byref(info)))
        return info

    def render(self, frame):
        """Asynchronous launch of the fused ray-gen / traversal / shading kernel."""
        self._frame = frame
        f = frame.c_struct()
        self._check(self.lib.rt_render(self.ctx, C.byref(f)))

    def sync(self):
        ms = C.c_float()
        self._check(self.lib.rt_sync(self.ctx, C.byref(ms)))
        return ms.value

    def frame_times(self):
        """(whole rt_render, frame kernel only) device ms of the last frame on this rank."""
        a, b = C.c_float(), C.c_float()
        self._check(self.lib.rt_frame_times(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stream_handle(self):
        h = C.c_void_p()
        self._check(self.lib.rt_stream_handle(self.ctx, C.byref(h)))
        return int(h.value or 0)

    def host_image(self, nbytes):
        """Collective (rt_host_image_create): a host buffer shared by every rank's process; returns this rank's mapping as
        a uint8 numpy array.  Pass slices of it (the same on every rank) as `into` of render_into: every rank then copies
        its own bands to the host over its own PCIe link."""
        p = C.c_void_p()
        self._check(self.lib.rt_host_image_create(self.ctx, int(nbytes), C.byref(p)))
        self._shared = np.ctypeslib.as_array((C.c_uint8 * int(nbytes)).from_address(p.value))
        return self._shared

    def _image(self, fr, into):
        W, H = fr.width, fr.height
        img = A.rt_image()
        out = {}
        into = into or {}
        root = self.world == 1 or self.rank == 0 or bool(into)       # (ranks != 0 pass planes only for the shared host image)

        def buf(name, shape, dt):
            a = into.get(name)
            if a is None:
                a = np.empty(shape, dt)
            out[name] = a
            return a

        if root:
            if fr.outputs & A.RT_OUT_RGB_F32:
                img.rgb = _ptr(buf("rgb", (H, W, 3), np.float32), A.f32p)
            if fr.outputs & A.RT_OUT_RGB8:
                img.rgb8 = _ptr(buf("rgb8", (H, W, 3), np.uint8), A.u8p)
            if fr.outputs & A.RT_OUT_TRI_ID:
                img.tri_id = _ptr(buf("tri_id", (H, W), np.int32), A.i32p)
            if fr.outputs & A.RT_OUT_T:
                img.t = _ptr(buf("t", (H, W), np.float32), A.f32p)
        return img, out

    def download(self, into=None):
        """Blocking read-back of the planes requested in Frame.outputs -> dict of numpy arrays.
        `into` may map plane name -> preallocated (e.g. pinned) array."""
        img, out = self._image(self._frame, into)
        self._check(self.lib.rt_download_image(self.ctx, C.byref(img)))
        out["rays_primary"], out["rays_shadow"], out["gpu_ms"] = int(img.rays_primary), int(img.rays_shadow), float(img.gpu_ms)
        return out

    def render_into(self, frame, into=None):
        """rt_render_into: render + download in one blocking call (band-pipelined on a single GPU)."""
        self._frame = frame
        f = frame.c_struct()
        img, out = self._image(frame, into)
        self._check(self.lib.rt_render_into(self.ctx, C.byref(f), C.byref(img)))
        out["rays_primary"], out["rays_shadow"], out["gpu_ms"] = int(img.rays_primary), int(img.rays_shadow), float(img.gpu_ms)
        return out

    def frame_stats(self):
        """(node_visits, tri_tests, node_lines, tri_blocks) of the last frame rendered with a *_STATS variant."""
        v = [C.c_uint64() for _ in range(4)]
        self._check(self.lib.rt_frame_stats(self.ctx, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    def download_bvh(self):
        info = A.rt_build_info()
        self._check(self.lib.rt_build_info_get(self.ctx, C.byref(info)))
        nodes = np.zeros((int(info.num_nodes), 16), np.uint32)
        geom = np.zeros((int(info.num_triangles), 12), np.float32)
        ids = np.zeros(int(info.num_triangles), np.int32)
        self._check(self.lib.rt_debug_download_bvh(self.ctx, nodes.ctypes.data, geom.ctypes.data, _ptr(ids, A.i32p)))
        return nodes, geom, ids
