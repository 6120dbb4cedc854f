#!/usr/bin/env python
"""bench.py — Mrays/s of closest-hit queries (BVH + triangle) on BASELINE config C4.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3|c2|c5]

One "step" = one frame of the hot path: per-pixel ray generation, BVH traversal, Möller–Trumbore
closest hit, HW2 direct shading with one shadow ray per lit hit, 8-bit resolve (and, for N > 1, the
NCCL tile gather to rank 0).  Default workload (config.workload = "c4"): synthetic 1,000,000-triangle
terrain, 3840x2160, 1 spp with the reference jitter pair, primary + shadow rays.

value   = rays (primary + shadow) per second, scene/BVH resident in HBM, L2 flushed between steps,
          device time from CUDA events on the library's stream, max over ranks.
e2e     = the same metric through the public C ABI with host buffers: rt_render_into (frame description,
          lights and jitter copied host->device; render; 8-bit frame copied into pinned host memory,
          band-pipelined on one GPU), every step, wall clock.  This is what the reference's own "GPU Render Time"
          brackets (render + D2H copy, GPUandCPU/src/main.cu:370-378).
--impl reference: the reference's own CPU implementation of the path (oracle/_ref/libref_hw2.so,
          compiled in place from the reference sources; falls back to the oracle port) on all host
          cores, each step a bounded row sample of the same frame.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = ("c4", "c5", "c3", "c3fill", "c2", "c4small", "c3bounce", "c3fillbounce", "n1sphere")
FROG = os.path.join(ROOT, "tests", "golden", "frog_mesh.npz")


def make_workload(name, leaf_max=2, variant=0, shadows=True):
    """BASELINE.json configs: scene + frame + description.  c4 is the headline (metric is quoted on it);
    the others are parity/measurement cases selectable with --workload."""
    from raytracinginonesemester_b200 import _abi as A, api, scenes
    flags = A.RT_BUILD_LEAF_MAX(leaf_max)
    terr = {"c4": (1000, 500, 3840, 2160, 1), "c5": (2500, 2000, 7680, 4320, 16), "c4small": (200, 100, 960, 540, 1)}
    if name in terr:
        nx, ny, W, H, spp = terr[name]
        scene = lambda: scenes.terrain_scene(nx, ny, build_flags=flags)
        frame = scenes.terrain_frame(W, H, spp=spp, outputs=A.RT_OUT_RGB8, kernel_variant=variant, shadows=shadows)
        return dict(name=name, scene=scene, frame=frame, triangles=nx * ny * 2, mode="hw2",
                    desc="synthetic terrain(%d,%d,seed 42), HW2 shading, primary + shadow rays, BVH" % (nx, ny))
    if name == "n1sphere":     # CPUOnly/config/sphere.json as shipped (4 802 triangles, mirror spheres, max_bounces 4) at its own 360x240
        g = np.load(os.path.join(ROOT, "tests", "golden", "cpuonly_scenes.npz"))
        sc, fr = scenes.cpuonly_case(g, "sphere", width=360, height=240, outputs=A.RT_OUT_RGB8)
        fr.kernel_variant = variant
        return dict(name=name, scene=lambda: sc, frame=fr, triangles=int(sc.indices.shape[0]), mode="cpuonly", fixture=("sphere", g),
                    desc="HW2/CPUOnly config/sphere.json (mirror recursion depth 4, hard shadows), 360x240, RT_MODE_HW2_CPU over the BVH")
    d = np.load(FROG)

    def frog(extra=0):
        return api.Scene(d["positions"], d["indices"], normals=d["normals"], tri_obj_ids=d["tri_obj_ids"],
                         materials=[api.make_material(**scenes.FROG_MATERIAL)], build_flags=flags | extra)
    ntri = int(d["indices"].shape[0])
    if name == "c2":      # HW1 frog, 1080p, brute force: every ray tests every triangle (HW1/src/render.cpp:89-107)
        frame = scenes.hw1_frame(1920, 1080, accel=A.RT_ACCEL_BRUTE, outputs=A.RT_OUT_RGB8)
        return dict(name=name, scene=lambda: frog(A.RT_BUILD_NO_BVH), frame=frame, triangles=ntri, mode="hw1",
                    desc="HW1 frog.obj (19 858 triangles), 1920x1080, brute-force ray-triangle, HW1 shade()")
    if name in ("c3bounce", "c3fillbounce"):   # assets/json_files/frog.json AS SHIPPED: max_bounces 8, diffuse_bounce defaults to true (scene.h:18), 1920x1080
        frame = scenes.frog_frame(1920, 1080, filling=name == "c3fillbounce", outputs=A.RT_OUT_RGB8, shadows=shadows, quantiser=A.RT_QUANT_PPM_LROUND)
        frame.max_depth, frame.diffuse_bounce, frame.kernel_variant = 8, True, variant
        return dict(name=name, scene=frog, frame=frame, triangles=ntri, mode="hw2",
                    desc="HW2-BVH frog.json as shipped (max_bounces 8, hash-RNG diffuse bounces) at 1920x1080, %s" % ("frame-filling view" if name == "c3fillbounce" else "stock view"))
    filling = name == "c3fill"
    frame = scenes.frog_frame(3840, 2160, filling=filling, outputs=A.RT_OUT_RGB8, shadows=shadows, quantiser=A.RT_QUANT_PPM_LROUND)
    frame.kernel_variant = variant
    return dict(name=name, scene=frog, frame=frame, triangles=ntri, mode="hw2",
                desc="HW2-BVH frog.obj at 3840x2160, depth 1, %s" % ("frame-filling view (focal 170 mm)" if filling else "stock frog.json view (3.2 % of pixels hit)"))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread every 5 ms
    (nvidia_ml_py), falling back to a streaming nvidia-smi query when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.sm, self.max_mhz, self.reasons, self.proc, self.nvml = [], None, set(), None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nvml = None
            try:
                self.rows = []
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.t = threading.Thread(target=self._read, daemon=True)
                self.t.start()
            except Exception:
                self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml:
            self._stop.set()
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------- reference arm ----
def reference_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "libref_hw2.so")
    if os.path.exists(p):
        lib = C.CDLL(p)
        lib.ref_hw2_world.restype = C.c_void_p
        lib.ref_hw2_build.restype = C.c_double
        return lib
    return None


def cpuonly_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "libref_cpuonly.so")
    return C.CDLL(p) if os.path.exists(p) else None


def hw1_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "libref_hw1.so")
    if os.path.exists(p):
        lib = C.CDLL(p)
        lib.ref_hw1_render.restype = C.c_uint64
        return lib
    return None


class CpuReference:
    """The reference's CPU renderer of the path on host threads: kind 'reference' when the in-place build of the
    reference sources is present (oracle/_ref), else the oracle port.  HW2 workloads run render()'s per-pixel body
    (query.cu:136-165) over the reference LBVH; the HW1 workload runs the loop of HW1/src/render.cpp:72-116."""

    def __init__(self, wl):
        from raytracinginonesemester_b200 import _abi as A
        self.A, self.wl = A, wl
        self.scene, self.frame = wl["scene"](), wl["frame"]
        self.cores = os.cpu_count() or 1
        self.hw1 = wl["mode"] == "hw1"
        self.cpuonly = wl["mode"] == "cpuonly"
        self.lib = hw1_lib() if self.hw1 else cpuonly_lib() if self.cpuonly else reference_lib()
        self.kind = "reference" if self.lib else "port"
        scene = self.scene
        t0 = time.perf_counter()
        if self.hw1 or self.cpuonly:
            self.h = None
        elif self.lib:
            f32p, u32p, i32p = A.f32p, A.u32p, A.i32p
            nrm = scene.normals.ctypes.data_as(f32p) if scene.normals is not None else None
            self.h = self.lib.ref_hw2_world(scene.positions.ctypes.data_as(f32p), nrm, C.c_uint64(scene.positions.shape[0]),
                                            scene.indices.ctypes.data_as(u32p), C.c_uint64(scene.indices.shape[0]),
                                            scene.tri_obj_ids.ctypes.data_as(i32p))
            self.lib.ref_hw2_build(C.c_void_p(self.h))
        if not self.lib:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import orclib
            self.orclib = orclib
            self.h = None if (self.hw1 or self.cpuonly) else orclib.oracle_bvh(scene)
        self.build_s = time.perf_counter() - t0

    def census(self, row_begin, row_step):
        """(primary, shadow) rays of a row-strided pass, counted by the reference shim itself (ref_hw2_count_rays_rows:
        the reference's SearchBVH + the two conditions that gate IsInShadow); None when only the port is available."""
        if not self.lib or self.hw1 or self.cpuonly or self.frame.max_depth > 1:
            return None
        fr, A = self.frame, self.A
        cp = fr.cam.params
        f3 = lambda v: np.array(v, np.float32)
        cpos, look, up = f3(cp["pos"]), f3(cp["look_at"]), f3(cp["up"])
        larr = (A.rt_light * max(1, len(fr.lights)))(*fr.lights)
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.ref_hw2_count_rays_rows(C.c_void_p(self.h), cpos.ctypes.data_as(A.f32p), look.ctypes.data_as(A.f32p), up.ctypes.data_as(A.f32p),
                                         C.c_double(cp["focal_mm"]), C.c_double(cp["sensor_mm"]), fr.width, fr.height, fr.spp, larr, len(fr.lights),
                                         row_begin, row_step, self.cores, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def rays_in_rows(self, row_begin, row_step):
        """Times one row-strided pass; returns (primary rays, seconds)."""
        fr, A, scene = self.frame, self.A, self.scene
        W, H = fr.width, fr.height
        rows = len(range(row_begin, H, row_step))
        cp = fr.cam.params
        f3 = lambda v: np.array(v, np.float32)
        cpos, look, up = f3(cp["pos"]), f3(cp["look_at"]), f3(cp["up"])
        t0 = time.perf_counter()
        if self.lib and self.cpuonly:
            # the CPUOnly renderer's pixel loop (brute force over all triangles, as the reference runs it), rows interleaved over host threads
            name, g = self.wl["fixture"]
            c, li = g[name + "_camera"], g[name + "_light"]

            class L(C.Structure):
                _fields_ = [("position", C.c_float * 3), ("color", C.c_float * 3), ("intensity", C.c_float)]
            l = L(); l.position[:] = [float(v) for v in li[0:3]]; l.color[:] = [float(v) for v in li[3:6]]; l.intensity = float(li[6])
            marr = (A.rt_material * len(scene.materials))(*scene.materials)
            rgb = np.zeros((H, W, 3), np.float32)
            cposf, lookf, upf = f3(c[0:3]), f3(c[3:6]), f3(c[6:9])
            rows_all = list(range(row_begin, H, row_step))

            def work(t):
                for y in rows_all[t::self.cores]:
                    self.lib.ref_cpu_render_rows(scene.positions.ctypes.data_as(A.f32p), scene.normals.ctypes.data_as(A.f32p) if scene.normals is not None else None,
                                                 C.c_uint64(scene.positions.shape[0]), scene.indices.ctypes.data_as(A.u32p), C.c_uint64(scene.indices.shape[0]),
                                                 scene.tri_obj_ids.ctypes.data_as(A.i32p), marr, len(scene.materials), cposf.ctypes.data_as(A.f32p), lookf.ctypes.data_as(A.f32p),
                                                 upf.ctypes.data_as(A.f32p), C.c_double(float(c[9])), C.c_double(float(c[10])), C.c_double(float(c[11])), W, H, C.byref(l), 1,
                                                 int(fr.max_depth), y, H, rgb.ctypes.data_as(A.f32p), None, None)
            th = [threading.Thread(target=work, args=(t,)) for t in range(self.cores)]
            [t.start() for t in th]; [t.join() for t in th]
        elif self.lib and self.hw1:
            l = fr.lights[0]
            lp, lc = f3(list(l.position)), f3(list(l.color))
            rgb8 = np.zeros((H, W, 3), np.uint8)
            nrm = scene.normals if scene.normals is not None else np.zeros_like(scene.positions)
            self.lib.ref_hw1_render(scene.positions.ctypes.data_as(A.f32p), nrm.ctypes.data_as(A.f32p), scene.indices.ctypes.data_as(A.u32p),
                                    C.c_uint64(scene.indices.shape[0]), cpos.ctypes.data_as(A.f32p), look.ctypes.data_as(A.f32p),
                                    up.ctypes.data_as(A.f32p), C.c_double(cp["focal_mm"]), C.c_double(cp["sensor_mm"]), W, H,
                                    lp.ctypes.data_as(A.f32p), lc.ctypes.data_as(A.f32p), fr.spp, C.c_uint(42), row_begin, row_step, self.cores,
                                    None, rgb8.ctypes.data_as(A.u8p), None, None)
        elif self.lib:
            ms = f3(fr.miss_color)
            rgb = np.zeros((H, W, 3), np.float32)
            marr = (A.rt_material * len(scene.materials))(*scene.materials)
            larr = (A.rt_light * len(fr.lights))(*fr.lights)
            self.lib.ref_hw2_render_rows(C.c_void_p(self.h), cpos.ctypes.data_as(A.f32p), look.ctypes.data_as(A.f32p), up.ctypes.data_as(A.f32p),
                                         C.c_double(cp["focal_mm"]), C.c_double(cp["sensor_mm"]), W, H, ms.ctypes.data_as(A.f32p), int(fr.max_depth), fr.spp,
                                         marr, len(scene.materials), larr, len(fr.lights), int(fr.diffuse_bounce), row_begin, row_step, self.cores,
                                         rgb.ctypes.data_as(A.f32p), None, None)
        else:
            self.orclib.oracle_render(scene, fr, bvh=self.h, threads=self.cores, row_begin=row_begin, row_step=row_step, want=("rgb",))
        dt = time.perf_counter() - t0
        return rows * W * fr.spp, dt


def reference_cuda_leg(wl, rays, flush_bytes=256 << 20):
    """The reference's OWN CUDA renderer (renderBatchCUDA + normalizeCUDA, GPUandCPU/include/query.cu:12-128, and its
    Thrust LBVH build) compiled in place for sm_100 with the reference's flags (oracle/_ref/libref_hw2_cuda.so) and timed
    on this GPU on the same scene and frame: the GPU bar of SURVEY §8d.  Measurement only — never on the product path.
    Its GPU build jitters with a per-pixel wang hash (query.cu:36-45) instead of the mt19937 table, and uses
    --use_fast_math, so its image is close to but not bit-comparable with the CPU reference."""
    from raytracinginonesemester_b200 import _abi as A
    p = os.path.join(ROOT, "oracle", "_ref", "libref_hw2_cuda.so")
    if wl["mode"] != "hw2" or not os.path.exists(p):
        return None
    lib = C.CDLL(p)
    lib.ref_cuda_world_create.restype = C.c_void_p
    lib.ref_cuda_build_ms.restype = C.c_double
    scene, fr = wl["scene"](), wl["frame"]
    W, H, spp = fr.width, fr.height, fr.spp
    f3 = lambda v: np.array(v, np.float32)
    marr = (A.rt_material * len(scene.materials))(*scene.materials)
    nrm = scene.normals.ctypes.data_as(A.f32p) if scene.normals is not None else None
    h = lib.ref_cuda_world_create(scene.positions.ctypes.data_as(A.f32p), nrm, C.c_uint64(scene.positions.shape[0]),
                                  scene.indices.ctypes.data_as(A.u32p), C.c_uint64(scene.indices.shape[0]),
                                  scene.tri_obj_ids.ctypes.data_as(A.i32p), marr, len(scene.materials))
    if not h:
        return {"unavailable": "ref_cuda_world_create failed"}
    cp = fr.cam.params
    cpos, look, up, ms = f3(cp["pos"]), f3(cp["look_at"]), f3(cp["up"]), f3(fr.miss_color)
    larr = (A.rt_light * len(fr.lights))(*fr.lights)
    reps = 10 if W * H * spp <= 3840 * 2160 else 2
    dev, e2e = np.zeros(reps, np.float32), np.zeros(reps, np.float32)
    rc = lib.ref_cuda_render(C.c_void_p(h), cpos.ctypes.data_as(A.f32p), look.ctypes.data_as(A.f32p), up.ctypes.data_as(A.f32p),
                             C.c_double(cp["focal_mm"]), C.c_double(cp["sensor_mm"]), W, H, ms.ctypes.data_as(A.f32p), 1, spp,
                             larr, len(fr.lights), 1, 2, reps, C.c_uint64(flush_bytes), dev.ctypes.data_as(A.f32p), e2e.ctypes.data_as(A.f32p), None)
    build_ms = float(lib.ref_cuda_build_ms(C.c_void_p(h)))
    lib.ref_cuda_world_free(C.c_void_p(h))
    if rc != 0:
        return {"unavailable": "ref_cuda_render failed"}
    return {"impl": "reference CUDA renderer (renderBatchCUDA, one thread per pixel, 512-entry local stack, fp64 slabs) built in place for sm_100 with --use_fast_math",
            "value": rays / (float(dev.mean()) * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": float(dev.mean()), "frames": reps,
            "e2e_ms_per_step": float(e2e.mean()), "e2e_value": rays / (float(e2e.mean()) * 1e-3) / 1e6,
            "e2e_note": "render() + D2H of the float image = the reference's own 'GPU Render Time' (main.cu:370-378)",
            "lbvh_build_ms": build_ms, "l2": "flushed before each frame (256 MiB memset)",
            "rays": "the product's device count for the same frame (one shadow ray per lit hit - the same rule)"}


METRIC = "Mrays/s closest-hit (BVH+tri)"


def cpu_sample(ref, seconds, max_passes=64):
    """Bounded CPU sample of the frame: row-strided passes sized from a thin calibration pass."""
    fr = ref.frame
    W, H, spp = fr.width, fr.height, fr.spp
    cal_rays, cal_dt = ref.rays_in_rows(5, max(1, H // 8))
    rows = max(1, int(cal_rays / cal_dt * seconds / (W * spp)))
    stp = max(1, H // rows)
    return stp


def reference_record(workload, steps, warmup, seconds_per_step=6.0, n_gpus=1):
    """The reference's CPU path on `workload`: K steps, each a bounded row sample of the frame (~seconds_per_step of CPU work on
    all host threads).  Rays = primary + shadow, both counted by the reference shim on the sampled rows."""
    wl = make_workload(workload)
    ref = CpuReference(wl)
    fr = wl["frame"]
    W, H, spp = fr.width, fr.height, fr.spp
    step = cpu_sample(ref, seconds_per_step)
    for _ in range(min(warmup, 1)):
        ref.rays_in_rows(1 % step, step)
    tot_rays, tot_s, prim, shad = 0, 0.0, 0, 0
    for k in range(steps):
        rays, dt = ref.rays_in_rows(k % step, step)
        tot_rays += rays
        tot_s += dt
    cen = ref.census(0, step)                      # untimed: shadow rays per primary ray on a sampled row set
    if cen is not None:
        prim, shad = cen
        ratio, how = shad / max(prim, 1), "counted by the reference shim on the sampled rows"
    else:
        ratio, how = 0.0, "primary rays only (no counters in this build)"
    mrays = tot_rays * (1.0 + ratio) / tot_s / 1e6
    return {
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s", "n_gpus": n_gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * tot_s / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "description": wl["desc"], "triangles": wl["triangles"], "width": W, "height": H, "spp": spp,
                   "rays": "primary+shadow" if ratio else "primary"},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": ref.cores, "kind": ref.kind,
                         "sample": "every %d-th row of the %dx%d frame per step (%d rows), full-frame rate extrapolated; shadow rays = primary x %.4f, %s" % (step, W, H, len(range(0, H, step)), ratio, how),
                         "lbvh_build_s": ref.build_s},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference(args, rank, world):
    if rank != 0:
        return
    line = reference_record(args.workload, args.steps, args.warmup, n_gpus=args.gpus)
    if args.gpus >= 8 and args.workload == "c4" and not args.no_c5:
        # BASELINE config 5 (10M triangles, 8K, 16 spp) beside the headline, like the product arm: a thin sample
        try:
            line["configs"] = [reference_record("c5", max(1, min(args.steps, 3)), 0, seconds_per_step=4.0, n_gpus=args.gpus)]
        except Exception as e:
            line["configs"] = [{"config": {"workload": "c5"}, "unavailable": "%s: %s" % (type(e).__name__, e)}]
    print(json.dumps(line))


# ------------------------------------------------------------------------------ our arm ----
def bench_workload(args, r, dist, rank, world, local_rank, workload, steps, with_cpu=True):
    """Benchmarks one workload on the already-created renderer(s); returns the record (rank 0) or None."""
    import torch
    from raytracinginonesemester_b200 import _abi as A, api, parallel
    wl = make_workload(workload, args.leaf_max, args.variant, not args.no_shadows)
    frame = wl["frame"]
    W, H, spp = frame.width, frame.height, frame.spp
    brute = frame.accel == A.RT_ACCEL_BRUTE
    scene = wl["scene"]() if rank == 0 else None
    first = r.upload_scene(scene)          # first call in the process: pays CUDA module loading and the pool's first allocations
    r.upload_scene(scene)
    t0 = time.perf_counter()
    info = r.upload_scene(scene)           # steady state (what another scene or a re-upload costs)
    upload_wall = time.perf_counter() - t0
    gather = {A.RT_GATHER_NCCL: "nccl send/recv + unpack", A.RT_GATHER_PEER: "peer stores into rank 0's image over NVLink; band flags and the ready/done handshake inside the frame kernel"}[r.gather_mode()] if world > 1 else "none"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    pinned = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True).numpy() if rank == 0 else None
    # N > 1: the e2e frame goes to a host buffer shared by all ranks' processes (rt_host_image_create): every rank copies the bands
    # it rendered over its own PCIe link; rank 0 reads the frame from its mapping
    shared = r.host_image(3 * W * H)[:3 * W * H].reshape(H, W, 3) if world > 1 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # All device work of a step is enqueued on the library's own stream (rt_stream_handle): the L2 flush, an event,
    # rt_render (frame kernel + tile delivery to rank 0), an event.  The K timed steps are enqueued back to back
    # and bracketed by barrier + synchronize on both sides; a step's time is the CUDA-event interval around its
    # rt_render (the flush is outside it), max over ranks.
    ext = torch.cuda.ExternalStream(r.stream_handle())

    def enqueue_step(do_flush=True):
        with torch.cuda.stream(ext):
            if do_flush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r.render(frame)
            e1.record()
        return e0, e1

    def timed(n, do_flush=True):
        barrier()
        ev = [enqueue_step(do_flush) for _ in range(n)]
        barrier()
        return [a.elapsed_time(b) for a, b in ev]

    timed(args.warmup if args.profile else max(args.warmup, 3))
    # traversal statistics (untimed).  (1) the frame kernel's own counters (STATS instance): wide entries / triangle blocks it
    # requests; (2) SURVEY §8(d)'s N_node / N_tri: a SINGLE-RAY walk of the same BVH, counted by the per-ray kernel, whose
    # counters equal the host walk of the downloaded BVH (tests/test_gpu_parity.py) — the gap between the two is the
    # redundancy a packet pays (every lane tests every triangle any lane reaches).
    single = None
    if not brute:
        frame.kernel_variant = A.RT_VARIANT_PER_RAY_STATS
        r.render(frame)
        r.download(into={"rgb8": pinned} if rank == 0 else None)
        single = r.frame_stats()
        frame.kernel_variant = (A.RT_VARIANT_PER_RAY_STATS if args.variant >= 10 else
                                A.RT_VARIANT_STATS if args.variant in (A.RT_VARIANT_DEFAULT, A.RT_VARIANT_PERSIST) else
                                A.RT_VARIANT_FRUSTUM_STATS if args.variant == A.RT_VARIANT_FRUSTUM else A.RT_VARIANT_PACKET_STATS)
    r.render(frame)
    st = r.download(into={"rgb8": pinned} if rank == 0 else None)
    nv, nt, nl, nb = r.frame_stats()
    if brute:      # every ray tests every triangle; a block of 128 rays streams the 48-byte blocks once through shared memory
        ntri = wl["triangles"]
        nv, nl = 0, 0
        nt = st["rays_primary"] * ntri
        nb = (st["rays_primary"] // 128) * ntri
    frame.kernel_variant = args.variant
    cnt = torch.tensor([st["rays_primary"], st["rays_shadow"], nv, nt, nl, nb] + list(single[:2] if single else (0, 0)), dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(cnt)
    rays_primary, rays_shadow, nv_all, nt_all, nl_all, nb_all, n_node_1, n_tri_1 = [float(x) for x in cnt.cpu()]
    rays = rays_primary + rays_shadow

    # rt_render (the device-timed steps): RT_VARIANT_DEFAULT is the block-per-tile launch; the persistent kernel when asked for by name
    persistent = (not brute) and args.variant in (A.RT_VARIANT_PERSIST, A.RT_VARIANT_PERSIST_EXACT_MT, A.RT_VARIANT_PERSIST_OCC8, A.RT_VARIANT_PERSIST_OCC10)
    if world == 1:
        launches_per_step = 1
    elif r.gather_mode() == A.RT_GATHER_PEER:
        launches_per_step = 1 if persistent else 3      # per rank: the frame kernel; other kernels need a flag kernel either side
    else:
        launches_per_step = 1 + world                   # rank 0: frame kernel + one unpack per rank
    sampler = ClockSampler(local_rank) if rank == 0 else None
    step_ms = timed(steps)
    warm_ms = timed(min(steps, 5), do_flush=False)
    # frame kernel alone vs the whole rt_render, per rank (host-synchronised frames, untimed for the headline)
    ktimes = []
    for _ in range(3):
        flush.zero_()
        barrier()
        r.render(frame)
        ktimes.append(r.frame_times())
    kern_ms = float(np.mean([k for _, k in ktimes]))
    clocks = sampler.stop() if sampler else None      # clocks were sampled during the device-timed steps
    # end to end through the C ABI with host buffers (render + copy into pinned host memory), wall clock per rank; the
    # ranks are aligned before every step (outside the timed region) and the step's time is the max over ranks
    e2e_s, start_late = [], []

    def aligned_start():
        """Ranks leave a collective tens of microseconds apart, and on a 0.4 ms frame that skew would be counted twice (the other
        ranks wait for rank 0 to enter the frame, rank 0 waits for the last rank's copies).  So the ranks agree on a start TIME:
        rank 0 proposes one 400 us ahead on the host's monotonic clock (one clock for all processes of the node) and everybody
        spins until it.  Returns how late this rank was (0 when it made it)."""
        barrier()
        if dist is None:
            return 0.0
        tt = torch.zeros(1, dtype=torch.float64, device="cuda")
        if rank == 0:
            tt[0] = time.perf_counter() + 400e-6
        dist.broadcast(tt, 0)
        target = float(tt.item())
        late = time.perf_counter() - target
        while time.perf_counter() < target:
            pass
        return max(late, 0.0)

    for i in range(-max(args.warmup, 3), steps):      # the warm-up calls pay rt_render_into's one-time stream / event / band set-up
        late = aligned_start()
        t0 = time.perf_counter()
        e2e_out = r.render_into(frame, into={"rgb8": shared} if world > 1 else {"rgb8": pinned})     # rt_render_into: the user-facing "frame to host memory" call
        if i >= 0:
            e2e_s.append(time.perf_counter() - t0)
            start_late.append(late)
    e2e_kernel_ms = r.frame_times()[1]                 # the frame kernel inside the last end-to-end step (this rank)
    # PARITY OF WHAT WAS TIMED (world > 1): rank 0 renders the same frame alone and compares it with the gathered planes —
    # the 8-bit frame the e2e loop just delivered, and one extra untimed frame with ids and t
    gather_parity = None
    if world > 1:
        allp = A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T
        keep = frame.outputs
        frame.outputs = allp
        r.render(frame)
        got = r.download()
        frame.outputs = keep
        if rank == 0:
            solo = api.Renderer(local_rank)
            solo.upload_scene(scene)
            frame.outputs = allp
            solo.render(frame)
            ref = solo.download()
            frame.outputs = keep
            solo.close()
            gather_parity = {"rgb8_equal": bool(np.array_equal(got["rgb8"], ref["rgb8"])), "tri_id_equal": bool(np.array_equal(got["tri_id"], ref["tri_id"])),
                             "t_equal": bool(np.array_equal(got["t"], ref["t"])), "e2e_rgb8_equal": bool(np.array_equal(shared, ref["rgb8"])),
                             "pixels": int(W * H), "against": "the same frame rendered by rank 0 alone (single-GPU context, same scene)"}
        barrier()
    per_rank = torch.zeros(world, dtype=torch.float64, device="cuda")
    per_rank[rank] = kern_ms
    if dist is not None:
        dist.all_reduce(per_rank)
    t = torch.tensor([step_ms, e2e_s + [0.0] * (len(step_ms) - len(e2e_s)), start_late + [0.0] * (len(step_ms) - len(start_late))], dtype=torch.float64, device="cuda")
    tw = torch.tensor(warm_ms + [kern_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    step_ms = t[0].cpu().numpy()
    e2e_s = t[1].cpu().numpy()
    start_late_us = 1e6 * float(t[2].max())            # worst lateness of any rank at any timed step's agreed start (0 = all on time)
    ms = float(step_ms.mean())
    step_spread = {"median_ms": float(np.median(step_ms)), "min_ms": float(step_ms.min()), "max_ms": float(step_ms.max())}
    value = rays / (ms * 1e-3) / 1e6
    e2e_value = rays / float(e2e_s.mean()) / 1e6
    e2e_spread = {"median_ms": 1e3 * float(np.median(e2e_s)), "min_ms": 1e3 * float(e2e_s.min()), "max_ms": 1e3 * float(e2e_s.max())}
    if rank != 0:
        return None

    peak, peak_kind = measured_peaks()
    peak = peak * world
    # ncu capture of the dominant kernel on this workload (profiles/traffic.json, re-captured whenever the kernel changes; the
    # record names the kernel and the profile it came from): DRAM bytes and warp instructions per launch
    traffic, traffic_src, winst, tkern = None, None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        if tj and args.variant == 0:
            # The instruction count of a frame does not depend on which rank renders which tile (same packets, same kernel code:
            # the persistent launch of the multi-GPU path executed 1.704e9 against 1.700e9), so the single-GPU capture also
            # gives the N-rank issue roofline; the DRAM bytes are per launch of ONE GPU and are only reported at N = 1.
            traffic_src, tkern = tj["source"], tj.get("kernel")
            traffic = float(tj["dram_bytes_per_launch"]) if world == 1 else None
            winst = float(tj.get("warp_instructions_per_launch", 0)) or None
    except Exception:
        pass
    # algorithmic bytes, SURVEY §8(d): B_ray = 64 N_node + 48 N_tri + 16 with N_node / N_tri of a single-ray walk of the uploaded BVH
    b_ray_8d = (64.0 * n_node_1 + 48.0 * n_tri_1) / max(rays, 1.0) + 16.0 if not brute else 48.0 * nt_all / max(rays, 1.0) + 16.0
    hbm_achieved = rays * b_ray_8d / (ms * 1e-3) / 1e9
    # bytes the frame kernel itself requests at L1 (frustum traversal: one 32-byte wide entry per lane-box test, counted in 64-byte
    # units; one 48-byte triangle block per warp test), one shading block per hit pixel, the 8-bit frame
    bytes_requested = 64.0 * nl_all + 48.0 * nb_all + 48.0 * rays_primary + 3.0 * W * H
    compulsory = float(info.arena_bytes) + 3.0 * W * H        # every byte of the arena once + the frame
    hbm = {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
           "peak_kind": peak_kind + (" x %d GPUs" % world if world > 1 else ""),
           "bytes_per_ray": b_ray_8d, "n_node_single_ray": n_node_1 / max(rays, 1.0), "n_tri_single_ray": n_tri_1 / max(rays, 1.0),
           "counted_by": "per-ray kernel walking the uploaded BVH one ray at a time (== host walk of the downloaded BVH, tested)",
           "kernel_requested_bytes_per_launch_at_l1": bytes_requested, "kernel_tri_tests_per_ray": nt_all / max(rays, 1.0),
           "packet_redundancy_tri_tests": (nt_all / n_tri_1) if n_tri_1 else None,
           "dram_bytes_per_launch_ncu": traffic, "dram_frac_of_peak": (traffic / (ms * 1e-3) / 1e9 / peak) if traffic else None,
           "compulsory_bytes_per_launch": compulsory,
           "note": "the working set is cache-resident (ncu: L1 hit 78 %, L2 hit 84-92 %, DRAM < 1 % of peak): this kernel is NOT HBM-bound; "
                   "the figure above 1.0 x peak that SURVEY 8(d) anticipates would mean 'served from cache'"}
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "description": wl["desc"], "triangles": wl["triangles"], "width": W, "height": H, "spp": spp,
                   "rays": "primary+shadow" if rays_shadow else "primary", "rays_primary": rays_primary, "rays_shadow": rays_shadow,
                   "l2": "flushed between timed steps (256 MiB memset); warm-L2 figure in value_warm_l2",
                   "tile_sharding": "16x8-px tiles in %d contiguous bands per rank, band c -> rank c %% %d" % (args.chunks or parallel.DEFAULT_CHUNKS_PER_RANK, world), "gather": gather, "leaf_max": args.leaf_max, "variant": args.variant},
        "ms_per_step_spread": step_spread,
        "primary_mrays_s": rays_primary / (ms * 1e-3) / 1e6,
        "value_warm_l2": rays / (float(tw[:-1].mean()) * 1e-3) / 1e6,
        "frame_kernel_ms_max_rank": float(tw[-1]), "gather_overhead_ms": ms - float(tw[-1]),
        "frame_kernel_ms_per_rank": [round(float(x), 4) for x in per_rank.cpu()],
        "bvh_nodes": int(info.num_nodes), "bvh_build_ms": float(info.build_ms), "bvh_build_first_call_ms": float(first.build_ms),
        "scene_upload_ms": float(info.upload_ms),
        "scene_upload_wall_s": upload_wall,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": 1e3 * float(e2e_s.mean()),
                "h2d_bytes_per_step": int(C.sizeof(A.rt_frame) + 28 * len(frame.lights) + 8 * spp),
                "d2h_bytes_per_step": int(3 * W * H + 32), "spread": e2e_spread, "frame_kernel_ms_rank0": float(e2e_kernel_ms), "start_late_us_max": start_late_us,
                "delivery": ("every rank copies its own bands into one shared page-locked host buffer (rt_host_image_create), %d PCIe links" % world) if world > 1 else "band-pipelined copy into pinned host memory"},
        "gpu_launches": int(steps * launches_per_step),
        "clocks": clocks,
    }
    if gather_parity is not None:
        line["gather_parity"] = gather_parity
    if brute:   # 51 flop per ray-triangle test in the reference's unfused formulation (27 mul, 23 add/sub, 1 div)
        tests_s = nt_all / (ms * 1e-3)
        pk = 148 * 128 * 2 * 1.965e9 / 1e12 * world
        line["roofline"] = {"bound": "fp32-issue", "achieved": tests_s * 51 / 1e12, "peak": pk, "unit": "TFLOP/s", "frac": tests_s * 51 / 1e12 / pk, "traffic": traffic,
                            "ray_triangle_tests_per_s": tests_s, "flop_per_test": 51,
                            "note": "peak counts FMA as 2 flop; the exactly-rounded test cannot use FMA, so 0.5 is its ceiling; triangles are streamed through shared memory once per 128 rays",
                            "hbm": hbm}
    else:
        # The bound the profile shows: warp-instruction issue (4 schedulers per SM, one warp instruction per cycle each).
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        pk = 148 * 4 * mhz * 1e6 / 1e9 * world
        if winst:
            line["roofline"] = {"bound": "issue", "achieved": winst / (ms * 1e-3) / 1e9, "peak": pk, "unit": "G warp-instr/s", "frac": winst / (ms * 1e-3) / 1e9 / pk,
                                "traffic": traffic, "warp_instructions_per_launch": winst, "source": traffic_src, "kernel": tkern,
                                "peak_kind": "148 SMs x 4 schedulers x SM clock sampled during the run",
                                "note": "issue-slot roofline: smsp__inst_executed.sum of the ncu capture named in `source` / this run's kernel time; traffic = dram bytes of the same capture" +
                                        ("" if world == 1 else "; N ranks: the whole frame's instruction count (single-GPU capture; it does not depend on the sharding) over the max-over-ranks time, peak = N GPUs; no per-rank DRAM capture (never ncu a multi-rank run)"),
                                "hbm": hbm}
        else:   # no capture for this workload / variant / rank count: only the byte roofline can be formed from this run
            line["roofline"] = dict(hbm, traffic=traffic, issue_peak_g_warp_instr_s=pk,
                                    note="no ncu capture on file for this workload/variant/rank count (profiles/traffic.json): HBM byte roofline of SURVEY 8(d) only; the kernel is issue-bound (see the c4 record) - " + hbm["note"])
    if world == 1 and with_cpu and not args.no_cpu_baseline and not args.profile:
        ref = CpuReference(make_workload(workload))
        stp = cpu_sample(ref, 12.0)
        n, dt, passes = 0, 0.0, 0
        while dt < 10.0 and passes < 64:                  # bounded sample: ~10 s of CPU work
            a, b = ref.rays_in_rows(passes % stp, stp)
            n += a; dt += b; passes += 1
        cen = ref.census(0, stp)
        ratio = (cen[1] / max(cen[0], 1)) if cen else rays_shadow / max(rays_primary, 1.0)
        line["cpu_baseline"] = {"value": n * (1 + ratio) / dt / 1e6, "unit": "Mrays/s", "cores": ref.cores, "kind": ref.kind,
                                "sample": "%d pass(es) over every %d-th row of the %dx%d frame, %.1f s of CPU work; shadow rays = primary x %.4f (%s)" %
                                          (passes, stp, W, H, dt, ratio, "counted by the reference shim on the sampled rows" if cen else "device count")}
    line["_rays"] = rays
    line["_e2e_value"] = e2e_value
    return line


def run_ours(args, rank, world, local_rank):
    import torch
    from raytracinginonesemester_b200 import _abi as A
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    from raytracinginonesemester_b200 import parallel
    torch.cuda.set_device(local_rank)
    dist, rank, world, local_rank = parallel.init_process_group("nccl")
    r = parallel.make_renderer(dist, rank, world, local_rank)
    if world > 1:
        r.set_sharding(args.chunks)
        r.set_gather({"auto": A.RT_GATHER_AUTO, "nccl": A.RT_GATHER_NCCL, "peer": A.RT_GATHER_PEER}[args.gather])
    line = bench_workload(args, r, dist, rank, world, local_rank, args.workload, args.steps)
    extra = None
    if world >= 8 and args.workload == "c4" and not args.no_c5 and not args.profile:
        # BASELINE config 5 (10M triangles, 7680x4320, 16 spp, 8 GPUs) as a driver-timed record beside the headline
        try:
            extra = bench_workload(args, r, dist, rank, world, local_rank, "c5", max(3, min(args.steps, 10)), with_cpu=False)
        except Exception as e:
            extra = {"config": {"workload": "c5"}, "unavailable": "%s: %s" % (type(e).__name__, e)}
    r.close()
    if rank == 0:
        rays, e2e_value = line.pop("_rays"), line.pop("_e2e_value")
        if extra is not None:
            extra.pop("_rays", None); extra.pop("_e2e_value", None)
            line["configs"] = [extra]
        if world == 1 and not args.no_cpu_baseline and not args.profile:
            try:
                rc = reference_cuda_leg(make_workload(args.workload), rays)
            except Exception as e:      # measurement of the other implementation must never take the product's line down
                rc = {"unavailable": "%s: %s" % (type(e).__name__, e)}
            if rc is not None:
                if "value" in rc:
                    rc["speedup_device"] = line["value"] / rc["value"]
                    rc["speedup_e2e"] = e2e_value / rc["e2e_value"]
                line["reference_cuda"] = rc
        bad = line.get("gather_parity") and not all(v for k, v in line["gather_parity"].items() if k.endswith("_equal"))
        print(json.dumps(line))
        if bad:
            sys.stderr.write("bench.py: the gathered frame differs from the single-GPU frame\n")
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0 and line.get("gather_parity") and not all(v for k, v in line["gather_parity"].items() if k.endswith("_equal")):
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=WORKLOADS)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--leaf-max", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="--gpus 8: skip the BASELINE config 5 record (configs[0] of the line)")
    ap.add_argument("--chunks", type=int, default=0, help="multi-GPU tile ownership bands per rank (0 = library default)")
    ap.add_argument("--gather", default="auto", choices=["auto", "nccl", "peer"], help="multi-GPU tile delivery to rank 0")
    ap.add_argument("--no-shadows", action="store_true", help="primary rays only (diagnostic; not the headline workload)")
    ap.add_argument("--profile", action="store_true", help="short run for ncu: honour --warmup < 3, skip the CPU baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
