"""Statistical fixture for the soft shadows of RT_MODE_HW2_CPU, produced by the UNMODIFIED HW2/HW2/CPUOnly
reference sources compiled in place (oracle/_ref/libref_cpuonly.so).

The reference samples its disk lights with a process-wide std::mt19937 seeded by std::random_device
(CPUOnly/include/raytracer.h:12-16): no two runs agree, so the fixture holds, per pixel and channel, the MEAN and the
STANDARD DEVIATION of 64 reference runs of config/sphere_area.json (radius 0.15, 8 shadow samples) at 120x80, plus the
scene as the tests consume it.  tests/golden/cpuonly_area.npz.  Authoring container only (needs /root/reference)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden_cpuonly as G  # noqa: E402  (scene loading through the reference's own loader)
from raytracinginonesemester_b200 import _abi as A  # noqa: E402

lib, fp, f3, L = G.lib, G.fp, G.f3, G.L
RUNS, W, H = 64, 120, 80

j, pos, nrm, idx, obj, mats = G.load_scene("sphere_area.json")
cam, li = j["camera"], j["light"]
marr = (A.rt_material * len(mats))(*mats)
lights = (L * 1)(L(tuple(li["position"]), tuple(li["color"]), float(li["intensity"])))
radius = np.array([li["radius"]], np.float32)
nsamp = np.array([li["shadow_samples"]], np.int32)
sensor_h, sensor_w = float(cam.get("sensor_height_mm", 24.0)), float(cam.get("sensor_width_mm", 36.0))
mean = np.zeros((H, W, 3), np.float64)
meansq = np.zeros((H, W, 3), np.float64)
lib.ref_cpu_render_area_stats(fp(pos), fp(nrm), C.c_uint64(len(pos)), idx.ctypes.data_as(A.u32p), C.c_uint64(len(idx)), obj.ctypes.data_as(A.i32p),
                              marr, len(mats), fp(f3(cam["position"])), fp(f3(cam["look_at"])), fp(f3(cam["up"])),
                              C.c_double(cam["focal_length_mm"]), C.c_double(sensor_h), C.c_double(sensor_w), W, H, lights,
                              fp(radius), nsamp.ctypes.data_as(A.i32p), 1, int(j["settings"]["max_bounces"]), RUNS,
                              mean.ctypes.data_as(C.POINTER(C.c_double)), meansq.ctypes.data_as(C.POINTER(C.c_double)))
std = np.sqrt(np.maximum(meansq - mean * mean, 0.0) * RUNS / (RUNS - 1))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cpuonly_area.npz"),
                    positions=pos, normals=nrm, indices=idx, tri_obj_ids=obj, materials=np.stack([G.mat13(m) for m in mats]),
                    camera=np.array(list(cam["position"]) + list(cam["look_at"]) + list(cam["up"]) + [cam["focal_length_mm"], sensor_h, sensor_w], np.float64),
                    light=np.array(list(li["position"]) + list(li["color"]) + [li["intensity"], li["radius"], li["shadow_samples"]], np.float64),
                    frame=np.array([W, H, int(j["settings"]["max_bounces"]), RUNS]), mean=mean.astype(np.float32), std=std.astype(np.float32))
pen = (std > 1e-6).any(-1)
print("cpuonly_area.npz: %d runs, %d of %d pixels have a penumbra (std > 0), max std %.4f" % (RUNS, pen.sum(), pen.size, std.max()))
