#!/usr/bin/env python
"""OBJ ingest: device parser (rt_dmesh_parse_obj) against the host loader (rt_mesh_load_obj) on the C4 terrain written out as
an OBJ file (1M triangles, 501 501 vertices, ~45 MB of text).  Prints one JSON line."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401
from raytracinginonesemester_b200 import Renderer, api, scenes  # noqa: E402
from raytracinginonesemester_b200.api import DeviceMesh, DeviceScene  # noqa: E402

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 500)
pos, idx = scenes.terrain(nx, ny)
t0 = time.time()
buf = ["v %.9g %.9g %.9g" % tuple(p) for p in pos] + ["f %d %d %d" % tuple(t) for t in (idx.astype(np.int64) + 1)]
text = ("\n".join(buf) + "\n").encode()
gen_s = time.time() - t0
path = os.path.join(tempfile.mkdtemp(), "terrain.obj")
open(path, "wb").write(text)
r = Renderer(0)
ms = []
for rep in range(4):
    t0 = time.perf_counter()
    dm, nid = DeviceMesh.parse_obj(r, text)
    wall = time.perf_counter() - t0
    ms.append((dm.stats()[0], wall * 1e3))
    st = dm.stats()
    if rep < 3:
        dm.close()
t0 = time.perf_counter()
hp, hn, hi, ho, hid = api.load_obj(path)
host_s = time.perf_counter() - t0
dp, dn, di, do = dm.download()
same = bool(np.array_equal(hp.view(np.uint32), dp.view(np.uint32)) and np.array_equal(hi, di) and np.array_equal(ho, do))
t0 = time.perf_counter()
info = r.upload_scene(DeviceScene(dm, materials=[api.make_material(**scenes.TERRAIN_MATERIAL)]))
up_s = time.perf_counter() - t0
print(json.dumps({"triangles": int(hi.shape[0]), "vertices": int(hp.shape[0]), "text_bytes": len(text), "lines": st[1],
                  "device_parse_ms_incl_h2d": round(min(m[0] for m in ms[1:]), 3), "device_parse_wall_ms": round(min(m[1] for m in ms[1:]), 3),
                  "device_text_GB_s": round(len(text) / (min(m[0] for m in ms[1:]) * 1e-3) / 1e9, 2),
                  "numbers_converted_by_host_strtof": st[2], "host_loader_ms_1_thread": round(host_s * 1e3, 1),
                  "speedup_vs_host_loader": round(host_s * 1e3 / min(m[1] for m in ms[1:]), 1), "arrays_identical": same,
                  "upload_from_device_arrays_wall_ms": round(up_s * 1e3, 2), "bvh_build_ms": round(float(info.build_ms), 3)}))
