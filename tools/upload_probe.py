import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from raytracinginonesemester_b200 import api, scenes
r = api.Renderer(0)
t=time.perf_counter(); sc = scenes.terrain_scene(2500, 2000); print("gen", time.perf_counter()-t)
for i in range(3):
    t=time.perf_counter(); info = r.upload_scene(sc); print("upload wall %.3f s  build_ms %.2f upload_ms %.2f"%(time.perf_counter()-t, info.build_ms, info.upload_ms))
