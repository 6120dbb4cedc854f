import os, sys
sys.path.insert(0, "/root/repo")
import torch
from raytracinginonesemester_b200 import Renderer, scenes, _abi as A
r = Renderer(0)
r.upload_scene(scenes.terrain_scene(1000, 500))
fr = scenes.terrain_frame(3840, 2160)
fr.kernel_variant = A.RT_VARIANT_STATS
r.render(fr); r.sync()
nv, nt, nl, nb = r.frame_stats()
print("lib", os.environ.get("RT_B200_LIB"), "lane node tests", nv, "lane tri tests", nt, "node lines", nl, "tri blocks (warp tests)", nb, "lanes per warp test %.2f" % (nt / max(nb, 1)))
