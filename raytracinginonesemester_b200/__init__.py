"""B200-native ray-casting hot path (camera rays -> BVH -> Möller–Trumbore closest hit -> HW1/HW2
shading -> PPM), a drop-in for the renderers of nirajbabar/raytracinginonesemester.
The product is the CUDA library behind include/rt_api.h; this package is its thin host mirror."""
from ._abi import *  # noqa: F401,F403
from .api import (Frame, Renderer, RtError, Scene, camera_init, jitter_table, load_library, make_light,  # noqa: F401
                  make_material)
