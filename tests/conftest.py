import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def frog_scene():
    import numpy as np
    from raytracinginonesemester_b200 import api, scenes
    d = np.load(os.path.join(GOLDEN, "frog_mesh.npz"))
    return api.Scene(d["positions"], d["indices"], normals=d["normals"], tri_obj_ids=d["tri_obj_ids"],
                     materials=[api.make_material(**scenes.FROG_MATERIAL)])


@pytest.fixture(scope="session")
def renderer():
    """One context on cuda:0 through the C ABI.  Fails (does not skip) when the CUDA library is
    missing: the product has no CPU path."""
    from raytracinginonesemester_b200 import api
    r = api.Renderer(0)
    yield r
    r.close()
