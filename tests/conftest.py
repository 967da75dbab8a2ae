import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box via gpurun)")


@pytest.fixture(scope="session")
def small_bunny():
    from fspt_b200 import scenes
    return scenes.bunny_class(subdiv=4, atlas_res=64, env_size=(256, 128))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle
