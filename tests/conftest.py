import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def atmospheres():
    from tools import atmospheres as A
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = getattr(A, name)()
        return cache[name]
    return get


@pytest.fixture(scope="session")
def oracle_factory():
    from oracle_lib import Oracle

    def make(atm, l=0, photon_source=1):
        o = Oracle()
        depth = o.set_atmosphere(atm, l, photon_source)
        return o, depth
    return make


@pytest.fixture(scope="session")
def gpu_factory():
    """libartes_gpu context loaded with an atmosphere; fails (not skips) without the CUDA library."""
    from artes_b200.lib import GpuTransport
    from artes_b200 import host

    def make(atm, l=0, photon_source=1):
        g = GpuTransport((0,))
        g.set_grid(atm.rfront, atm.thetafront(), atm.thetaplane(), atm.phifront())
        depth = host.cell_depth(atm.rfront, atm.k_sca[l], atm.k_abs[l], atm.nr, atm.ntheta, atm.nphi, photon_source)
        g.set_wavelength(atm.k_sca[l], atm.k_abs[l], atm.uniq[l], atm.cell_to_uniq[l], depth)
        return g, depth
    return make
