import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vi-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand with gcc."""
    from oracle import vso
    vso.build()
    vso.lib()
    return vso


@pytest.fixture(scope="session")
def ctx():
    """One vislam_b200 context on cuda:0.  GPU tests call the product ONLY through the C ABI."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vislam_b200 as vb
    c = vb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def pair_small():
    """A small synthetic frame pair (EuRoC-shaped 752x480, 300 features) shared across tests."""
    from vislam_b200 import synth
    return synth.make_pair(n_feat=300, seed=1001)
