import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle.oracle import Oracle

    return Oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle.oracle import Oracle, have_reference

    if not have_reference():
        pytest.skip("oracle/_ref/libref_harness.so not built (needs /root/reference at build time)")
    return Oracle("reference")


@pytest.fixture(scope="session")
def gpu():
    import restir_b200 as rb

    rb.init(0)   # raises RestirError when the CUDA extension or a device is missing: no silent fallback
    return rb
