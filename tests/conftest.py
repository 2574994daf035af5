import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def port_oracle():
    """The plain-C restatement (oracle/bp_oracle.c); built on demand."""
    import oracle
    if not oracle.have_port():
        oracle.build()
    return oracle.PortOracle()


@pytest.fixture(scope="session")
def ref_oracle():
    """The unmodified reference C++ (oracle/_ref); skipped when it was never built (needs /root/reference)."""
    import oracle
    if not oracle.have_ref():
        if os.path.isdir("/root/reference/src_cpp"):
            oracle.build()
        else:
            pytest.skip("oracle/_ref/libref_bp.so not present and /root/reference absent")
    return oracle.RefOracle()
