import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _cuda_device_count() -> int:
    """CUDA devices visible to the driver, asked through libcudart directly (no torch import, no product code)."""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            break
        except OSError:
            continue
    else:
        return 0
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def pytest_collection_modifyitems(config, items):
    """On a box without a CUDA device, `gpu`-marked tests are skipped instead of failing in bpb_create."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (ldpc_b200 has no CPU decode path)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
