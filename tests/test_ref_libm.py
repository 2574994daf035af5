"""CPU: ldpc_b200/csrc/ref_libm.h (the glibc log / expm1 / tanh restatement the product-sum kernels use) is
compiled for the host and compared bit-for-bit with the live libm on random and edge-case arguments."""
import os
import platform
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_ref_libm_bit_exact_against_glibc(tmp_path):
    if platform.machine() != "x86_64":
        pytest.skip("the restatement follows the x86-64 FMA variants of glibc")
    flags = open("/proc/cpuinfo").read()
    if " fma " not in flags or " avx2 " not in flags:
        pytest.skip("host CPU lacks FMA/AVX2: glibc selects a different libm variant here")
    exe = str(tmp_path / "ref_libm_check")
    src = os.path.join(HERE, "native", "ref_libm_check.c")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", exe, src, "-lm"], check=True)
    res = subprocess.run([exe, "1500000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "log 0 expm1 0 tanh 0 chain 0 mismatches" in res.stdout
