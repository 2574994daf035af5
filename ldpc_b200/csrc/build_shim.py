"""Build the Cython binding ldpc_b200/_bp_shim (in place).  Run after libbp_b200.so exists:
    python ldpc_b200/csrc/build_shim.py
Links against ldpc_b200/libbp_b200.so with an $ORIGIN rpath, so the pair travels together."""
import os
import sys

from Cython.Build import cythonize
from setuptools import Extension, setup

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(ROOT)
ext = Extension("ldpc_b200._bp_shim", ["ldpc_b200/_bp_shim.pyx"], include_dirs=[os.path.join(ROOT, "include")],
                library_dirs=[os.path.join(ROOT, "ldpc_b200")], libraries=["bp_b200"],
                runtime_library_dirs=["$ORIGIN"], extra_compile_args=["-O2", "-w"])
tmp = os.path.join(ROOT, "ldpc_b200", "csrc", "_obj", "shim")
sys.argv = [sys.argv[0], "build_ext", "--inplace", "-q", "--build-temp", tmp, "--build-lib", os.path.join(tmp, "lib")]
setup(name="ldpc_b200_shim", ext_modules=cythonize([ext], quiet=True, language_level=3), script_args=sys.argv[1:])
