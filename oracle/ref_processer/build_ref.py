"""TEST INFRASTRUCTURE: compile the UNMODIFIED reference ``dataset/processer.pyx`` (Cython/C++: affine crop + augmentation + label
rasterisation of the train1 input pipeline) from the sources where they lie under /root/reference into
``oracle/_ref/ref_processer*.so`` (git-ignored).  Cython writes its generated .cpp next to the .pyx, and /root/reference is
read-only, so the .pyx is copied to a scratch directory for the build; nothing of it enters the repository.  Build flags: -O2
-ffp-contract=off -mno-fma after the file's own "-O3 -march=native", so that float32 products and sums round as written
(the golden vectors made with it are committed; the GPU box needs neither this module nor /root/reference).

    python oracle/ref_processer/build_ref.py      ->  oracle/_ref/ref_processer.cpython-312-x86_64-linux-gnu.so
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref")

_SETUP = '''
import numpy as np
from Cython.Build import cythonize
from setuptools import Extension, setup
ext = Extension("ref_processer", ["ref_processer.pyx"], include_dirs=[np.get_include()], language="c++",
                extra_compile_args=["-O2", "-ffp-contract=off", "-mno-fma"])
setup(name="ref_processer", ext_modules=cythonize([ext], language_level=3, quiet=True), script_args=["build_ext", "--inplace", "-q"])
'''


def build() -> str:
    src = os.path.join(REF, "dataset", "processer.pyx")
    if not os.path.exists(src):
        raise FileNotFoundError(src)
    os.makedirs(OUT, exist_ok=True)
    have = glob.glob(os.path.join(OUT, "ref_processer*.so"))
    if have and os.path.getmtime(have[0]) >= os.path.getmtime(src):
        return have[0]
    with tempfile.TemporaryDirectory() as work:
        shutil.copy(src, os.path.join(work, "ref_processer.pyx"))
        with open(os.path.join(work, "setup_ref.py"), "w") as f:
            f.write(_SETUP)
        r = subprocess.run([sys.executable, "setup_ref.py"], cwd=work, capture_output=True, text=True)
        built = glob.glob(os.path.join(work, "ref_processer*.so"))
        if r.returncode != 0 or not built:
            raise RuntimeError("reference processer.pyx did not build:\n" + (r.stdout + r.stderr)[-3000:])
        dst = os.path.join(OUT, os.path.basename(built[0]))
        shutil.copy(built[0], dst)
    return dst


def load():
    """Import the compiled reference module (needs /root/reference on sys.path for its ``import util_func``)."""
    path = build()
    for p in (REF, OUT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    return importlib.import_module("ref_processer")


if __name__ == "__main__":
    print(build())
