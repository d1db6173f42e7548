"""`pip install .` for the B200-native torchlsq (replaces /root/reference/setup.py + setup_utils.py + torch_patch.py:
no multi-arch list, no CPU build, no header patching).

The build step is one `make`: hand-written CUDA compiled for sm_100a into the C-ABI library
`torchlsq/libtorchlsq_b200.so` (include/lsq_b200.h), and the thin C++ dispatcher / autograd binding above it
(csrc/torch_binding.cpp, g++ against the installed torch headers) into `torchlsq/_C.so`, which the package loads with
torch.ops.load_library as the reference loads its `_C`.  nvcc and g++ must be on PATH (or pass NVCC= / CXX=).  The in-tree
libraries are what the repo's tests and bench use; installing is optional.

    pip install --no-build-isolation ./lsqfakequantize-pytorch_b200
"""
import os
import shutil
import subprocess
from pathlib import Path

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py

HERE = Path(__file__).resolve().parent
LIBS = ("libtorchlsq_b200.so", "_C.so")


class BuildWithKernels(build_py):
    def run(self):
        env = dict(os.environ)
        subprocess.check_call(["make", "-C", str(HERE / "csrc"), f"-j{os.cpu_count() or 4}"], env=env)
        super().run()
        for lib in LIBS:
            dst = Path(self.build_lib) / "torchlsq" / lib
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copy2(HERE / "torchlsq" / lib, dst)


setup(
    name="torchlsq",
    version="2.1+b200.r2",
    description="LSQ+ fake-quantize for PyTorch: hand-written sm_100a (B200) kernels behind the torchlsq 2.1 API",
    packages=find_packages(include=["torchlsq", "torchlsq.*"]),
    package_data={"torchlsq": list(LIBS)},
    include_package_data=True,
    python_requires=">=3.9",
    install_requires=[],          # torch is expected in the environment (_C.so links against the torch it was built with)
    cmdclass={"build_py": BuildWithKernels},
    zip_safe=False,
)
