"""In-tree native build + editable install.  `python setup.py build_ext --inplace` (or `pip install -e .`
with no build isolation) compiles csrc/ into asyncframework_b200/lib/*.so via asyncframework_b200/build.py
(nvcc -gencode arch=compute_100a,code=sm_100a; g++ for the host runtime)."""
import os
import sys

from setuptools import Command, find_packages, setup
from setuptools.command.build_ext import build_ext as _build_ext

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


class build_ext(_build_ext):
    def run(self):
        from asyncframework_b200 import build as b
        b.build_all(force=False, verbose=True)


class BuildNative(Command):
    description = "compile the C++ runtime and the sm_100a CUDA kernels in-tree"
    user_options = []

    def initialize_options(self): pass
    def finalize_options(self): pass

    def run(self):
        from asyncframework_b200 import build as b
        b.build_all(force=True, verbose=True)


setup(
    name="asyncframework_b200",
    version="0.1.0",
    description="Blackwell-native asynchronous-optimization engine (ASGD / ASAGA parameter server over NVLink)",
    packages=find_packages(include=["asyncframework_b200", "asyncframework_b200.*"]),
    package_data={"asyncframework_b200": ["lib/*.so"]},
    python_requires=">=3.10",
    install_requires=["torch", "numpy", "cloudpickle"],
    extras_require={"crypto": ["cryptography"]},
    cmdclass={"build_ext": build_ext, "build_native": BuildNative},
    entry_points={"console_scripts": ["afb-launch=asyncframework_b200.launch:main", "afb-submit=asyncframework_b200.submit:main",
                                      "afb-shell=asyncframework_b200.shell:main", "afb-master=asyncframework_b200.deploy.master:main",
                                      "afb-worker=asyncframework_b200.deploy.worker:main",
                                      "afb-history-server=asyncframework_b200.utils.status_server:main"]},
)
