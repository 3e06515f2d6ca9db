"""Build and inject the host-side debug build of the kernels (tests/hostsim).

TEST INFRASTRUCTURE: lets the CPU test-suite drive the package's Python host
logic and the kernels' arithmetic without a GPU.  The package itself never
does this; outside these tests `pydisort` raises if CUDA is unavailable."""
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "pd_hostsim.cpp")
LIB = os.path.join(HERE, "hostsim", "libpd_hostsim.so")
CSRC = os.path.join(os.path.dirname(HERE), "pythonic_disort_b200", "csrc")


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC],
                       check=True)
    return LIB


class use:
    """Context manager: route pythonic_disort_b200.api through the host build."""

    def __enter__(self):
        from pythonic_disort_b200 import _lib, api
        self.api = api
        self.prev = api._backend
        self.lib = _lib.bind(build())
        backend = (self.lib, torch.device("cpu"))
        api._backend = lambda: backend  # the package has no such seam: the test replaces the function
        return self

    def __exit__(self, *exc):
        self.api._backend = self.prev
