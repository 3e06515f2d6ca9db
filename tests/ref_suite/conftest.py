"""Runs the reference's unmodified pydisotest files against pythonic_disort_b200 (see README.md in this directory)."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))  # tests/: hostsim_backend
sys.path.insert(0, HERE)                   # ARTS_data

import pythonic_disort_b200  # noqa: E402
from pythonic_disort_b200 import subroutines  # noqa: E402

# the DISORT 4.0.99 result files are links into tests/golden/stamnes; copy them if a snapshot dropped the links
_src = os.path.join(os.path.dirname(HERE), "golden", "stamnes")
for _f in os.listdir(_src):
    _dst = os.path.join(HERE, "Stamnes_results", _f)
    if not os.path.exists(_dst):
        import shutil
        if os.path.islink(_dst):
            os.unlink(_dst)
        shutil.copyfile(os.path.join(_src, _f), _dst)

sys.modules.setdefault("PythonicDISORT", pythonic_disort_b200)
sys.modules.setdefault("PythonicDISORT.subroutines", subroutines)


def pytest_generate_tests(metafunc):
    if "pd_backend" in metafunc.fixturenames:
        metafunc.parametrize("pd_backend", [pytest.param("host"), pytest.param("cuda", marks=pytest.mark.gpu)],
                             indirect=True)


@pytest.fixture(autouse=True)
def pd_backend(request, monkeypatch):
    monkeypatch.chdir(HERE)
    if request.param == "cuda":
        import torch
        assert torch.cuda.is_available(), "-m gpu needs a GPU"
        yield "cuda"
    else:
        import hostsim_backend
        with hostsim_backend.use():
            yield "host"
