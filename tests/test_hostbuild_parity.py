"""CPU tests (-m "not gpu"): the package's Python host logic driving a host
build of the very same kernel routines (tests/hostsim, one lane per group)
against the reference goldens.  This checks the arithmetic of the CUDA code
and the batching / squeezing / error behaviour of the host wrapper; the GPU
run of the same checks is tests/test_gpu_parity.py."""
import numpy as np
import pytest

import golden_io
import hostsim_backend
import parity_suite
import pythonic_disort_b200 as pd

SUITE = golden_io.suite_names()


@pytest.fixture(scope="module", autouse=True)
def host_build():
    with hostsim_backend.use():
        yield


@pytest.mark.parametrize("name", SUITE)
def test_reference_suite_outputs(name):
    parity_suite.check_suite_record_outputs(pd.pydisort, name, max_records=6)


@pytest.mark.parametrize("name", [n for n in SUITE if golden_io.load_test(n)[1]])
def test_stamnes_criteria(name):
    parity_suite.check_stamnes(pd.pydisort, name)


@pytest.mark.parametrize("name", ["sw", "lw", "ha", "tp9c16"])
def test_ensembles(name):
    parity_suite.check_ensemble_vs_golden(pd.pydisort, name)
