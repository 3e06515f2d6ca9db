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


@pytest.mark.parametrize("name", ["ha", "sw"])
def test_interpolate_at_user_polar_angles(name):
    parity_suite.check_interpolate_vs_golden(pd, name)


@pytest.mark.parametrize("name,ncol,first", [("sw", 4, 1000), ("lw", 16, 5000), ("ha", 1, 700), ("tp1", 6, 0), ("tp9c", 2, 0)])
def test_ensemble_slices_vs_live_oracle(name, ncol, first):
    """Small version of the GPU suite's 256 / 1,024 / 64-column comparison (same code path, no process pool)."""
    parity_suite.check_ensemble_vs_live_oracle(pd.pydisort, name, ncol, first)


@pytest.mark.parametrize("name,ncol", [("sw", 3), ("lw", 8), ("ha", 1), ("tp9c", 2), ("tp1", 6)])
def test_interface_levels_come_from_the_sweep_and_agree_with_the_assembled_solution(name, ncol):
    parity_suite.check_interface_levels_vs_assembled(pd.pydisort, name, ncol)


def test_actinic_fluxes():
    parity_suite.check_actinic_vs_golden(pd)


def test_thermal_source_inputs():
    parity_suite.check_thermal_inputs_vs_golden(pd)


def test_hapke_fourier_modes():
    parity_suite.check_hapke_modes_vs_golden(pd)


def test_production_paths_cover_the_production_shapes_and_match_the_generic_kernels():
    """N = 4 / 8 items go through the Cholesky + Jacobi path and every system through the interface-radiance
    elimination (no hand-backs on the ensembles); the size-generic kernels (PD_FLAG_GENERIC_KERNELS) give the same
    answers."""
    from pythonic_disort_b200 import _lib, api, synthetic
    lib = api._backend()[0]
    lib.pd_hostsim_count.restype = __import__("ctypes").c_long

    def generic(*args, **kwargs):
        return pd.pydisort(*args, _kernel_flags=_lib.PD_FLAG_GENERIC_KERNELS, **kwargs)

    for name, ncol in (("sw", 3), ("lw", 6)):
        ens = synthetic.make(name, ncol)
        lib.pd_hostsim_reset()
        fast = parity_suite.run_batched(pd.pydisort, ens)
        counts = [lib.pd_hostsim_count(i) for i in range(4)]
        assert counts[0] > 0 and counts[1] == 0 and counts[2] > 0 and counts[3] == 0, (name, counts)
        lib.pd_hostsim_reset()
        slow = parity_suite.run_batched(generic, ens)
        counts = [lib.pd_hostsim_count(i) for i in range(4)]
        assert counts[0] == 0 and counts[2] == 0, (name, counts)
        for key in fast:
            scale = np.max(np.abs(slow[key]))
            assert np.max(np.abs(fast[key] - slow[key])) <= 1e-11 * scale, (name, key)
