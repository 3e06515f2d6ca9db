"""CPU tests for the synthetic ensembles: generator determinism and the oracle
against the reference outputs stored in tests/golden/ensemble_*.npz."""
import os
import warnings

import numpy as np
import pytest

import golden_io
from oracle import disort_oracle
from pythonic_disort_b200 import synthetic


def test_subset_property():
    a = synthetic.make("sw", 8)
    b = synthetic.make("sw", 3, first=5)
    for x, y in zip(a["args"], b["args"]):
        if isinstance(x, np.ndarray):
            np.testing.assert_array_equal(x[5:8], y)
    np.testing.assert_array_equal(a["kwargs"]["BDRF_Fourier_modes"][0][5:8], b["kwargs"]["BDRF_Fourier_modes"][0])


@pytest.mark.parametrize("name,ncheck", [("sw", 4), ("lw", 16), ("ha", 1), ("tp9c16", 1)])
def test_oracle_matches_reference_on_ensemble(name, ncheck):
    gold = np.load(os.path.join(golden_io.GOLDEN, f"ensemble_{name}.npz"))
    ens = synthetic.make(name, int(gold["ncol"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = synthetic.run_reference_like(disort_oracle.pydisort, ens, columns=range(ncheck))
    for key, val in got.items():
        for b in range(ncheck):
            err, _, masked = golden_io.parity(val[b], gold[key][b])
            assert err <= 1e-9, (name, key, b, err)
