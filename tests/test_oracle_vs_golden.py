"""CPU tests: the oracle (oracle/disort_oracle.py) against the golden vectors
produced by the unmodified reference, and against Stamnes' DISORT 4.0.99
results with the reference suite's own pass criteria (pydisotest/1_test.py:78-81)."""
import warnings

import numpy as np
import pytest

import golden_io
from oracle import disort_oracle

SUITE = golden_io.suite_names()


def _to_np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


@pytest.mark.parametrize("name", SUITE)
def test_oracle_matches_reference_outputs(name):
    records, _ = golden_io.load_test(name)
    records = records[:12]  # 8ARTS_A holds 101 near-identical no-scattering solves
    for rec in records:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = disort_oracle.pydisort(*rec["args"], **rec["kwargs"])
        tol = golden_io.conditioning_tolerance(rec["args"][1])
        for call, got in golden_io.run_calls(out, rec):
            scale = golden_io.group_scale(call["outs"])
            for g, r in zip(got, call["outs"]):
                scale_err, _, _ = golden_io.parity(np.squeeze(_to_np(g)), np.squeeze(r), scale=scale)
                assert scale_err <= tol, (name, call["fn"], call["anti"], scale_err, tol)


@pytest.mark.parametrize("name", [n for n in SUITE if golden_io.load_test(n)[1] and n != "9corrections"])
def test_oracle_passes_stamnes_criteria(name):
    records, compares = golden_io.load_test(name)
    from pythonic_disort_b200.subroutines import _compare
    for cmp in compares:
        rec = records[cmp["record"]]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = disort_oracle.pydisort(*rec["args"], **rec["kwargs"])
        res = _compare(golden_io.stamnes(cmp["file"]), cmp["mu_to_compare"], cmp["reorder_mu"],
                       out[1], out[2], out[4] if cmp["has_u"] else None)
        for d, ratio in zip(res[0:6:2], res[1:6:2]):
            assert np.max(ratio[d > 1e-3], initial=0) < 1e-3
        if cmp["has_u"]:
            assert np.max(res[7][res[6] > 1e-3], initial=0) < 1e-2


def test_oracle_corrections_improve_on_stamnes():
    """pydisotest/9_test.py:368-370 -- delta-M + NT must shrink the mean error."""
    records, compares = golden_io.load_test("9corrections")
    from pythonic_disort_b200.subroutines import _compare
    res = []
    for cmp in compares:
        rec = records[cmp["record"]]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = disort_oracle.pydisort(*rec["args"], **rec["kwargs"])
        res.append(_compare(golden_io.stamnes(cmp["file"]), cmp["mu_to_compare"], cmp["reorder_mu"],
                            out[1], out[2], out[4]))
    plain, corrected = res
    assert np.mean(plain[0] - corrected[0]) > 0
    assert np.mean(plain[2] - corrected[2]) > 0
    assert np.mean(plain[6] - corrected[6]) > 0


def test_oracle_with_interpolate_reproduces_the_reference_at_user_polar_angles():
    """Row f1: the oracle's u through the host-side interpolate() against the reference's interpolate() output
    (tests/golden/interpolate.npz) -- pins the CPU arm of the HA bench, which evaluates at user polar angles."""
    import os
    import warnings

    import golden_io
    from oracle import disort_oracle
    from pythonic_disort_b200 import synthetic
    gold = np.load(os.path.join(golden_io.GOLDEN, "interpolate.npz"))
    ens = synthetic.make("ha", 1)
    ens["mu_user"] = gold["mu_user"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = synthetic.run_reference_like(disort_oracle.pydisort, ens, at_user_mu=True)["u"][0]
    err, _, _ = golden_io.parity(got, gold["ha_u"][0])
    assert err <= 1e-9, err
