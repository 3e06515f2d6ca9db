"""CPU tests of the host-side logic around the batched solve: the pipelined ensemble driver
(pythonic_disort_b200.ensemble.solve_ensemble: chunking, assembly of the outputs, deferred input checks) and the
rules that decide which inputs carry the column axis (api.carries_batch_axis / slice_columns /
parallel.shard_inputs).  The kernels behind it are the host build of the CUDA sources (tests/hostsim)."""
import warnings

import numpy as np
import pytest

import hostsim_backend
import parity_suite
import pythonic_disort_b200 as pd
from pythonic_disort_b200 import api, ensemble, parallel, synthetic


@pytest.fixture(scope="module", autouse=True)
def host_build():
    with hostsim_backend.use():
        yield


@pytest.mark.parametrize("name,ncol,chunk", [("sw", 5, 2), ("lw", 7, 3), ("tp9c", 3, 2), ("ha", 2, 1)])
def test_solve_ensemble_equals_the_batched_call(name, ncol, chunk):
    ens = synthetic.make(name, ncol)
    ref = parity_suite.run_batched(pd.pydisort, ens)
    outputs = ("flux_up", "flux_down", "u0") + (("u",) if "u" in ens["outputs"] else ())
    res = ensemble.solve_ensemble(*ens["args"], tau=ens["tau_eval"], phi=ens["phi_eval"], outputs=outputs, chunk=chunk,
                                  **ens["kwargs"])
    assert res.chunks == -(-ncol // chunk)
    for key in res:
        np.testing.assert_array_equal(res[key], ref[key])
    # buffers of a previous result are recycled, asynchronous mode gives the same numbers after wait()
    again = ensemble.solve_ensemble(*ens["args"], tau=ens["tau_eval"], phi=ens["phi_eval"], outputs=outputs,
                                    chunk=chunk, out=res, wait=False, **ens["kwargs"])
    again.wait()
    assert all(again.tensors[k] is res.tensors[k] for k in res.tensors)
    np.testing.assert_array_equal(again["flux_up"], ref["flux_up"])


@pytest.mark.parametrize("name,ncol,chunk", [("sw", 4, 3), ("lw", 7, 3), ("ha", 2, 1)])
def test_inputs_described_per_layer_and_expanded_on_the_device(name, ncol, chunk):
    parity_suite.check_compact_inputs(pd, name, ncol, chunk)


def test_solve_ensemble_at_user_polar_angles():
    ens = synthetic.make("ha", 2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pd.pydisort(*ens["args"], **ens["kwargs"])
    ref = pd.subroutines.interpolate(out[4])(ens["mu_user"], ens["tau_eval"], ens["phi_eval"])
    res = ensemble.solve_ensemble(*ens["args"], tau=ens["tau_eval"], phi=ens["phi_eval"], mu=ens["mu_user"],
                                  outputs=("u",), chunk=1, **ens["kwargs"])
    np.testing.assert_array_equal(res["u"], ref)


def test_solve_ensemble_raises_the_reference_errors_after_the_last_chunk():
    ens = synthetic.make("lw", 4)
    args = list(ens["args"])
    args[1] = args[1].copy()
    args[1][3, 5] = 1.5  # omega out of range in the second chunk
    with pytest.raises(ValueError, match="Single-scattering albedo"):
        ensemble.solve_ensemble(*args, tau=ens["tau_eval"], chunk=2, **ens["kwargs"])
    with pytest.raises(ValueError, match="outside the tau range"):
        ensemble.solve_ensemble(*ens["args"], tau=ens["tau_eval"] + 100.0, chunk=2, **ens["kwargs"])


def _tiny(B, L, NQuad, **kw):
    rng = np.random.default_rng(7)
    tau = np.cumsum(0.1 + rng.random((B, L)), axis=1)
    omega = 0.2 + 0.6 * rng.random((B, L))
    g = 0.3 + 0.4 * rng.random(L)
    leg = g[:, None] ** np.arange(NQuad + 1)[None, :]  # SHARED phase function [L, NLeg_all]
    return (tau, omega, NQuad, leg, 0.5, 1.0, 0.0), kw


def test_shared_inputs_are_not_cut_when_a_dimension_happens_to_equal_B():
    """with B == L a shared Leg_coeffs_all [L, NLeg_all] used to be sliced like a per-column input."""
    B = L = 4
    args, kw = _tiny(B, L, 4)
    a, k, (lo, hi) = parallel.shard_inputs(B, args, kw, rank=1, world=2)
    assert (lo, hi) == (2, 4) and a[0].shape == (2, L) and a[3].shape == args[3].shape
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        full = pd.pydisort(*args, **kw)[1](0.05)
        part = pd.pydisort(*a, **k)[1](0.05)
    np.testing.assert_array_equal(part, full[lo:hi])


def test_boundary_values_with_B_equal_N():
    """with B == N a reference-shaped b_pos [N] was silently read as B per-column scalars."""
    B, NQuad = 4, 8
    args, kw = _tiny(B, 3, NQuad)
    b_vec = np.array([0.1, 0.2, 0.3, 0.4])
    with pytest.raises(ValueError, match="ambiguous"):
        pd.pydisort(*args, b_pos=b_vec, **kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        per_col = pd.pydisort(*args, b_pos=b_vec[:, None], **kw)[1](0.05)            # [B, 1]: one value per column
        shared = pd.pydisort(*args, b_pos=np.tile(b_vec, (B, 1)), **kw)[1](0.05)     # [B, N]: the same vector everywhere
        single = [pd.pydisort(args[0][b], args[1][b], *args[2:], b_pos=b_vec)[1](0.05) for b in range(B)]
        single_iso = [pd.pydisort(args[0][b], args[1][b], *args[2:], b_pos=float(b_vec[b]))[1](0.05) for b in range(B)]
    np.testing.assert_allclose(shared, np.array(single), rtol=1e-13)
    np.testing.assert_allclose(per_col, np.array(single_iso), rtol=1e-13)
    # the sharding rule reads the shapes the same way
    assert api.carries_batch_axis("b_pos", b_vec[:, None], B, 4, 8)
    assert api.carries_batch_axis("b_pos", np.tile(b_vec, (B, 1)), B, 4, 8)
    assert not api.carries_batch_axis("b_pos", np.array([0.1, 0.2, 0.3]), 5, 3, 6)


def test_unsupported_stream_counts_are_rejected_up_front():
    """NQuad beyond what one CTA's shared memory holds used to fail late with an opaque code."""
    leg = np.ones((1, 201)) * 0.0
    leg[0, 0] = 1.0
    with pytest.raises(ValueError, match="NQuad"):
        pd.pydisort(np.array([1.0]), np.array([0.5]), 200, leg, 0.5, 1.0, 0.0)
