"""CPU test of the N>1 path: two gloo ranks each solve their shard of an
ensemble (host build of the kernels), all-gather the fluxes, and the result
must equal the single-process solve bit for bit."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import warnings

    import torch.distributed as dist

    import hostsim_backend
    import pythonic_disort_b200 as pd
    from pythonic_disort_b200 import parallel, synthetic
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ens = synthetic.make("lw", 11)
    args, kwargs, (lo, hi) = parallel.shard_inputs(ens["B"], ens["args"], ens["kwargs"])
    with hostsim_backend.use(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = pd.pydisort(*args, **kwargs)
        Fp = out[1](ens["tau_eval"][lo:hi])
    full = parallel.all_gather_columns(Fp, ens["B"])
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    # the same shard through the one-call host pipeline, with the inputs that have a per-layer description passed as such
    # (inputs.HenyeyGreenstein / LevelSource are cut like the arrays they stand for)
    from pythonic_disort_b200 import ensemble
    cargs, ckw = list(ens["args"]), dict(ens["kwargs"])
    cargs[3] = ens["compact"]["Leg_coeffs_all"]
    ckw["s_poly_coeffs"] = ens["compact"]["s_poly_coeffs"]
    a2, k2, _ = parallel.shard_inputs(ens["B"], tuple(cargs), ckw)
    with hostsim_backend.use(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = ensemble.solve_ensemble(*a2, tau=ens["tau_eval"][lo:hi], outputs=("flux_up",), chunk=4, **k2)
    np.save(os.path.join(out_dir, f"compact{rank}.npy"), parallel.all_gather_columns(res["flux_up"], ens["B"]))
    dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_single_process(tmp_path):
    import warnings

    import hostsim_backend
    import pythonic_disort_b200 as pd
    from pythonic_disort_b200 import parallel, synthetic
    assert parallel.shard_range(11, 0, 2) == (0, 6) and parallel.shard_range(11, 1, 2) == (6, 11)
    hostsim_backend.build()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ens = synthetic.make("lw", 11)
    with hostsim_backend.use(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = pd.pydisort(*ens["args"], **ens["kwargs"])[1](ens["tau_eval"])
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        assert got.shape == ref.shape
        np.testing.assert_array_equal(got, ref)
        compact = np.load(os.path.join(str(tmp_path), f"compact{r}.npy"))
        np.testing.assert_allclose(compact, ref, rtol=0, atol=1e-12 * np.max(np.abs(ref)))
