#!/usr/bin/env python
"""Benchmark of the pydisort hot path on B200 (contract: see task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload sw|sw_flux|lw|ha|tp1|tp9c] [--columns B] [--chunk C] [--scaling weak|strong]
                    [--no-cpu] [--no-others]

One "step" = the whole hot path (prologue -> eigen stage -> boundary-condition stage -> evaluation of fluxes and,
where the workload asks, NT-corrected intensities on the L+1 interface levels) over one synthetic column ensemble.
Default workload: SURVEY.md 8(d) config 3, the shortwave ensemble the headline metric is quoted on: 65,536 columns
x 60 layers, NQuad=16, NLeg_all=32, NFourier=16, delta-M + NT, per-column Lambertian albedo; weak scaling (every
rank owns its own 65,536 columns).

value : columns/s, inputs resident in HBM, outputs left in HBM (CUDA events, max over ranks).
e2e   : columns/s through the package's one-call host interface, pythonic_disort_b200.ensemble.solve_ensemble():
        pinned HOST inputs -> HOST outputs, every H2D / D2H byte inside the timed region.  Inputs are the
        reference-shaped arrays (Leg_coeffs_all [B, L, NLeg_all], s_poly_coeffs [B, L, 2], ...).
e2e_compact_inputs : the same call with the inputs that have a per-layer description passed as such
        (inputs.HenyeyGreenstein: asymmetry parameters, inputs.LevelSource: level values of the thermal source) and
        expanded on the device: what the link-bound LW ensemble reaches when the host ships 242 instead of 844 doubles
        per column.  Reported beside `e2e`, never in its place.
roofline : the dominant kernel's algorithmic FP64 FLOP/s (SURVEY 8(d) counts) against the FP64 FMA peak measured in
        this run by pd_fp64_probe (MEASURED_PEAKS.json has no FP64 entry); peak_dfma / peak_dmma are both given;
        `hbm` is the same kernel against MEASURED_PEAKS.json:hbm_gbs; `traffic` is the ncu DRAM byte count of that
        kernel per launch from profiles/r2_traffic.json (made by tools/traffic_from_ncu.py from a capture of this build).
cpu_baseline : the oracle port (oracle/disort_oracle.py) on the host cores.
other_workloads : the remaining BASELINE.json configs, short runs: LW (1,048,576 columns) and HA (16,384 columns) with
        the ensemble FIXED and sharded over the N ranks (strong scaling, configs 4-5), SW flux-only, TP1 x 4096 and
        TP9c x 4096 (configs 1-2, replicated per rank).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

WORKLOADS = {
    "sw": dict(ens="sw", columns=65536, metric="columns/s (60 layers, NQuad=16)",
               desc="shortwave ensemble: 60 layers, NQuad=16, NLeg_all=32, NFourier=16, "
               "delta-M + NT, Lambertian; fluxes + NT-corrected intensities at 61 levels x 16 mu x 3 phi"),
    "sw_flux": dict(ens="sw", columns=65536, only_flux=True,
                    desc="shortwave ensemble, only_flux=True; fluxes at 61 levels"),
    "lw": dict(ens="lw", columns=1048576, desc="longwave ensemble: 60 layers, NQuad=8, thermal source, flux-only"),
    "ha": dict(ens="ha", columns=16384, desc="high-accuracy: 100 layers, NQuad=32, NFourier=32, Hapke BDRF, "
               "intensities at 101 levels x 6 user polar angles (mu = +-0.1, +-0.5, +-0.9; interpolate() on the device) x 5 phi"),
    "tp1": dict(ens="tp1", columns=6 * 4096, desc="pydisotest test problem 1a-1f (1 layer, NQuad=16, isotropic "
                "scattering, beam) x 4096; fluxes + intensities at 5 levels x 3 phi"),
    "tp9c": dict(ens="tp9c", columns=4096, desc="pydisotest test problem 9c (6 layers, NQuad=8, beam + thermal + "
                 "Lambertian + Dirichlet) x 4096; fluxes + intensities at 7 levels x 3 phi"),
}
SHAPES = {"sw": (60, 16), "lw": (60, 8), "ha": (100, 32), "tp1": (1, 16), "tp9c": (6, 8)}  # L, NQuad


# ---------------------------------------------------------------------------
# algorithmic work: FLOPs per column by stage, bytes per column
# ---------------------------------------------------------------------------
def algorithmic_flops(L, NQuad, NLeg, NF, beam, thermal, nlev, nphi):
    N = NQuad // 2
    eig = part = setup = 0.0
    for m in range(NF):
        nm = NLeg - m
        setup += L * (4 * N * N * nm + (4 * N * nm if beam else 0) + 6 * N * N)
        eig += L * (2 * N**3 + 25 * N**3 + 2 * N**3 + 4 * N * N)
        if beam:
            part += L * ((2.0 / 3.0) * (2 * N) ** 3 + 2 * (2 * N) ** 2)
        if thermal and m == 0:
            part += L * (2 * (2 * N) ** 3 + 4 * (2 * N) ** 2)
    bc = NF * (2.0 * (2 * N * L) * (3 * N - 1) * (3 * N) + 6.0 * (2 * N * L) * (3 * N - 1) + 4.0 * N * N * L)
    ev = nlev * NF * (2 * (2 * N) ** 2 + 44 * N + 4 * N * nphi)
    return dict(eigen_stage=setup + eig + part, bc_stage=bc, eval=ev, total=setup + eig + part + bc + ev)


L2_POLICY = ("inputs and solved state of a step are far larger than L2 (>= 1 GB per chunk except tp1/tp9c, whose "
             "4096-column state is 50-90 MB against 126 MB of L2: those two are L2-warm numbers); no explicit flush")


def workload_config(workload, cols_total, per_gpu, chunk, chunk_e2e):
    """The `config` object of a bench line: identical for the GPU arm and the `--impl reference` arm of one workload."""
    return {"workload": workload, "description": WORKLOADS[workload]["desc"], "columns_total": cols_total,
            "columns_per_gpu": per_gpu, "chunk_columns": chunk, "chunk_columns_e2e": chunk_e2e,
            "seed": "pythonic_disort_b200/synthetic.py", "l2_policy": L2_POLICY}


def chunk_sizes(workload, per_gpu, chunk=0):
    """(columns per pydisort() call of the device-resident arm, of the host pipeline): the first as many as the solved
    state allows, the second at least six chunks per ensemble (ensemble.default_chunk); --chunk overrides both."""
    from pythonic_disort_b200 import ensemble
    wl = WORKLOADS[workload]
    L, NQuad = SHAPES[wl["ens"]]
    NF = 1 if wl.get("only_flux") or wl["ens"] == "lw" else NQuad
    if chunk > 0:
        return min(chunk, per_gpu), min(chunk, per_gpu)
    return (min(ensemble.default_chunk(per_gpu, L, NQuad, NF, pipeline=False), per_gpu),
            min(ensemble.default_chunk(per_gpu, L, NQuad, NF), per_gpu))


def make_inputs(name, ncol, first, only_flux=False):
    from pythonic_disort_b200 import synthetic
    ens = synthetic.make(name, ncol, first)
    if only_flux:
        ens["kwargs"]["only_flux"] = True
        ens["kwargs"].pop("NT_cor", None)
        ens["outputs"] = ("flux",)
    return ens


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (also `--impl reference`)
# ---------------------------------------------------------------------------
def _cpu_worker(job):
    name, first, ncol, only_flux = job
    warnings.simplefilter("ignore")
    from oracle import disort_oracle
    from pythonic_disort_b200 import synthetic
    ens = make_inputs(name, ncol, first, only_flux)
    synthetic.run_reference_like(disort_oracle.pydisort, ens, at_user_mu=True)
    return ncol


def cpu_columns_per_second(wl, seconds_target=15.0, cores=None, pool=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count()
    name, only_flux = wl["ens"], wl.get("only_flux", False)
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
    try:
        pool.map(_cpu_worker, [(name, i, 1, only_flux) for i in range(cores)])      # warm-up: imports + 1 column
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(name, 100 + i, 1, only_flux) for i in range(cores)])
        per_col = max(time.perf_counter() - t0, 1e-4)
        per_worker = int(min(max(2, seconds_target / per_col), 4096))
        t0 = time.perf_counter()
        done = sum(pool.map(_cpu_worker, [(name, 1000 + i * per_worker, per_worker, only_flux) for i in range(cores)]))
        dt = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    return done / dt, cores, f"{done} columns of the same ensemble ({per_worker} per worker process), {dt:.1f} s"


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
# kernels launched per C-ABI call (label of the event mark api.py records around it)
def _launches(label, N, nt):
    return {"prologue": 1, "solve_eigen": 2 if N in (4, 8, 16) else 1, "solve_bc": 3 if N == 8 else 2 if N in (2, 4, 16) else 1,
            "eval_flux": 1, "eval_u0": 1, "eval_u": 2 if nt else 1, "interp_mu": 1}.get(label, 1)


class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        from pythonic_disort_b200 import parallel
        self.cores = parallel.bind_to_gpu_numa(self.local) if self.world > 1 else None
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        self.traffic = {}
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        except (OSError, ValueError):
            pass
        self.fp64_peak = self.measure_fp64_peak()

    def measure_fp64_peak(self):
        """FP64 FMA peak of this GPU (independent DFMA chains, CUDA events), TFLOP/s."""
        torch = self.torch
        from pythonic_disort_b200 import _lib
        lib = _lib.cuda_lib()
        sink = torch.zeros(8, dtype=torch.float64, device=self.dev)
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        lib.pd_fp64_probe(sink.data_ptr(), 1000, stream)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            flops = lib.pd_fp64_probe(sink.data_ptr(), 20000, stream)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, flops / (e0.elapsed_time(e1) * 1e-3))
        return best / 1e12

    def timed(self, nsteps, fn, finish=None):
        torch, dist = self.torch, self.dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(nsteps):
            fn(k)
        if finish:
            finish()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def run(self, workload, steps, warmup, strong, columns=0, chunk=0, cpu_seconds=15.0, do_cpu=True, pool=None):
        """One workload on this rank's columns; returns the result dict (rank 0 fills the CPU baseline)."""
        torch = self.torch
        import pythonic_disort_b200 as pd
        from pythonic_disort_b200 import api, ensemble, parallel
        wl = WORKLOADS[workload]
        only_flux = wl.get("only_flux", False)
        total = columns or wl["columns"]
        if strong:   # a fixed ensemble of `total` columns, contiguous shards (parallel.shard_range)
            lo, hi = parallel.shard_range(total, self.rank, self.world)
            if total <= 65536:
                full = make_inputs(wl["ens"], total, 0, only_flux)
                a, k, _ = parallel.shard_inputs(total, full["args"], full["kwargs"], self.rank, self.world)
                ens = dict(full, args=a, kwargs=k, B=hi - lo, tau_eval=full["tau_eval"][lo:hi],
                           compact={n: v[lo:hi] for n, v in full.get("compact", {}).items()})
            else:    # too big to build on every rank: the generator's subset property gives the same columns
                ens = make_inputs(wl["ens"], hi - lo, lo, only_flux)
            cols_total = total
        else:        # weak: every rank owns its own `total` columns of the (unbounded) seeded ensemble
            ens = make_inputs(wl["ens"], total, self.rank * total, only_flux)
            cols_total = total * self.world
        B = ens["B"]
        L, NQuad = SHAPES[wl["ens"]]
        N = NQuad // 2
        NF = 1 if (only_flux or ens["kwargs"].get("only_flux")) else NQuad
        want_u = "u" in ens["outputs"]
        chunk, chunk_e2e = chunk_sizes(workload, B, chunk)
        phi = ens["phi_eval"] if want_u else None
        mu_user = ens.get("mu_user") if want_u else None
        outputs = ("flux_up", "flux_down") + (("u",) if want_u else ())

        # ---- device-resident arm: inputs in HBM, outputs left in HBM ----
        def as_dev(x):
            return torch.as_tensor(x, device=self.dev) if isinstance(x, np.ndarray) else x

        dev_args = [as_dev(a) for a in ens["args"]]
        dev_kw = {k: ([as_dev(m) for m in v] if k == "BDRF_Fourier_modes" else as_dev(v)) for k, v in ens["kwargs"].items()}
        dev_tau = as_dev(ens["tau_eval"])
        phi_dev = torch.as_tensor(phi, device=self.dev) if phi is not None else None
        pending = []

        def step_dev(_k):
            for lo in range(0, B, chunk):
                hi = min(B, lo + chunk)
                ca, ck = api.slice_columns(dev_args, dev_kw, lo, hi, B)
                te = dev_tau[lo:hi]
                out = api.pydisort(*ca, **ck, _defer=pending)
                out[1](te)
                out[2](te)
                if want_u and mu_user is not None:
                    pd.subroutines.interpolate(out[4])(mu_user, te, phi_dev)
                elif want_u:
                    out[4](te, phi_dev)
                del out

        for k in range(warmup):
            step_dev(k)
        torch.cuda.synchronize()
        api._raise_for_pending(pending)
        pending.clear()
        sampler = ClockSampler(self.local)
        sampler.start()
        api._profile = []
        ms_dev = self.timed(steps, step_dev)
        marks, api._profile = api._profile, None
        clocks = sampler.stop()
        pending.clear()
        kernel_ms, launches = {}, 0
        nt = bool(ens["kwargs"].get("NT_cor")) and want_u
        for (l0, ev0), (l1, ev1) in zip(marks[:-1], marks[1:]):
            if l1 != "begin":
                kernel_ms[l1] = kernel_ms.get(l1, 0.0) + ev0.elapsed_time(ev1)
                launches += _launches(l1, N, nt)
        del dev_args, dev_kw, dev_tau
        torch.cuda.empty_cache()

        # ---- end-to-end arm: the package's host interface (pinned host inputs -> host outputs) ----
        def pin(x):
            if isinstance(x, api.DeviceInput):
                return x.with_parts(tuple(pin(p) for p in x.parts()))
            return ensemble.pinned(x) if isinstance(x, np.ndarray) else x

        host_tau = ensemble.pinned(ens["tau_eval"])

        def e2e_arm(args_in, kw_in):
            host_args = [pin(a) for a in args_in]
            host_kw = {k: ([pin(m) for m in v] if k == "BDRF_Fourier_modes" else pin(v)) for k, v in kw_in.items()}
            results = [None, None]   # two sets of pinned output buffers: step k+1 is enqueued while step k drains

            def step_e2e(k):
                cur = ensemble.solve_ensemble(*host_args, tau=host_tau, phi=phi, mu=mu_user, outputs=outputs,
                                              chunk=chunk_e2e, out=results[k % 2], wait=False, **host_kw)
                prev = results[(k + 1) % 2]
                if prev is not None:
                    prev.wait()   # step k - 1 has reached the host (its deferred checks are raised here)
                results[k % 2] = cur

            def drain():
                for r in results:
                    if r is not None:
                        r.wait()

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                step_e2e(0)
                step_e2e(1)
                drain()
                ms = self.timed(steps, step_e2e, drain)
            return ms, results[0].h2d_bytes, results[0].d2h_bytes

        ms_e2e, h2d, d2h = e2e_arm(ens["args"], ens["kwargs"])
        # the same ensemble with the inputs that have one described per layer (Henyey-Greenstein asymmetry parameters,
        # level values of the thermal source) instead of as arrays: pydisort() expands them on the device (inputs.py)
        e2e_compact = None
        if ens.get("compact"):
            cargs, ckw = list(ens["args"]), dict(ens["kwargs"])
            for name, obj in ens["compact"].items():
                if name in api.POSITIONAL:
                    cargs[api.POSITIONAL.index(name)] = obj
                else:
                    ckw[name] = obj
            ms_c, h2d_c, d2h_c = e2e_arm(cargs, ckw)
            e2e_compact = {"value": cols_total / (ms_c / steps * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": h2d_c,
                           "d2h_bytes_per_step": d2h_c, "inputs": {n: type(o).__name__ for n, o in ens["compact"].items()},
                           "api": "pythonic_disort_b200.ensemble.solve_ensemble"}

        ms_step = ms_dev / steps
        value = cols_total / (ms_step * 1e-3)
        e2e_value = cols_total / (ms_e2e / steps * 1e-3)
        beam = wl["ens"] != "lw"
        nphi = len(phi) if phi is not None else 0
        fl = algorithmic_flops(L, NQuad, NQuad, NF, beam, wl["ens"] in ("lw", "tp9c"), ens["tau_eval"].shape[1], nphi)
        per_kernel = {k: v / steps for k, v in kernel_ms.items()}
        stage_flops = {"solve_eigen": fl["eigen_stage"], "solve_bc": fl["bc_stage"]}
        dom = max(stage_flops, key=lambda k: per_kernel.get(k, 0.0))
        dom_ms = per_kernel[dom]
        achieved = stage_flops[dom] * B / (dom_ms * 1e-3) / 1e12
        hbm_peak = self.peaks.get("hbm_gbs", 6650.0)
        item = NF * L
        bytes_k = {"solve_eigen": item * (NQuad + 1) * 8 + item * (2 * N * N + N + 2 * N) * 8,
                   "solve_bc": item * (2 * N * N + N + 2 * N) * 8 + item * 2 * N * 8}[dom]
        kname = {"solve_eigen": "k_stage_a_sym" if N in (4, 8) else "k_stage_a_j16" if N == 16 else "k_stage_a",
                 "solve_bc": "k_stage_b_tps" if N in (2, 4) else "k_layer_ops + k_stage_b_add" if N == 8 else
                 "k_stage_b_add" if N == 16 else "k_stage_b"}[dom]   # the kernels the stage's event time covers
        parts = [self.traffic.get(workload, {}).get(k.strip()) for k in kname.split("+")]
        per_col = sum(parts) if all(parts) else None
        peak = self.fp64_peak
        roofline = {"kernel": kname, "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak if peak > 0 else None,
                    "peak_dfma": peak, "peak_dmma": 37.1,
                    "peak_source": "peak / peak_dfma: pd_fp64_probe DFMA micro-benchmark measured in this run "
                    "(MEASURED_PEAKS.json has no FP64 entry); peak_dmma: tools/ubench/dmma_rate on this pool "
                    "(profiles/r1_dmma_ubench.txt) -- no kernel of this build issues DMMA",
                    "traffic": per_col * min(chunk, B) if per_col else None,
                    "traffic_bytes_per_column": per_col,
                    "traffic_source": "profiles/r2_traffic.json (ncu --set full of this build, dram__bytes_read.sum + "
                    "dram__bytes_write.sum per column; one launch = one chunk)" if per_col else None,
                    "hbm": {"achieved_gbs": bytes_k * B / (dom_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if self.peaks else "fallback",
                            "algorithmic_bytes_per_column": bytes_k},
                    "algorithmic_flops_per_column": stage_flops[dom], "kernel_ms_per_step": dom_ms,
                    "whole_path": {"algorithmic_mflop_per_column": fl["total"] / 1e6,
                                   "achieved_tflops_per_gpu": fl["total"] * B / (ms_step * 1e-3) / 1e12,
                                   "frac_of_fp64_peak": fl["total"] * B / (ms_step * 1e-3) / 1e12 / peak if peak > 0 else None},
                    "kernel_ms_per_step_all": per_kernel}
        cpu = {"value": None, "unit": "columns/s", "cores": None, "kind": "port", "sample": "skipped (N>1 or --no-cpu)"}
        if do_cpu and self.world == 1 and self.rank == 0:
            v, cores, sample = cpu_columns_per_second(wl, seconds_target=cpu_seconds, pool=pool)
            cpu = {"value": v, "unit": "columns/s", "cores": cores, "kind": "port", "sample": sample}
        return {
            "metric": wl.get("metric", f"columns/s ({workload})"), "value": value, "unit": "columns/s",
            "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "scaling": "strong" if strong else "weak",
            "config": workload_config(workload, cols_total, B, chunk, chunk_e2e),
            "e2e": {"value": e2e_value, "unit": "columns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "pythonic_disort_b200.ensemble.solve_ensemble"},
            "e2e_compact_inputs": e2e_compact,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }


def run_gpu(args):
    warnings.simplefilter("ignore")  # the ensembles trip the reference's "close to 1" warnings by design
    bench = Bench(args)
    strong = args.scaling == "strong"
    pool = None
    if bench.world == 1 and not args.no_cpu:
        import multiprocessing as mp
        pool = mp.get_context("spawn").Pool(os.cpu_count())
    main = bench.run(args.workload, args.steps, args.warmup, strong, args.columns, args.chunk, args.cpu_seconds,
                     not args.no_cpu, pool)
    others = {}
    if args.workload == "sw" and not args.no_others and not args.columns:
        plan = [("lw", True, 3, 3), ("ha", True, 1, 3), ("sw_flux", False, 3, 3), ("tp1", False, 3, 3), ("tp9c", False, 3, 3)]
        for name, strong_o, st, wu in plan:
            r = bench.run(name, st, wu, strong_o, 0, 0, min(args.cpu_seconds, 5.0), not args.no_cpu, pool)
            keep = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "e2e", "e2e_compact_inputs", "cpu_baseline",
                                      "gpu_launches")}
            keep["config"] = {k: r["config"][k] for k in ("description", "columns_total", "columns_per_gpu", "chunk_columns",
                                                           "chunk_columns_e2e")}
            rf = r["roofline"]
            keep["roofline"] = {k: rf[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic",
                                                   "traffic_bytes_per_column", "kernel_ms_per_step_all")}
            keep["roofline"]["whole_path_frac_of_fp64_peak"] = rf["whole_path"]["frac_of_fp64_peak"]
            others[name] = keep
    if pool is not None:
        pool.close()
        pool.join()
    if bench.rank == 0:
        line = dict(main)
        line.update({"n_gpus": bench.world, "higher_is_better": True, "vs_baseline": None, "dtype": "f64",
                     "data": "synthetic"})
        if others:
            line["other_workloads"] = others
        if bench.cores:
            line["cpu_affinity"] = f"{len(bench.cores)} cores local to the GPU (NVML)"
        print(json.dumps(line))
    if bench.world > 1:
        bench.dist.destroy_process_group()


def run_reference(args):
    """`--impl reference`: the reference algorithm's CPU port (oracle) on all host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = WORKLOADS[args.workload]
    vals = []
    cores = sample = None
    for _ in range(max(1, args.warmup and 1) + max(1, min(args.steps, 3))):
        v, cores, sample = cpu_columns_per_second(wl, seconds_target=args.cpu_seconds)
        vals.append(v)
    value = float(np.median(vals[1:])) if len(vals) > 1 else vals[0]
    total = args.columns or wl["columns"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    strong = args.scaling == "strong"
    per_gpu = -(-total // world) if strong else total   # what the GPU arm of the same command gives every rank
    line = {
        "impl": "reference", "metric": wl.get("metric", f"columns/s ({args.workload})"),
        "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, total * (1 if strong else world), per_gpu, *chunk_sizes(args.workload, per_gpu, args.chunk)),
        "cpu_baseline": {"value": value, "unit": "columns/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sw", choices=sorted(WORKLOADS))
    ap.add_argument("--columns", type=int, default=0, help="columns (per GPU when weak, in total when strong; default: the workload's full size)")
    ap.add_argument("--chunk", type=int, default=0, help="columns per pydisort() call (0 = sized from the workload)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other BASELINE workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
