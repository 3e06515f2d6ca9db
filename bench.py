#!/usr/bin/env python
"""Benchmark of the pydisort hot path on B200 (contract: see task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload sw|sw_flux|lw|ha] [--columns B] [--chunk C]

One "step" = the whole hot path (prologue -> eigen stage -> boundary-condition
stage -> evaluation of fluxes and, where the workload asks, NT-corrected
intensities on the L+1 interface levels) over one synthetic column ensemble.
Default workload: SURVEY.md 8(d) config 3, the shortwave ensemble the headline
metric is quoted on: 65,536 columns x 60 layers, NQuad=16, NLeg_all=32,
NFourier=16, delta-M + NT, per-column Lambertian albedo.

value : columns/s, inputs resident in HBM, outputs left in HBM (CUDA events).
e2e   : columns/s through pydisort() with pinned HOST inputs and HOST outputs
        (H2D and D2H copies inside the timed region).
roofline : the dominant kernel's algorithmic FP64 FLOP/s against the FP64 FMA
        peak measured live on this GPU by pd_fp64_probe (MEASURED_PEAKS.json
        has no FP64 entry), plus the HBM view against MEASURED_PEAKS.json.
cpu_baseline : the oracle port (oracle/disort_oracle.py) on the host cores.
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
    "sw": dict(ens="sw", columns=65536, desc="shortwave ensemble: 60 layers, NQuad=16, NLeg_all=32, NFourier=16, "
               "delta-M + NT, Lambertian; fluxes + NT-corrected intensities at 61 levels x 16 mu x 3 phi"),
    "sw_flux": dict(ens="sw", columns=65536, only_flux=True,
                    desc="shortwave ensemble, only_flux=True; fluxes at 61 levels"),
    "lw": dict(ens="lw", columns=1048576, desc="longwave ensemble: 60 layers, NQuad=8, thermal source, flux-only"),
    "ha": dict(ens="ha", columns=16384, desc="high-accuracy: 100 layers, NQuad=32, NFourier=32, Hapke BDRF, "
               "intensities at 101 levels x 6 user polar angles (mu = +-0.1, +-0.5, +-0.9; interpolate() on the device) x 5 phi"),
}


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


def cpu_columns_per_second(wl, seconds_target=15.0, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count()
    name, only_flux = wl["ens"], wl.get("only_flux", False)
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, [(name, i, 1, only_flux) for i in range(cores)])      # warm-up: imports + 1 column
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(name, 100 + i, 1, only_flux) for i in range(cores)])
        per_col = max(time.perf_counter() - t0, 1e-4)
        per_worker = int(min(max(2, seconds_target / per_col), 4096))
        t0 = time.perf_counter()
        done = sum(pool.map(_cpu_worker, [(name, 1000 + i * per_worker, per_worker, only_flux) for i in range(cores)]))
        dt = time.perf_counter() - t0
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
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import pythonic_disort_b200 as pd
    from pythonic_disort_b200 import _lib, api

    warnings.simplefilter("ignore")  # the ensembles trip the reference's "close to 1" warnings by design
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    ncol = args.columns or wl["columns"]
    only_flux = wl.get("only_flux", False)
    if args.chunk > 0:
        chunk = min(args.chunk, ncol)
    else:  # columns per pydisort() call: keep the solved state (K, G, Bv, C) of one call under ~32 GB
        shape = {"sw": (60, 8, 16), "lw": (60, 4, 1), "ha": (100, 16, 32)}[wl["ens"]]
        nf = 1 if only_flux else shape[2]
        per_col = nf * shape[0] * (2 * shape[1] ** 2 + 5 * shape[1]) * 8
        chunk = max(1024, min(ncol, 65536, int(32e9 / per_col) // 1024 * 1024))
        if wl["ens"] == "sw" and not only_flux:
            chunk = min(chunk, 16384)

    # weak scaling: every rank owns its own `ncol` columns of the (unbounded) seeded ensemble
    ens = make_inputs(wl["ens"], ncol, rank * ncol, only_flux)
    want_u = "u" in ens["outputs"]
    B = ens["B"]

    def split(x, lo, hi):
        return x[lo:hi] if isinstance(x, (np.ndarray, torch.Tensor)) and x.ndim >= 1 and x.shape[0] == B else x

    def as_host(x):
        return torch.from_numpy(np.ascontiguousarray(x)).pin_memory() if isinstance(x, np.ndarray) else x

    def as_dev(x):
        return torch.as_tensor(x, device=dev) if isinstance(x, np.ndarray) else x

    host_args = [as_host(a) for a in ens["args"]]
    host_kw = {k: ([as_host(m) for m in v] if k == "BDRF_Fourier_modes" else as_host(v)) for k, v in ens["kwargs"].items()}
    host_tau = as_host(ens["tau_eval"])
    dev_args = [as_dev(a) for a in ens["args"]]
    dev_kw = {k: ([as_dev(m) for m in v] if k == "BDRF_Fourier_modes" else as_dev(v)) for k, v in ens["kwargs"].items()}
    dev_tau = as_dev(ens["tau_eval"])
    phi = ens["phi_eval"]
    phi_dev = torch.as_tensor(phi, device=dev) if phi is not None else None
    mu_user = ens.get("mu_user")  # config 5: intensities at user polar angles (row f1)

    def one_chunk(a, kw, tau_eval, to_host, lo):
        """One pydisort() call + evaluation for columns [lo, lo + chunk); returns bytes copied (h2d, d2h)."""
        h2d = d2h = 0
        hi = min(B, lo + chunk)
        ca = [split(x, lo, hi) for x in a]
        ck = {k: ([split(m, lo, hi) for m in v] if k == "BDRF_Fourier_modes" else split(v, lo, hi))
              for k, v in kw.items()}
        te = tau_eval[lo:hi]
        if to_host:  # public API on (pinned) host buffers: pydisort copies in, the output functions copy out
            h2d += sum(x.numel() * 8 for x in ca if isinstance(x, torch.Tensor)) + te.numel() * 8
            h2d += sum(v.numel() * 8 for v in ck.values() if isinstance(v, torch.Tensor))
            h2d += sum(m.numel() * 8 for m in ck.get("BDRF_Fourier_modes", []) if isinstance(m, torch.Tensor))
        out = pd.pydisort(*ca, **ck)
        Fp = out[1](te)
        Fm, Fd = out[2](te)
        if want_u and mu_user is not None:
            uu = pd.subroutines.interpolate(out[4])(mu_user, te, phi if to_host else phi_dev)
        else:
            uu = out[4](te, phi if to_host else phi_dev) if want_u else None
        if to_host:  # host inputs -> the API returned NumPy arrays (device->host copies already done)
            assert isinstance(Fp, np.ndarray)
            d2h += (Fp.size + Fm.size + Fd.size + (uu.size if want_u else 0)) * 8
        del out
        return h2d, d2h

    # End-to-end arm: a three-stage pipeline over the chunks, the way a caller with host data drives the public API:
    # pinned host inputs -> device on a copy stream (one chunk ahead), pydisort() + output functions on device tensors
    # on the compute stream, results -> pinned host buffers on a second copy stream.  The copies of one chunk overlap
    # the kernels of its neighbours; every byte moved is inside the timed region.  (Two host threads on two streams do
    # not achieve this: their kernels interleave and both reach the copy phase at the same time -- tools/e2e_overlap.py.)
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    host_out = {}  # (slot, name) -> pinned buffer, slot = chunk parity
    d2h_done = {}  # running chunk number -> event of its device->host copies (buffers are reused every other chunk)
    chunk_no = [0]

    def to_device_async(x, cur):
        if not isinstance(x, torch.Tensor):
            return x
        d = x.to(dev, non_blocking=True)
        d.record_stream(cur)
        return d

    carry = {}          # chunk 0 of the next step, staged while the last chunk of this one computes
    steps_left = [0]    # set by the caller of `timed`: how many more end-to-end steps follow the current one

    def step_e2e(a, kw, tau_eval):
        cur = torch.cuda.current_stream(dev)
        starts = list(range(0, B, chunk))
        h2d = d2h = 0
        staged = {}
        steps_left[0] -= 1

        def stage(i):
            lo, hi = starts[i], min(B, starts[i] + chunk)
            with torch.cuda.stream(h2d_stream):
                ca = [to_device_async(split(x, lo, hi), cur) for x in a]
                ck = {k: ([to_device_async(split(m, lo, hi), cur) for m in v] if k == "BDRF_Fourier_modes"
                          else to_device_async(split(v, lo, hi), cur)) for k, v in kw.items()}
                te = to_device_async(tau_eval[lo:hi], cur)
                ev = torch.cuda.Event()
                ev.record(h2d_stream)
            nbytes = sum(x.numel() * 8 for x in ca if isinstance(x, torch.Tensor)) + te.numel() * 8
            nbytes += sum(v.numel() * 8 for v in ck.values() if isinstance(v, torch.Tensor))
            nbytes += sum(m.numel() * 8 for m in ck.get("BDRF_Fourier_modes", []) if isinstance(m, torch.Tensor))
            staged[i] = (ca, ck, te, ev, nbytes)

        if 0 in carry:
            staged[0] = carry.pop(0)
        else:
            stage(0)
        for i in range(len(starts)):
            ca, ck, te, ev, nbytes = staged.pop(i)
            if i + 1 < len(starts):
                stage(i + 1)
            elif steps_left[0] > 0:  # the next step's first chunk rides behind this step's last one
                stage(0)
                carry[0] = staged.pop(0)
            h2d += nbytes
            cur.wait_event(ev)
            out = pd.pydisort(*ca, **ck)
            res = {"Fp": out[1](te)}
            res["Fm"], res["Fd"] = out[2](te)
            if want_u and mu_user is not None:
                res["u"] = pd.subroutines.interpolate(out[4])(mu_user, te, phi_dev)
            elif want_u:
                res["u"] = out[4](te, phi_dev)
            evc = torch.cuda.Event()
            evc.record(cur)
            g = chunk_no[0]
            chunk_no[0] += 1
            if g - 2 in d2h_done:  # the pinned buffers of this parity are free once their previous copy has landed
                d2h_done.pop(g - 2).synchronize()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(evc)
                for name, t in res.items():
                    key = (g % 2, name)
                    if key not in host_out or host_out[key].shape != t.shape:
                        host_out[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    t.record_stream(d2h_stream)
                    host_out[key].copy_(t, non_blocking=True)
                    d2h += t.numel() * 8
                evd = torch.cuda.Event()
                evd.record(d2h_stream)
            d2h_done[g] = evd
            del out, res
        # no wait here: the copies of the last chunk overlap the first chunk of the next step; `timed` closes the
        # timed region only after the copy stream has drained
        return h2d, d2h

    def step(a, kw, tau_eval, to_host):
        """The hot path over all columns, chunk by chunk; returns bytes copied (h2d, d2h)."""
        if to_host:
            return step_e2e(a, kw, tau_eval)
        res = [one_chunk(a, kw, tau_eval, False, lo) for lo in range(0, B, chunk)]
        return sum(r[0] for r in res), sum(r[1] for r in res)

    def timed(nsteps, fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        extra = None
        for _ in range(nsteps):
            extra = fn()
        torch.cuda.current_stream(dev).wait_stream(d2h_stream)  # results of the last step have reached the host
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, extra

    # ---- FP64 peak of this GPU (DFMA chains, CUDA events) ----
    lib = _lib.cuda_lib()
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
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
    fp64_peak_tflops = best / 1e12

    # ---- warm-up ----
    for _ in range(args.warmup):
        step(dev_args, dev_kw, dev_tau, False)
    torch.cuda.synchronize()

    # ---- device-resident timing, with per-kernel CUDA-event marks ----
    sampler = ClockSampler(local)
    sampler.start()
    api._profile = []
    ms_dev, _ = timed(args.steps, lambda: step(dev_args, dev_kw, dev_tau, False))
    marks, api._profile = api._profile, None
    clocks = sampler.stop()
    kernel_ms = {}
    for (l0, ev0), (l1, ev1) in zip(marks[:-1], marks[1:]):
        if l1 != "begin":
            kernel_ms[l1] = kernel_ms.get(l1, 0.0) + ev0.elapsed_time(ev1)
    launches_per_step = sum(1 for lab, _ in marks if lab != "begin") // max(args.steps, 1)
    if want_u and not only_flux:
        launches_per_step += (B + chunk - 1) // chunk  # pd_eval_u launches two kernels when NT is on

    # ---- end-to-end timing through the public API on host buffers ----
    step(host_args, host_kw, host_tau, True)  # warm pinned paths
    steps_left[0] = args.steps
    ms_e2e, (h2d, d2h) = timed(args.steps, lambda: step(host_args, host_kw, host_tau, True))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cols_total = B * world
    ms_step = ms_dev / args.steps
    value = cols_total / (ms_step * 1e-3)
    e2e_value = cols_total / (ms_e2e / args.steps * 1e-3)

    cfgL, NQuad = ens["L"], ens["NQuad"]
    NLeg = NQuad
    NF = 1 if only_flux or ens["name"] == "lw" else NQuad
    beam = ens["name"] != "lw"
    nphi = len(phi) if (phi is not None and want_u) else 0
    fl = algorithmic_flops(cfgL, NQuad, NLeg, NF, beam, ens["name"] == "lw", cfgL + 1, nphi)
    per_kernel = {k: v / args.steps for k, v in kernel_ms.items()}
    stage_flops = {"solve_eigen": fl["eigen_stage"], "solve_bc": fl["bc_stage"]}
    dom = max(stage_flops, key=lambda k: per_kernel.get(k, 0.0))
    dom_ms = per_kernel[dom]
    achieved = stage_flops[dom] * B / (dom_ms * 1e-3) / 1e12
    N = NQuad // 2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # algorithmic bytes of the dominant kernel per column: what it must read and write once
    item = NF * cfgL
    bytes_k = {"solve_eigen": item * (NLeg + 1) * 8 + item * (2 * N * N + N + 2 * N) * 8,
               "solve_bc": item * (2 * N * N + N + 2 * N) * 8 + item * 2 * N * 8}[dom]
    kname = {"solve_eigen": "k_stage_a_sym" if N in (4, 8) else "k_stage_a",
             "solve_bc": "k_stage_b_add" if N in (2, 4, 8, 16) else "k_stage_b"}[dom]
    # DRAM bytes per column of that kernel from the committed ncu capture (profiles/r1_traffic.json), if any
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        per_col = tr.get(args.workload, {}).get(kname)
        traffic = per_col * B if per_col else None
    except (OSError, ValueError):
        pass
    roofline = {"kernel": kname, "bound": "fp64",
                "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak_tflops if fp64_peak_tflops > 0 else None,
                "peak_source": "pd_fp64_probe DFMA micro-benchmark, measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "traffic": traffic,
                "hbm": {"achieved_gbs": bytes_k * B / (dom_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "algorithmic_bytes_per_column": bytes_k},
                "algorithmic_flops_per_column": stage_flops[dom], "kernel_ms_per_step": dom_ms,
                "whole_path": {"algorithmic_mflop_per_column": fl["total"] / 1e6,
                               "achieved_tflops": fl["total"] * B / (ms_step * 1e-3) / 1e12 / world * world,
                               "frac_of_fp64_peak": fl["total"] * cols_total / (ms_step * 1e-3) / 1e12 / (fp64_peak_tflops * world)},
                "kernel_ms_per_step_all": per_kernel}

    cpu_val, cores, sample = cpu_columns_per_second(wl, seconds_target=args.cpu_seconds) if world == 1 and not args.no_cpu \
        else (None, None, "skipped (N>1 or --no-cpu)")
    line = {
        "metric": "columns/s (60 layers, NQuad=16)" if ens["name"] == "sw" else f"columns/s ({args.workload})",
        "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"], "columns_per_gpu": B, "chunk_columns": chunk,
                   "seed": "pythonic_disort_b200/synthetic.py", "l2_policy": "inputs and state per step are far larger than L2 "
                   "(>= 1 GB per chunk); no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "columns/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_val, "unit": "columns/s", "cores": cores, "kind": "port", "sample": sample},
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    line = {
        "impl": "reference", "metric": "columns/s (60 layers, NQuad=16)" if wl["ens"] == "sw" else f"columns/s ({args.workload})",
        "value": value, "unit": "columns/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": wl["desc"]},
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
    ap.add_argument("--columns", type=int, default=0, help="columns per GPU (default: the workload's full size)")
    ap.add_argument("--chunk", type=int, default=0, help="columns per pydisort() call (0 = sized from the workload)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
