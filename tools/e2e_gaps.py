#!/usr/bin/env python
"""Where the end-to-end pipeline (ensemble.solve_ensemble) spends a step: time inside the library's launches on the
compute stream against the gaps between them (waits for an upload, allocator work, launch latency), and the host
time needed to enqueue a step.  python tools/e2e_gaps.py [workload] [columns] [chunk]"""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pythonic_disort_b200 import api, ensemble, parallel, synthetic  # noqa: E402

warnings.simplefilter("ignore")
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
if world > 1:   # under torchrun: one rank per GPU, all ranks run the same pipeline at the same time (no communication)
    parallel.bind_to_gpu_numa(local)
torch.cuda.set_device(local)
name = sys.argv[1] if len(sys.argv) > 1 else "sw"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
ens = synthetic.make(name, B)
if os.environ.get("PD_ONLY_FLUX"):   # the flux-only variant of a workload (bench.py: sw_flux)
    ens["kwargs"]["only_flux"] = True
    ens["kwargs"].pop("NT_cor", None)
    ens["outputs"] = ("flux",)
pin = lambda x: ensemble.pinned(x) if hasattr(x, "shape") and getattr(x, "ndim", 0) > 0 else x  # noqa: E731
args = [pin(a) for a in ens["args"]]
kw = {k: ([pin(m) for m in v] if k == "BDRF_Fourier_modes" else pin(v)) for k, v in ens["kwargs"].items()}
tau = pin(ens["tau_eval"])
outputs = ("flux_up", "flux_down") + (("u",) if "u" in ens["outputs"] else ())
res = [None, None]


def step(k, wait):
    res[k % 2] = ensemble.solve_ensemble(*args, tau=tau, phi=ens["phi_eval"], outputs=outputs, chunk=chunk, out=res[k % 2],
                                         wait=wait, **kw)


for k in range(3):
    step(k, True)
torch.cuda.synchronize()
for mode in ("one step, waited", "three steps back to back"):
    api._profile = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 1 if mode.startswith("one") else 3
    e0.record()
    t0 = time.perf_counter()
    for k in range(n):
        step(k, n == 1)
    host_ms = (time.perf_counter() - t0) * 1e3
    for r in res:
        r.wait()
    e1.record()
    torch.cuda.synchronize()
    marks, api._profile = api._profile, None
    inside = {}
    gaps = 0.0
    for (l0, ev0), (l1, ev1) in zip(marks[:-1], marks[1:]):
        dt = ev0.elapsed_time(ev1)
        if l1 == "begin":
            gaps += dt
        else:
            inside[l1] = inside.get(l1, 0.0) + dt
    wall = e0.elapsed_time(e1)
    first = e0.elapsed_time(marks[0][1])
    last = marks[-1][1].elapsed_time(e1)
    print(f"[rank {rank}] {mode}: wall {wall / n:.1f} ms per step; inside launches {sum(inside.values()) / n:.1f}; gaps between launches "
          f"{gaps / n:.1f}; before the first launch {first:.1f}; after the last launch (download tail) {last:.1f}; "
          f"host enqueue {host_ms / n:.1f} ms per step")
    if rank == 0:
        print("   ", {k: round(v / n, 1) for k, v in inside.items()})
