"""Experiment: where does the end-to-end arm of bench.py lose time against the device-resident arm?
Runs the three-stage pipeline (H2D one chunk ahead / compute / D2H behind) with each copy stage switched on and off."""
import sys, time, warnings
sys.path.insert(0, '.')
import numpy as np, torch
import pythonic_disort_b200 as pd
from pythonic_disort_b200 import synthetic
warnings.simplefilter('ignore')
NCH, CH = 4, 16384
B = NCH * CH
ens = synthetic.make('sw', B)
dev = torch.device('cuda', 0)
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory() if isinstance(x, np.ndarray) else x
args = [pin(a) for a in ens['args']]
kw = {k: ([pin(m) for m in v] if k == 'BDRF_Fourier_modes' else pin(v)) for k, v in ens['kwargs'].items()}
te = pin(ens['tau_eval']); phi = torch.as_tensor(ens['phi_eval'], device=dev)
sl = lambda x, lo, hi: x[lo:hi] if isinstance(x, torch.Tensor) and x.ndim >= 1 and x.shape[0] == B else x
h2d_s, d2h_s = torch.cuda.Stream(), torch.cuda.Stream()
host_out = {}
dev_in = {}

def run(do_h2d, do_d2h):
    cur = torch.cuda.current_stream()
    staged, done = {}, {}
    def stage(i):
        lo, hi = i * CH, (i + 1) * CH
        if not do_h2d and i in dev_in:
            staged[i] = dev_in[i] + (None,)
            return
        with torch.cuda.stream(h2d_s):
            f = lambda x: x.to(dev, non_blocking=True) if isinstance(x, torch.Tensor) else x
            ca = [f(sl(a, lo, hi)) for a in args]
            ck = {k: ([f(sl(m, lo, hi)) for m in v] if k == 'BDRF_Fourier_modes' else f(sl(v, lo, hi))) for k, v in kw.items()}
            t = f(te[lo:hi])
            ev = torch.cuda.Event(); ev.record(h2d_s)
        dev_in[i] = (ca, ck, t)
        staged[i] = (ca, ck, t, ev)
    h2d_s.wait_stream(cur)
    stage(0)
    for i in range(NCH):
        if i + 1 < NCH: stage(i + 1)
        ca, ck, t, ev = staged.pop(i)
        if ev is not None: cur.wait_event(ev)
        out = pd.pydisort(*ca, **ck)
        res = {'Fp': out[1](t)}
        res['Fm'], res['Fd'] = out[2](t)
        res['u'] = out[4](t, phi)
        if do_d2h:
            evc = torch.cuda.Event(); evc.record(cur)
            if i - 2 in done: done.pop(i - 2).synchronize()
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(evc)
                for name, x in res.items():
                    key = (i % 2, name)
                    if key not in host_out: host_out[key] = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
                    x.record_stream(d2h_s)
                    host_out[key].copy_(x, non_blocking=True)
                evd = torch.cuda.Event(); evd.record(d2h_s)
            done[i] = evd
        del out, res
    cur.wait_stream(d2h_s)

for cfg in [(True, True), (True, True), (False, False), (True, False), (False, True), (True, True)]:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    run(*cfg)
    torch.cuda.synchronize(); print('h2d=%s d2h=%s: %.1f ms' % (cfg + ((time.perf_counter() - t0) * 1e3,)))
