"""Experiment: do two host threads on two CUDA streams overlap the copies of one chunk with the kernels of another?"""
import sys, time, warnings, threading, concurrent.futures
sys.path.insert(0, '.')
import numpy as np, torch
import pythonic_disort_b200 as pd
from pythonic_disort_b200 import synthetic
warnings.simplefilter('ignore')
NCH, CH = 4, 16384
ens = synthetic.make('sw', NCH * CH)
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory() if isinstance(x, np.ndarray) else x
args = [pin(a) for a in ens['args']]
kw = {k: ([pin(m) for m in v] if k == 'BDRF_Fourier_modes' else pin(v)) for k, v in ens['kwargs'].items()}
te = pin(ens['tau_eval']); phi = ens['phi_eval']
B = NCH * CH
sl = lambda x, lo, hi: x[lo:hi] if isinstance(x, torch.Tensor) and x.ndim >= 1 and x.shape[0] == B else x
t00 = [0.0]
def chunk(i, log):
    lo, hi = i * CH, (i + 1) * CH
    t0 = time.perf_counter()
    out = pd.pydisort(*[sl(a, lo, hi) for a in args], **{k: ([sl(m, lo, hi) for m in v] if k == 'BDRF_Fourier_modes' else sl(v, lo, hi)) for k, v in kw.items()})
    t1 = time.perf_counter()
    Fp = out[1](te[lo:hi]); Fm = out[2](te[lo:hi])
    t2 = time.perf_counter()
    u = out[4](te[lo:hi], phi)
    t3 = time.perf_counter()
    log.append((i, round((t0 - t00[0]) * 1e3, 1), round((t1 - t00[0]) * 1e3, 1), round((t2 - t00[0]) * 1e3, 1), round((t3 - t00[0]) * 1e3, 1)))
for i in range(NCH): chunk(i, [])
torch.cuda.synchronize()
log = []; t00[0] = time.perf_counter()
for i in range(NCH): chunk(i, log)
torch.cuda.synchronize(); print('sequential ms', round((time.perf_counter() - t00[0]) * 1e3, 1)); print(log)
streams = [torch.cuda.Stream() for _ in range(2)]
pool = concurrent.futures.ThreadPoolExecutor(2)
def on_stream(i, log):
    with torch.cuda.stream(streams[i % 2]):
        chunk(i, log)
        streams[i % 2].synchronize()
for rep in range(2):
    log = []; torch.cuda.synchronize(); t00[0] = time.perf_counter()
    futs = [pool.submit(on_stream, i, log) for i in range(NCH)]
    [f.result() for f in futs]
    torch.cuda.synchronize(); print('two threads ms', round((time.perf_counter() - t00[0]) * 1e3, 1)); print(sorted(log))
