import sys, cProfile, pstats, warnings
sys.path.insert(0, '.')
import numpy as np, torch
import pythonic_disort_b200 as pd
from pythonic_disort_b200 import synthetic
warnings.simplefilter('ignore')
ens = synthetic.make('sw', 16384)
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory() if isinstance(x, np.ndarray) else x
args = [pin(a) for a in ens['args']]
kw = {k: ([pin(m) for m in v] if k == 'BDRF_Fourier_modes' else pin(v)) for k, v in ens['kwargs'].items()}
te = pin(ens['tau_eval']); phi = ens['phi_eval']
def step():
    out = pd.pydisort(*args, **kw)
    Fp = out[1](te); Fm = out[2](te); u = out[4](te, phi)
    return Fp, Fm, u
step(); step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print('e2e step s', time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
