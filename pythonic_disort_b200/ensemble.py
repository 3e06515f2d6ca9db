"""End-to-end solve of a column ensemble that lives in HOST memory: one call, host buffers in, host buffers out.

The reference's seam is one call per column (src/PythonicDISORT/pydisort.py:13-29) returning output functions; a
user with a million columns on the host calls it in a loop and evaluates the functions on the level grid.  Here the
same request is one call, ``solve_ensemble(<pydisort arguments with a leading column axis>, tau=..., ...)``, which
streams the ensemble through the GPU in chunks as a three-stage pipeline:

    copy stream A:   host inputs of chunk i+1  ->  device            (while chunk i computes)
    compute stream:  pydisort() + output functions of chunk i         (same kernels, same host wrapper)
    copy stream B:   results of chunk i-1      ->  pinned host output (while chunk i computes)

No host synchronisation happens inside the pipeline: structural decisions (is there a beam, delta-M, a thermal
source, ...) are taken once from the host arrays, and the value checks / status words the kernels produce are read
back once after the last chunk (a single 4-byte read behind a 400 MB result copy would otherwise stall the host,
and with it the next chunk's launches, for the length of that copy -- which is what kept the round-1 pipeline
9 % below the device-resident rate).  Inputs that are pinned torch tensors are copied asynchronously as they are;
NumPy / pageable inputs go through the driver's staging copy (correct, but not overlapped).  Results land in pinned
host tensors (``out=`` lets a caller recycle them) and are returned as NumPy views.
"""
import contextlib
import warnings

import numpy as np
import torch

from . import api
from .api import POSITIONAL, _F64

_streams = {}


def _copy_streams(dev):
    key = (dev.type, dev.index)
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return _streams[key]


class _NoStream:
    """Stand-in for streams / events when the backend has none (the CPU test-suite drives the same chunking and
    assembly logic through the host build of the kernels; the product always runs on CUDA)."""

    def wait_stream(self, *a): pass
    def wait_event(self, *a): pass
    def record(self, *a): pass
    def synchronize(self): pass


def pinned(x):
    """A pinned (page-locked) host copy of ``x`` as a torch tensor: the input form ``solve_ensemble`` can copy
    asynchronously.  Arrays that are already pinned tensors are returned as they are."""
    if isinstance(x, torch.Tensor):
        return x if (x.is_cuda or x.is_pinned()) else x.pin_memory()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).pin_memory()


def default_chunk(B, L, NQuad, NFourier, pipeline=True):
    """Columns per pydisort() call: solved state of a chunk (K, G, Bv, C) below ~24 GB, a multiple of 1024 (one that
    divides B if there is one) and, for the host pipeline (``pipeline=True``), at least six chunks per ensemble so that
    the first upload and the last download are small next to the rest."""
    N = NQuad // 2
    per_col = NFourier * L * (2 * N * N + 5 * N) * 8
    limit = max(1024, min(int(24e9 / per_col), 1 << 17) // 1024 * 1024)
    if pipeline and B >= 6 * 1024:
        limit = min(limit, max(1024, (B // 6) // 1024 * 1024))
    if B <= limit:
        return max(1, B)
    for c in range(limit, max(1023, limit // 2), -1024):
        if B % c == 0:
            return c
    return limit


class EnsembleResult(dict):
    """Dict of output name -> NumPy array (views of pinned host tensors).  With ``wait=False`` the arrays are being
    filled by the copy stream until :meth:`wait` returns (it also raises what the deferred input checks found)."""

    def __init__(self):
        super().__init__()
        self._done = None
        self._pending = []
        self._flags = None   # host copy of the deferred checks' values (pinned; filled by the download stream)
        self.tensors = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.chunks = 0

    def wait(self):
        if self._done is not None:
            self._done.synchronize()
            self._done = None
            pending, self._pending = self._pending, []
            flags, self._flags = self._flags, None
            if pending:
                api._raise_for_pending(pending, None if flags is None else flags.tolist())
        return self


def _host_flag(x):
    return api._host_any_nonzero(x)


def solve_ensemble(tau_arr, omega_arr, NQuad, Leg_coeffs_all, mu0, I0, phi0, *, tau, phi=None, mu=None,
                   outputs=("flux_up", "flux_down"), chunk=None, out=None, wait=True, **kwargs):
    """Solve a batch of columns held on the host and evaluate the requested outputs on ``tau`` (``[B, ntau]`` or
    ``[ntau]``), streaming chunks of columns through the GPU.

    Positional arguments and ``kwargs`` are those of :func:`pythonic_disort_b200.pydisort` in batch mode
    (``tau_arr`` is ``[B, NLayers]``).  ``outputs`` picks from ``"flux_up"``, ``"flux_down"`` (-> entries
    ``flux_down_diffuse`` and ``flux_down_direct``), ``"u0"`` and ``"u"`` (needs ``phi``); with ``mu`` given, ``u``
    and ``u0`` are interpolated to those polar-angle cosines on the device (``subroutines.interpolate``).
    Returns an :class:`EnsembleResult`: ``res["flux_up"]`` is ``[B, ntau]``, ``res["u"]`` is
    ``[B, NQuad or len(mu), ntau, nphi]``, ...  ``out`` may be a previous result whose pinned buffers are reused.
    ``wait=False`` returns as soon as everything is enqueued (call ``.wait()`` before reading), which lets
    consecutive ensembles overlap."""
    lib, dev = api._backend()
    cuda = dev.type == "cuda"
    args = (tau_arr, omega_arr, NQuad, Leg_coeffs_all, mu0, I0, phi0)
    if not (api._is_array(tau_arr) and tau_arr.ndim == 2):
        raise ValueError("solve_ensemble needs a batch of columns: `tau_arr` must be [B, NLayers].")
    B, L = int(tau_arr.shape[0]), int(tau_arr.shape[1])
    NQuad = int(NQuad)
    only_flux = bool(kwargs.get("only_flux", False))
    NF = 1 if only_flux else int(kwargs.get("NFourier") or NQuad)
    want = set(outputs)
    unknown = want - {"flux_up", "flux_down", "u0", "u"}
    if unknown:
        raise ValueError(f"unknown outputs {sorted(unknown)}")
    if "u" in want and only_flux:
        raise ValueError("`u` was requested together with only_flux=True")
    if "u" in want and phi is None:
        raise ValueError("`u` needs the azimuthal angles `phi`")
    if chunk is None:
        chunk = default_chunk(B, L, NQuad, NF)
    chunk = max(1, min(int(chunk), B))

    # structural decisions, once, from the host arrays (pydisort.py:199-217, :316, :375)
    hints = dict(f_nonzero=_host_flag(kwargs.get("f_arr", 0)), s_nonzero=_host_flag(kwargs.get("s_poly_coeffs", ())),
                 beam=_host_flag(I0), b_pos_zero=not _host_flag(kwargs.get("b_pos", 0)),
                 b_neg_zero=not _host_flag(kwargs.get("b_neg", 0)))
    tau_q = tau if isinstance(tau, torch.Tensor) else np.asarray(tau, dtype=np.float64)
    tau_batched = tau_q.ndim == 2
    if tau_batched and tau_q.shape[0] != B:
        raise ValueError("`tau` must be [ntau] or [B, ntau].")
    ntau = int(tau_q.shape[-1]) if tau_q.ndim else 1
    phi_d = None
    if phi is not None:
        phi_d = torch.as_tensor(np.atleast_1d(np.asarray(phi, dtype=np.float64)), device=dev)
    nphi = 0 if phi_d is None else int(phi_d.shape[0])
    nstream = NQuad if mu is None else len(np.atleast_1d(mu))
    shapes = {}
    if "flux_up" in want:
        shapes["flux_up"] = (B, ntau)
    if "flux_down" in want:
        shapes["flux_down_diffuse"] = shapes["flux_down_direct"] = (B, ntau)
    if "u0" in want:
        shapes["u0"] = (B, nstream, ntau)
    if "u" in want:
        shapes["u"] = (B, nstream, ntau, nphi)

    res = EnsembleResult()
    old = out.tensors if isinstance(out, EnsembleResult) else {}
    for name, shp in shapes.items():
        t = old.get(name)
        if t is None or tuple(t.shape) != shp or (cuda and not t.is_pinned()):
            t = torch.empty(shp, dtype=_F64, pin_memory=cuda)
        res.tensors[name] = t
        res[name] = t.numpy()

    # inputs shared by all columns go to the device once (no small copies inside the pipeline); BDRF callables are
    # tabulated once on the host (_solve_for_coeffs.py:121-134) instead of once per chunk
    N = NQuad // 2

    def shared_to_device(name, x):
        if isinstance(x, api.DeviceInput):
            return x if x.batched(B) else x.with_parts(tuple(shared_to_device("", p) for p in x.parts()))
        if api._is_array(x) and not api.carries_batch_axis(name, x, B, N, NF) and not (isinstance(x, torch.Tensor) and x.is_cuda):
            return torch.as_tensor(np.asarray(x, dtype=np.float64) if isinstance(x, np.ndarray) else x, dtype=_F64).to(dev)
        return x

    args = tuple(shared_to_device(n, a) for n, a in zip(POSITIONAL, args))
    modes = list(kwargs.get("BDRF_Fourier_modes", []))[:NF]
    if modes:
        mu_nodes = api.Gauss_Legendre_quad(N)[0]
        mu0_host = np.atleast_1d(mu0.cpu().numpy() if isinstance(mu0, torch.Tensor) else np.asarray(mu0, dtype=np.float64))
        tabs = []
        for fm in modes:
            if callable(fm) and not isinstance(fm, api.TabulatedBDRF):
                q = np.asarray(fm(mu_nodes, mu_nodes), dtype=float)
                q0 = None
                if hints["beam"]:
                    pts = mu0_host[:1] if np.all(mu0_host == mu0_host[0]) else mu0_host
                    q0 = np.asarray(fm(mu_nodes, pts), dtype=float)
                fm = api.TabulatedBDRF(q, q0)
            if isinstance(fm, api.TabulatedBDRF) and not (isinstance(fm.q, torch.Tensor) and fm.q.is_cuda):
                fm = api.TabulatedBDRF(torch.as_tensor(fm.q, dtype=_F64).to(dev),
                                       None if fm.q0 is None else torch.as_tensor(fm.q0, dtype=_F64).to(dev))
            tabs.append(fm)
        kwargs = dict(kwargs, BDRF_Fourier_modes=tabs)
    kwargs = {k: (v if k == "BDRF_Fourier_modes" else shared_to_device(k, v)) for k, v in kwargs.items()}

    if cuda:
        cur = torch.cuda.current_stream(dev)
        h2d_stream, d2h_stream = _copy_streams(dev)
        on = torch.cuda.stream
        new_event = torch.cuda.Event
    else:
        cur = h2d_stream = d2h_stream = _NoStream()
        on = lambda s_: contextlib.nullcontext()  # noqa: E731
        new_event = _NoStream
    # inputs a caller prepared ON THE DEVICE (on its current stream) must be complete before the copy stream slices them;
    # host inputs need no such ordering, and without it the first upload of this call overlaps whatever the compute
    # stream is still working on (the tail of the previous ensemble, when calls are issued back to back with wait=False)
    def _on_device(x):
        if isinstance(x, api.DeviceInput):
            return x.on_device()
        if isinstance(x, (list, tuple)):
            return any(_on_device(y) for y in x)
        if isinstance(x, api.TabulatedBDRF):
            return False  # tables were placed on the device above, on the current stream, and are only read by kernels on it
        return isinstance(x, torch.Tensor) and x.is_cuda and x.ndim > 0 and x.shape[0] == B

    if any(_on_device(x) for x in list(args) + list(kwargs.values()) + [tau_q]):
        h2d_stream.wait_stream(cur)
    moved = [0]

    def up(x):
        if isinstance(x, api.DeviceInput):  # only the description crosses the link; pydisort() expands it on the device
            return x.with_parts(tuple(up(p) for p in x.parts()))
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if isinstance(x, torch.Tensor) and not x.is_cuda:
            moved[0] += x.numel() * x.element_size()
            if not cuda:
                return x
            d = x.to(dev, non_blocking=True)
            d.record_stream(cur)
            return d
        return x

    def stage(lo, hi):
        ca, ck = api.slice_columns(args, kwargs, lo, hi, B)
        with on(h2d_stream):
            ca = tuple(up(a) for a in ca)
            ck = {k: ([up(m) if api._is_array(m) else m for m in v] if k == "BDRF_Fourier_modes" else up(v))
                  for k, v in ck.items()}
            tq = up(tau_q[lo:hi] if tau_batched else tau_q)
            ev = new_event()
            ev.record(h2d_stream)
        return ca, ck, tq, ev

    starts = list(range(0, B, chunk))
    nxt = stage(starts[0], min(B, starts[0] + chunk))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # per-chunk warnings would repeat; the deferred check re-issues them once
        for i, lo in enumerate(starts):
            hi = min(B, lo + chunk)
            ca, ck, tq, ev = nxt
            if i + 1 < len(starts):
                nxt = stage(starts[i + 1], min(B, starts[i + 1] + chunk))
            cur.wait_event(ev)
            sol = api.pydisort(*ca, **ck, _defer=res._pending, _hints=hints)
            got = {}
            if "flux_up" in want:
                got["flux_up"] = sol[1](tq)
            if "flux_down" in want:
                got["flux_down_diffuse"], got["flux_down_direct"] = sol[2](tq)
            if "u0" in want:
                got["u0"] = sol[3](tq) if mu is None else sol[3].at_mu(mu, tq)
            if "u" in want:
                got["u"] = sol[4](tq, phi_d) if mu is None else sol[4].at_mu(mu, tq, phi_d)
            done = new_event()
            done.record(cur)
            with on(d2h_stream):
                d2h_stream.wait_event(done)
                for name, t in got.items():
                    t = torch.as_tensor(t).reshape((hi - lo,) + tuple(res.tensors[name].shape[1:]))
                    if cuda:
                        t.record_stream(d2h_stream)
                    res.tensors[name][lo:hi].copy_(t, non_blocking=True)
                    res.d2h_bytes += t.numel() * 8
            del sol, got
    res.h2d_bytes = moved[0]
    res.chunks = len(starts)
    if res._pending:
        # the values of the deferred checks travel with the results: reduced on the compute stream behind the last chunk,
        # downloaded by the copy stream, read by wait() from host memory
        flags = api._pending_flags(res._pending)
        done = new_event()
        done.record(cur)
        res._flags = torch.empty(flags.shape, dtype=flags.dtype, pin_memory=cuda)
        with on(d2h_stream):
            d2h_stream.wait_event(done)
            if cuda:
                flags.record_stream(d2h_stream)
            res._flags.copy_(flags, non_blocking=True)
    res._done = new_event()
    res._done.record(d2h_stream)
    return res.wait() if wait else res
