"""Host side of the drop-in boundary: ``pydisort`` with the reference's call
signature (src/PythonicDISORT/pydisort.py:13-29) extended by a leading
batch-of-columns dimension, returning the same output functions
(src/PythonicDISORT/_assemble_intensity_and_fluxes.py:170-613).

Everything numerical happens in libpydisort_b200.so (CUDA, sm_100a) through
ctypes; PyTorch only owns device memory and the stream.  There is no CPU or
PyTorch implementation to fall back to.

Batch conventions (B = number of columns)
  * ``tau_arr`` 2-D ``[B, NLayers]`` switches batch mode on; with a 0-D/1-D
    ``tau_arr`` the call is the reference's single-column call and outputs are
    squeezed exactly like the reference's.
  * In batch mode every other array input may carry a leading ``B`` axis
    (``omega_arr [B, L]``, ``Leg_coeffs_all [B, L, NLeg_all]``, ``f_arr [B, L]``,
    ``s_poly_coeffs [B, L, Ns]``, ``mu0 / I0 / phi0 [B]``) or keep the reference's
    unbatched shape, in which case it is shared by all columns.
  * ``b_pos`` / ``b_neg`` per column: ``[B]`` or ``[B, 1]`` (isotropic), ``[B, N]``
    or ``[B, N, NFourier]``.
  * ``BDRF_Fourier_modes`` entries: a scalar, a callable ``f(mu, -mu_p)``, a
    ``subroutines.TabulatedBDRF`` or (batch mode) an array ``[B]`` of per-column
    Lambertian albedos.
  * Output functions take ``tau`` as a scalar / 1-D array shared by all columns
    or ``[B, ntau]``; results have shape ``[B, <reference shape>]``.
  * NumPy in -> NumPy out; if any input is a CUDA tensor, outputs are CUDA tensors.
"""
import ctypes
import warnings
from math import pi

import numpy as np
import torch

from . import _lib
from .subroutines import Gauss_Legendre_quad, TabulatedBDRF

_F64 = torch.float64
_const_cache = {}
_interp_cache = {}
_profile = None  # bench.py sets this to a list to collect (label, CUDA event) marks around each launch


def _mark(label, dev):
    if _profile is not None and dev.type == "cuda":
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(dev))
        _profile.append((label, ev))


def _backend():
    """The CUDA library and the current device; there is nothing else to run on."""
    if not torch.cuda.is_available():
        raise RuntimeError("pythonic_disort_b200 needs a CUDA device (B200, sm_100a); it has no CPU fallback")
    return _lib.cuda_lib(), torch.device("cuda", torch.cuda.current_device())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if dev.type == "cuda" else None


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def norm_assoc_legendre_table(NF, NLeg, x):
    """P~_l^m(x_i) = sqrt((l-m)!/(l+m)!) P_l^m(x_i) as ``[NF, NLeg, len(x)]``
    (zero for l < m); same recurrence as csrc/pd_prologue.cuh."""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    out = np.zeros((NF, NLeg, len(x)))
    s = np.sqrt(np.maximum(0.0, 1.0 - x * x))
    pmm = np.ones_like(x)
    for m in range(NF):
        if m > 0:
            pmm = pmm * (-np.sqrt((2.0 * m - 1.0) / (2.0 * m)) * s)
        if m >= NLeg:
            break
        out[m, m] = pmm
        pm1, p = np.zeros_like(x), pmm
        for l in range(m, NLeg - 1):
            nxt = ((2.0 * l + 1.0) * x * p - np.sqrt(float((l + m) * (l - m))) * pm1) / np.sqrt(
                float((l + 1 - m) * (l + 1 + m)))
            out[m, l + 1] = nxt
            pm1, p = p, nxt
    return out


def _constants(NQuad, NLeg, NF, dev):
    key = (NQuad, NLeg, NF, str(dev))
    if key not in _const_cache:
        N = NQuad // 2
        mu, W = Gauss_Legendre_quad(N)
        ptab = norm_assoc_legendre_table(NF, NLeg, mu)
        _const_cache[key] = (mu, W, torch.as_tensor(mu, dtype=_F64, device=dev),
                             torch.as_tensor(W, dtype=_F64, device=dev),
                             torch.as_tensor(ptab, dtype=_F64, device=dev).contiguous())
    return _const_cache[key]


def _content_tag(x):
    """Cheap 'has this array been modified in place' tag for the flux memo."""
    if isinstance(x, torch.Tensor):
        return x._version
    if isinstance(x, np.ndarray):
        return hash(x.tobytes()) if x.size <= 4096 else (x.ctypes.data, x.shape, float(x.flat[0]), float(x.flat[-1]))
    return None


def _is_dev_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


class _Solution:
    """Solved state of a batch of columns (device tensors) + evaluation launches."""

    def __init__(self):
        pass

    # -- helpers -----------------------------------------------------------------
    def _tau_points(self, tau):
        t = torch.as_tensor(tau, dtype=_F64, device=self.dev) if not isinstance(tau, torch.Tensor) \
            else tau.to(device=self.dev, dtype=_F64)
        scalar = t.ndim == 0
        if t.ndim == 0:
            t = t[None]
        if t.ndim == 1:
            tq = t[None, :].expand(self.B, t.shape[0])
        elif t.ndim == 2 and self.batched and t.shape[0] == self.B:
            tq = t
        else:
            raise ValueError("tau must be a scalar, a 1-D array, or (batch mode) a [B, ntau] array.")
        tq = tq.contiguous()
        if isinstance(tau, torch.Tensor) and tau.is_cuda or self.defer is not None:
            # device-resident query points: the range test runs on the device and is read back with the results
            # (or, in a pipelined ensemble solve, once after the last chunk) -- no stream drain here
            self._note("tau_range", ((tq < 0) | (tq > self.tau[:, -1:])).any())
        else:
            th = tau.detach().cpu().numpy() if isinstance(tau, torch.Tensor) else np.asarray(tau, dtype=np.float64)
            if self._tau_max is None:
                self._tau_max = self.tau[:, -1].cpu().numpy()
            over = np.any(th > self._tau_max[:, None]) if th.ndim == 2 else np.max(th, initial=0.0) > self._tau_max.min()
            if np.any(th < 0) or over:
                raise ValueError("tau input outside the tau range given for the atmosphere (check `tau_arr`).")
        return tq, scalar

    def _note(self, kind, flag):
        """Queue a device-side check; `raise_pending` turns it into the reference's exception."""
        (self.defer if self.defer is not None else self.pending).append((kind, flag))

    def raise_pending(self):
        if self.pending:
            pending, self.pending[:] = list(self.pending), []
            _raise_for_pending(pending)

    def _finish(self, out, ref_dims):
        """Squeeze like the reference (``np.squeeze(x)[()]``) but never the batch axis."""
        keep = [d for d in ref_dims if d != 1]
        out = out.reshape([self.B] + keep)
        if not self.batched:
            out = out[0]
        if self.want_torch:
            if self.pending:
                self.raise_pending()
            return out
        if out.is_cuda and out.numel() >= 1 << 16:
            # large result: device -> pinned host buffer (torch's caching host allocator recycles it once the
            # caller drops the array), several times faster than a pageable copy
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream(out.device).synchronize()
            return host.numpy()
        res = out.cpu().numpy()
        return res[()] if res.ndim == 0 else res

    def _state(self):
        st = _lib.pd_state()
        for name, t in (("tau", self.tau), ("taus", self.taus), ("scale_tau", self.scale_tau), ("colp", self.colp),
                        ("K", self.K), ("G", self.G), ("Bv", self.Bv), ("dth", self.dth), ("C", self.C),
                        ("mu_nodes", self.mu_d), ("w_nodes", self.w_d), ("Uif", getattr(self, "Uif", None))):
            setattr(st, name, t.data_ptr() if t is not None else None)
        return st

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"libpydisort_b200: {what} failed with code {rc}")

    # -- launches ----------------------------------------------------------------
    def interp_mu(self, vals, mu):
        """``vals`` [B, 2N, ...] on the device -> [B, nmu, ...] at the polar-angle cosines ``mu`` (row f1)."""
        key = (tuple(np.atleast_1d(np.asarray(mu, dtype=np.float64)).tolist()), self.NQuad, str(self.dev))
        if key not in _interp_cache:
            _interp_cache[key] = torch.as_tensor(barycentric_weight_matrix(mu, self.NQuad // 2), dtype=_F64,
                                                 device=self.dev).contiguous()
        W = _interp_cache[key]
        nmu = W.shape[0]
        vals = vals.contiguous()
        M = int(np.prod(vals.shape[2:])) if vals.ndim > 2 else 1
        out = torch.empty((self.B, nmu) + tuple(vals.shape[2:]), dtype=_F64, device=self.dev)
        _mark("begin", self.dev)
        self._check(self.lib.pd_interp_mu(self.B, self.NQuad, M, nmu, _ptr(W), _ptr(vals), _ptr(out), _stream(self.dev)),
                    "pd_interp_mu")
        _mark("interp_mu", self.dev)
        return out, nmu

    def eval_flux(self, tau, anti):
        # flux_up(tau) then flux_down(tau) on the same array object: one launch serves the pair.  The memo is consumed
        # by its first hit, so it never outlives that pair (an array edited in place between later calls, or a caller
        # writing into a returned tensor, cannot meet a stale entry); up / diffuse / direct are disjoint rows of `out`.
        memo, self._flux_memo = getattr(self, "_flux_memo", None), None
        if memo is not None and memo[0] is tau and memo[1] == bool(anti) and memo[4] == _content_tag(tau):
            return memo[2], memo[3]
        tq, _ = self._tau_points(tau)
        ntau = tq.shape[1]
        out = torch.empty((3, self.B, ntau), dtype=_F64, device=self.dev)
        st = self._state()
        _mark("begin", self.dev)
        self._check(self.lib.pd_eval_flux(ctypes.byref(self.cfg), ctypes.byref(st), _ptr(tq), ntau, int(bool(anti)),
                                          _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream(self.dev)), "pd_eval_flux")
        _mark("eval_flux", self.dev)
        self._flux_memo = (tau, bool(anti), out, ntau, _content_tag(tau)) \
            if isinstance(tau, (np.ndarray, torch.Tensor)) else None
        return out, ntau

    def eval_u0(self, tau, anti, want_recl):
        tq, _ = self._tau_points(tau)
        ntau = tq.shape[1]
        u0 = torch.empty((self.B, self.NQuad, ntau), dtype=_F64, device=self.dev)
        recl = torch.empty((self.B, ntau), dtype=_F64, device=self.dev) if want_recl else None
        st = self._state()
        _mark("begin", self.dev)
        self._check(self.lib.pd_eval_u0(ctypes.byref(self.cfg), ctypes.byref(st), _ptr(tq), ntau, int(bool(anti)),
                                        _ptr(u0), _ptr(recl), _stream(self.dev)), "pd_eval_u0")
        _mark("eval_u0", self.dev)
        return u0, recl, ntau

    def eval_u(self, tau, phi, anti, nt, want_last):
        tq, _ = self._tau_points(tau)
        ntau = tq.shape[1]
        ph = torch.as_tensor(phi, dtype=_F64, device=self.dev) if not isinstance(phi, torch.Tensor) \
            else phi.to(device=self.dev, dtype=_F64)
        if ph.ndim == 0:
            ph = ph[None]
        if ph.ndim != 1:
            raise ValueError("phi must be a scalar or a 1-D array.")
        ph = ph.contiguous()
        nphi = ph.shape[0]
        u = torch.empty((self.B, self.NQuad, ntau, nphi), dtype=_F64, device=self.dev)
        ulast = torch.empty((self.B, self.NQuad, ntau), dtype=_F64, device=self.dev) if want_last else None
        st = self._state()
        _mark("begin", self.dev)
        self._check(self.lib.pd_eval_u(ctypes.byref(self.cfg), ctypes.byref(st), _ptr(tq), ntau, _ptr(ph), nphi,
                                       int(bool(anti)), int(bool(nt)), _ptr(self.omega), _ptr(self.f),
                                       _ptr(self.leg_all), _ptr(self.omega_s), _ptr(self.wleg), _ptr(u), _ptr(ulast),
                                       _stream(self.dev)), "pd_eval_u")
        _mark("eval_u", self.dev)
        return u, ulast, ph, ntau, nphi


_gl16_cache = {}


def _gl16(dev):
    """16-point Gauss-Legendre rule (nodes then weights) on the device, for the Planck band integrals."""
    if dev not in _gl16_cache:
        x, w = np.polynomial.legendre.leggauss(16)
        _gl16_cache[dev] = torch.as_tensor(np.concatenate([x, w]), dtype=_F64, device=dev)
    return _gl16_cache[dev]


def planck_band(T, WVNMLO, WVNMHI):
    """Band-integrated blackbody emission of every temperature in the tensor ``T`` (any shape), on the device
    (row f2; ``subroutines.blackbody_contrib_to_BCs``, subroutines.py:354-377)."""
    lib, dev = _backend()
    Tt = T.to(device=dev, dtype=_F64).contiguous()
    out = torch.empty_like(Tt)
    if Tt.numel():
        rc = lib.pd_planck_band(Tt.numel(), _ptr(Tt), float(WVNMLO), float(WVNMHI), _ptr(_gl16(dev)), _ptr(out),
                                _stream(dev))
        if rc != 0:
            raise RuntimeError(f"libpydisort_b200: pd_planck_band failed with code {rc}")
    return out


def s_poly_coeffs(tau_arr, TEMPER, WVNMLO, WVNMHI):
    """``generate_s_poly_coeffs`` (subroutines.py:413-454) for a batch: ``tau_arr`` [B, L] (or [L]), ``TEMPER``
    [B, L+1] (or [L+1]) tensors -> ``s_poly_coeffs`` [B, L, 2] (or [L, 2]) on the device, ready for ``pydisort``."""
    lib, dev = _backend()
    tau = tau_arr.to(device=dev, dtype=_F64)
    tem = TEMPER.to(device=dev, dtype=_F64)
    single = tau.ndim == 1
    if single:
        tau, tem = tau[None], tem[None]
    if tau.ndim != 2 or tem.ndim != 2 or tem.shape[0] != tau.shape[0] or tem.shape[1] != tau.shape[1] + 1:
        raise ValueError("Missing temperature specification at some boundaries / interfaces.")
    tau, tem = tau.contiguous(), tem.contiguous()
    B, L = tau.shape
    out = torch.empty((B, L, 2), dtype=_F64, device=dev)
    rc = lib.pd_s_poly_coeffs(B, L, _ptr(tau), _ptr(tem), float(WVNMLO), float(WVNMHI), _ptr(_gl16(dev)), _ptr(out),
                              _stream(dev))
    if rc != 0:
        raise RuntimeError(f"libpydisort_b200: pd_s_poly_coeffs failed with code {rc}")
    return out[0] if single else out


def hapke_modes(N, NF, mup, B0, HH, W, n_panels):
    """Fourier modes of the Hapke BDRF on the device (row f4): ``mup`` is a 1-D tensor of incidence cosines (the
    quadrature nodes, or one beam cosine per column); returns a tensor ``[NF, N, len(mup)]``."""
    from .subroutines import Gauss_Legendre_quad
    lib, dev = _backend()
    mu = torch.as_tensor(Gauss_Legendre_quad(N)[0], dtype=_F64, device=dev)
    mp = mup.to(device=dev, dtype=_F64).contiguous().reshape(-1)
    out = torch.empty((NF, N, mp.numel()), dtype=_F64, device=dev)
    rc = lib.pd_hapke_modes(N, mp.numel(), NF, int(n_panels), _ptr(_gl16(dev)), _ptr(mu), _ptr(mp), float(B0), float(HH),
                            float(W), _ptr(out), _stream(dev))
    if rc != 0:
        raise RuntimeError(f"libpydisort_b200: pd_hapke_modes failed with code {rc}")
    return out


def barycentric_weight_matrix(mu, N):
    """Weight matrix [nmu, 2N] of ``subroutines.interpolate`` (subroutines.py:614-705): row o holds the barycentric
    Lagrange weights of the N Gauss-Legendre streams of the hemisphere of ``mu[o]`` (``mu > 0``: the upward streams
    ``mu_i``, columns 0..N-1; otherwise the downward streams ``-mu_i``, columns N..2N-1) and zeros elsewhere."""
    from .subroutines import Gauss_Legendre_quad
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    nodes = Gauss_Legendre_quad(N)[0]
    diff = nodes[:, None] - nodes[None, :]
    np.fill_diagonal(diff, 1.0)
    bw = 1.0 / np.prod(diff, axis=1)  # w_j = 1 / prod_{k != j} (x_j - x_k); identical for the mirrored node set up to sign
    W = np.zeros((mu.shape[0], 2 * N))
    for o, x in enumerate(mu):
        up = x > 0
        xj = nodes if up else -nodes
        wj = bw if up else bw * (-1.0) ** (N - 1)
        d = x - xj
        hit = np.nonzero(d == 0.0)[0]
        if hit.size:
            row = np.zeros(N)
            row[hit[0]] = 1.0
        else:
            t = wj / d
            row = t / np.sum(t)
        W[o, (0 if up else N):(N if up else 2 * N)] = row
    return W


def _bc_tensor(b, name, B, N, NF, batched, T, is_zero=None):
    """Dirichlet boundary values as [B, NFb, N] (pydisort.py:199-202, :274-285).

    Reference shapes: scalar, [N] (mode 0) or [N, NFourier].  Batch mode adds the per-column forms [B, 1]
    (isotropic), [B, N] and [B, N, NFourier]; a 1-D [B] array is taken as per-column isotropic values.  A shape
    that could be read either way (B == N, or (B, N) == (N, NFourier)) is rejected instead of guessed."""
    t = T(b)
    if is_zero is None:
        is_zero = not _host_any_nonzero(b) if not (isinstance(b, torch.Tensor) and b.is_cuda) else t.numel() == 0
    if t.numel() == 0 or is_zero:
        return torch.zeros((B, 1, N), dtype=_F64, device=t.device)
    which = "bottom" if name == "b_pos" else "top"
    err = ValueError(f"The shape of the {which} boundary condition is incorrect.")
    if t.numel() == 1 and t.ndim <= 1:
        return t.reshape(1, 1, 1).expand(B, 1, N).contiguous()
    if batched:
        amb = ValueError(f"`{name}` of shape {tuple(t.shape)} is ambiguous for a batch of {B} columns with "
                         f"{N} streams per hemisphere: pass per-column values as [B, 1], [B, N] or [B, N, NFourier].")
        if t.ndim == 1 and t.shape[0] == B:
            if B == N:
                raise amb
            return t[:, None, None].expand(B, 1, N).contiguous()
        if t.ndim == 2 and t.shape == (B, 1):
            return t[:, :, None].expand(B, 1, N).contiguous()
        if t.ndim == 2 and t.shape == (B, N):
            if (B, N) == (N, NF):
                raise amb
            return t[:, None, :].contiguous()
        if t.ndim == 3 and t.shape == (B, N, NF):
            return t.transpose(1, 2).contiguous()
    if t.ndim == 1 and t.shape[0] == N:
        return t[None, None, :].expand(B, 1, N).contiguous()
    if t.ndim == 2 and t.shape == (N, NF):
        return t.t()[None].expand(B, NF, N).contiguous()
    raise err


# Which inputs of a batched call carry the leading column axis is decided from each argument's RANK (the unbatched
# rank is fixed by the reference's signature), never from "shape[0] happens to equal B".
from .inputs import DeviceInput, HenyeyGreenstein, LevelSource  # noqa: E402,F401

POSITIONAL = ("tau_arr", "omega_arr", "NQuad", "Leg_coeffs_all", "mu0", "I0", "phi0")
_UNBATCHED_NDIM = {"tau_arr": 1, "omega_arr": 1, "Leg_coeffs_all": 2, "f_arr": 1, "s_poly_coeffs": 2,
                   "mu0": 0, "I0": 0, "phi0": 0}


def _is_array(x):
    return isinstance(x, (np.ndarray, torch.Tensor))


def carries_batch_axis(name, x, B, N=None, NF=None):
    """True if input `name` of a batched call (B columns) has the leading column axis."""
    if name == "BDRF_mode":
        return _is_array(x) and x.ndim == 1 and x.shape[0] == B
    if isinstance(x, DeviceInput):
        return x.batched(B)
    if not _is_array(x):
        return False
    if name in _UNBATCHED_NDIM:
        return x.ndim == _UNBATCHED_NDIM[name] + 1 and x.shape[0] == B
    if name in ("b_pos", "b_neg"):  # same reading order as _bc_tensor (which rejects the ambiguous shapes)
        if x.numel() if isinstance(x, torch.Tensor) else x.size:
            if x.ndim == 3:
                return x.shape[0] == B
            if x.ndim == 2:
                return tuple(x.shape) in ((B, 1), (B, N))
            if x.ndim == 1:
                return x.shape[0] == B and B != 1
    return False


def slice_columns(args, kwargs, lo, hi, B=None):
    """Columns [lo, hi) of a batched pydisort() call: every input that carries the column axis is sliced, shared
    inputs are passed through.  Used by the ensemble pipeline and by parallel.shard_inputs."""
    tau = args[0]
    if not (_is_array(tau) and tau.ndim == 2):
        raise ValueError("slice_columns needs a batched call: `tau_arr` must be [B, NLayers].")
    if B is None:
        B = tau.shape[0]
    NQuad = int(args[2])
    N = NQuad // 2
    NF = 1 if kwargs.get("only_flux") else int(kwargs.get("NFourier") or NQuad)

    def cut(name, x):
        return x[lo:hi] if carries_batch_axis(name, x, B, N, NF) else x

    def cut_mode(m):
        if isinstance(m, TabulatedBDRF) and m.q0 is not None and _is_array(m.q0) and m.q0.ndim == 2 and \
                m.q0.shape[1] == B and B > 1:
            return TabulatedBDRF(m.q, m.q0[:, lo:hi])
        return cut("BDRF_mode", m)

    out_args = tuple(cut(n, a) for n, a in zip(POSITIONAL, args))
    out_kw = {}
    for k, v in kwargs.items():
        out_kw[k] = [cut_mode(m) for m in v] if k == "BDRF_Fourier_modes" else cut(k, v)
    return out_args, out_kw


def _bdrf_tables(modes, B, N, mu_pos, mu0_t, beam, batched, T):
    """Tabulate the BDRF Fourier modes at the quadrature nodes
    (_solve_for_coeffs.py:121-134) as device tensors q[(B), NBDRF, N, N] and
    q0[(B), NBDRF, N].  Scalars and per-column albedo arrays are expanded on the
    device; callables are evaluated once on the host."""
    qs, q0s, percol = [], [], False
    mu0_host = None
    for fm in modes:
        if isinstance(fm, TabulatedBDRF):
            q = T(fm.q)[None]
            q0 = T(fm.q0.T if fm.q0 is not None else np.zeros((1, N)))  # [k, N]
            if q0.shape[0] not in (1, B):
                raise ValueError("TabulatedBDRF.q0 must have 1 or B columns.")
        elif callable(fm):
            q = T(np.asarray(fm(mu_pos, mu_pos), dtype=float))[None]
            if beam:
                if mu0_host is None:
                    mu0_host = mu0_t.cpu().numpy()
                same = bool(np.all(mu0_host == mu0_host[0]))
                pts = mu0_host[:1] if same else mu0_host
                q0 = T(np.ascontiguousarray(np.asarray(fm(mu_pos, np.asarray(pts)), dtype=float).T))  # [k, N]
            else:
                q0 = torch.zeros((1, N), dtype=_F64, device=mu0_t.device)
        else:
            a = T(fm)
            if a.ndim == 0:
                q, q0 = a.reshape(1, 1, 1).expand(1, N, N), a.reshape(1, 1).expand(1, N)
            elif batched and a.shape == (B,):
                q, q0 = a[:, None, None].expand(B, N, N), a[:, None].expand(B, N)
            else:
                raise ValueError("BDRF Fourier modes must be scalars, callables, TabulatedBDRF or [B] arrays.")
        percol = percol or q.shape[0] > 1 or q0.shape[0] > 1
        qs.append(q)
        q0s.append(q0)
    Bq = B if percol else 1
    Q = torch.stack([q.expand(Bq, N, N) for q in qs], dim=1).contiguous()
    Q0 = torch.stack([q0.expand(Bq, N) for q0 in q0s], dim=1).contiguous()
    return Q, Q0, percol


def pydisort(
    tau_arr, omega_arr,
    NQuad,
    Leg_coeffs_all,
    mu0, I0, phi0,
    NLeg=None,
    NFourier=None,
    b_pos=0,
    b_neg=0,
    only_flux=False,
    f_arr=0,
    NT_cor=False,
    BDRF_Fourier_modes=[],
    s_poly_coeffs=np.array([[]]),
    use_banded_solver_NLayers=10,
    autograd_compatible=False,
    _kernel_flags=0,
    _defer=None,
    _hints=None,
):
    """Solve the 1-D RTE for a column or a batch of columns on the GPU.

    Arguments, defaults, input checks and returned functions follow the
    reference ``PythonicDISORT.pydisort`` (pydisort.py:13-128); see the module
    docstring for the batch extension.  Returns
    ``(mu_arr, flux_up, flux_down, u0)`` plus ``u`` unless ``only_flux``.
    ``use_banded_solver_NLayers`` is accepted and validated but unused (one
    block solver covers both of the reference's LAPACK paths).  ``_kernel_flags`` is a test hook: extra
    ``pd_config.flags`` bits (``_lib.PD_FLAG_GENERIC_KERNELS`` runs the size-generic kernels).

    Host synchronisation: inputs that live on the host are inspected on the host; device tensors are never read
    back to decide anything (a device ``f_arr`` / ``s_poly_coeffs`` / ``I0`` is taken as "may be non-zero", which
    the kernels resolve per column).  The value checks of pydisort.py:223-291 run in the prologue kernel and are
    read back once, after the solve kernels have been enqueued; ``ensemble.solve_ensemble`` defers even that to the
    end of its pipeline (``_defer`` / ``_hints`` are its private hooks)."""
    if autograd_compatible:
        raise NotImplementedError("autograd_compatible=True is not supported by the CUDA implementation.")
    lib, dev = _backend()
    ins = (tau_arr, omega_arr, Leg_coeffs_all, mu0, I0, phi0, b_pos, b_neg, f_arr, s_poly_coeffs)
    want_torch = any(x.on_device() if isinstance(x, DeviceInput) else _is_dev_tensor(x) for x in ins)

    def T(x):
        if isinstance(x, torch.Tensor):
            return x.to(device=dev, dtype=_F64, non_blocking=True)
        a = np.asarray(x, dtype=np.float64)
        if a.size == 1:  # a fill kernel, not a host->device copy (which would queue behind any large upload in flight)
            return torch.full(a.shape, float(a.reshape(-1)[0]), dtype=_F64, device=dev)
        return torch.as_tensor(a).to(dev, non_blocking=True)

    # ---- shapes (pydisort.py:184-217) ------------------------------------------
    tau = T(tau_arr)
    batched = tau.ndim == 2
    if tau.ndim == 0:
        tau = tau[None]
    if tau.ndim == 1:
        tau = tau[None, :]
    if tau.ndim != 2:
        raise ValueError("`tau_arr` must be a scalar, [NLayers] or [B, NLayers].")
    B, L = tau.shape
    NQuad = int(NQuad)
    if NLeg is None:
        NLeg = NQuad
    if only_flux:
        NFourier = 1
    elif NFourier is None:
        NFourier = NQuad
    NLeg, NFourier = int(NLeg), int(NFourier)
    N = NQuad // 2

    def per_layer(x, trailing, msg):
        """[..., L, *trailing] with an optional leading B axis -> [B, L, *trailing]."""
        t = T(x)
        nd = 1 + len(trailing)
        if t.ndim == nd - 1 and L == 1:      # a single layer given without its layer axis
            t = t[None]
        if t.ndim == nd:
            if t.shape[0] != L:
                raise ValueError(msg)
            return t[None].expand((B,) + tuple(t.shape))
        if t.ndim == nd + 1 and batched and t.shape[0] == B:
            if t.shape[1] != L:
                raise ValueError(msg)
            return t
        raise ValueError(msg)

    omega = per_layer(omega_arr, (), "The zeroth dimension of the shape of `omega_arr` does not match the number of "
                      "layers which is deduced from the length of `tau_arr`.")
    if isinstance(Leg_coeffs_all, DeviceInput):  # described by a few numbers per layer: expanded by a kernel (inputs.py)
        Leg_coeffs_all = Leg_coeffs_all.expand(lib, dev, T, tau)
    leg_in = T(Leg_coeffs_all)
    if leg_in.ndim == 1:
        leg_in = leg_in[None, :]
    leg = per_layer(leg_in, (None,), "The zeroth dimension of the shape of `Leg_coeffs_all` does not match the number "
                    "of layers which is deduced from the length of `tau_arr`.")
    NLeg_all = leg.shape[-1]

    hints = _hints or {}
    f_t = T(f_arr)
    if f_t.ndim == 0:
        f_t = f_t[None]
    f_nonzero = hints["f_nonzero"] if "f_nonzero" in hints else _host_any_nonzero(f_arr)
    if f_nonzero:
        f = per_layer(f_t, (), "The length of `f_arr` does not match the number of layers which is deduced from the "
                      "length of `tau_arr`.")
    else:
        f = None

    s_nonzero = hints["s_nonzero"] if "s_nonzero" in hints else _host_any_nonzero(s_poly_coeffs)
    if isinstance(s_poly_coeffs, DeviceInput):
        s_poly_coeffs = s_poly_coeffs.expand(lib, dev, T, tau)
    s_t = T(s_poly_coeffs)
    if s_t.ndim == 1:
        s_t = s_t[None, :]
    Ns = 0 if (s_t.numel() == 0 or not s_nonzero) else int(s_t.shape[-1])
    if Ns > 0:
        s_poly = per_layer(s_t, (None,), "The zeroth dimension of the shape of `s_poly_coeffs` does not match the "
                           "number of layers which is deduced from the length of `tau_arr`.")
    else:
        s_poly = None

    def per_column(x, name):
        t = T(x)
        if t.ndim == 0:
            return t[None].expand(B)
        if t.ndim == 1 and t.shape[0] == B and (batched or B == 1):
            return t
        raise ValueError(f"`{name}` must be a scalar or (batch mode) a [B] array.")

    mu0_t, I0_t, phi0_t = per_column(mu0, "mu0"), per_column(I0, "I0"), per_column(phi0, "phi0")

    # ---- structural checks (pydisort.py:223-291; value checks run on the device) ----
    if not NLeg > 0:
        raise ValueError("The number of phase function Legendre coefficients must be positive.")
    if not NLeg <= NLeg_all:
        raise ValueError("`NLeg` cannot be larger than the number of phase function Legendre coefficients provided.")
    if not NQuad >= 2:
        raise ValueError("There must be at least two streams.")
    if not NQuad % 2 == 0:
        raise ValueError("The number of streams must be even.")
    if NQuad > _lib.PD_MAX_NQUAD:
        raise ValueError(f"NQuad = {NQuad} is not supported by the CUDA kernels: at most {_lib.PD_MAX_NQUAD} streams "
                         "(one system's elimination panel has to fit the shared memory of one SM).")
    if not NFourier > 0:
        raise ValueError("The number of Fourier modes to use in the solution must be positive.")
    if not NFourier <= NLeg:
        raise ValueError("The number of Fourier modes to use in the solution must be less than or equal to the "
                         "number of phase function Legendre coefficients used.")
    if NFourier > 64 and not only_flux:
        warnings.warn("`NFourier` is large and may cause errors, consider decreasing `NFourier` to 64 and it "
                      "probably should be even less. By default `NFourier` equals `NQuad`.")
    if not NLeg <= NQuad:
        raise ValueError("There should be more streams than the number of phase function Legendre coefficients used.")
    if not use_banded_solver_NLayers >= 3:
        raise ValueError("The minimum threshold `use_banded_solver_NLayers` is 3, else the matrix will not be banded.")
    bpos = _bc_tensor(b_pos, "b_pos", B, N, NFourier, batched, T, hints.get("b_pos_zero"))
    bneg = _bc_tensor(b_neg, "b_neg", B, N, NFourier, batched, T, hints.get("b_neg_zero"))
    NFb = max(bpos.shape[1], bneg.shape[1])
    if bpos.shape[1] != NFb:
        bpos = torch.cat([bpos, torch.zeros((B, NFb - 1, N), dtype=_F64, device=dev)], dim=1)
    if bneg.shape[1] != NFb:
        bneg = torch.cat([bneg, torch.zeros((B, NFb - 1, N), dtype=_F64, device=dev)], dim=1)

    beam = hints["beam"] if "beam" in hints else _host_any_nonzero(I0)
    nt_static = bool(NT_cor) and not only_flux and NLeg < NLeg_all and f is not None
    mu_h, W_h, mu_d, w_d, ptab = _constants(NQuad, NLeg, NFourier, dev)

    # ---- BDRF tables ---------------------------------------------------------------
    modes = list(BDRF_Fourier_modes)[:NFourier]
    NBDRF = len(modes)
    flags = (_lib.PD_FLAG_BEAM if beam else 0) | (_lib.PD_FLAG_ISO if Ns > 0 else 0) | \
        (_lib.PD_FLAG_DELTA_M if f is not None else 0) | (int(_kernel_flags) & _lib.PD_FLAG_GENERIC_KERNELS)
    bdrf_q = bdrf_q0 = None
    if NBDRF:
        bdrf_q, bdrf_q0, percol = _bdrf_tables(modes, B, N, mu_h, mu0_t, beam, batched, T)
        if percol:
            flags |= _lib.PD_FLAG_BDRF_PERCOL

    cfg = _lib.pd_config(B, L, NQuad, NLeg, NLeg_all, NFourier, NBDRF, Ns, NFb, flags)

    # ---- device buffers ------------------------------------------------------------
    sol = _Solution()
    sol.lib, sol.dev, sol.cfg = lib, dev, cfg
    sol.B, sol.L, sol.N, sol.NQuad, sol.NF = B, L, N, NQuad, NFourier
    sol.batched, sol.want_torch = batched, want_torch
    sol.defer, sol.pending, sol._tau_max = _defer, [], None
    sol.tau_arr_in = tau_arr
    sol.tau = tau.contiguous()
    sol.omega = omega.contiguous()
    sol.leg_all = leg.contiguous()
    sol.f = f.contiguous() if f is not None else torch.zeros((B, L), dtype=_F64, device=dev)
    sol.mu_d, sol.w_d = mu_d, w_d
    new = lambda *shape: torch.empty(shape, dtype=_F64, device=dev)
    sol.taus, sol.omega_s, sol.wleg, sol.scale_tau = new(B, L + 1), new(B, L), new(B, L, NLeg), new(B, L)
    sol.s_s = new(B, L, Ns) if Ns > 0 else None
    sol.colp = new(B, _lib.PD_NCOLP)
    bpos_s, bneg_s = new(B, NFb, N), new(B, NFb, N)
    pmu0 = torch.zeros((B, NFourier, NLeg), dtype=_F64, device=dev)
    checks = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = _stream(dev)
    # keep every buffer handed to the library alive (and contiguous) in a local until the call returns
    sp_c = s_poly.contiguous() if s_poly is not None else None
    f_c = f.contiguous() if f is not None else None
    mu0_c, I0_c, phi0_c = mu0_t.contiguous(), I0_t.contiguous(), phi0_t.contiguous()
    bpos, bneg = bpos.contiguous(), bneg.contiguous()
    _mark("begin", dev)
    rc = lib.pd_prologue(ctypes.byref(cfg), _ptr(sol.tau), _ptr(sol.omega), _ptr(sol.leg_all),
                         _ptr(f_c), _ptr(sp_c), _ptr(mu0_c), _ptr(I0_c), _ptr(phi0_c),
                         _ptr(bpos), _ptr(bneg), _ptr(mu_d), int(bool(NT_cor)),
                         _ptr(sol.taus), _ptr(sol.omega_s), _ptr(sol.wleg), _ptr(sol.scale_tau), _ptr(sol.s_s),
                         _ptr(sol.colp), _ptr(bpos_s), _ptr(bneg_s), _ptr(pmu0), _ptr(checks), stream)
    sol._check(rc, "pd_prologue")
    _mark("prologue", dev)

    sol.K = new(B, NFourier, L, N)
    # the two distinct N x N blocks of every (column, mode, layer) item, sector-interleaved over groups of 32 items
    # (include/pydisort_b200.h, "layout of G"): a flat buffer of ceil(items / 32) * 32 items
    sol.G = new(-(-(B * NFourier * L) // 32) * 32 * 2 * N * N)
    sol.Bv = new(B, NFourier, L, NQuad) if beam else None
    sol.dth = new(B, L, Ns, NQuad) if Ns > 0 else None
    sol.C = new(B, NFourier, L, NQuad)
    # stream radiances of every mode at the L + 1 interfaces: what the output functions return at levels that are interfaces
    sol.Uif = new(B, L + 1, NFourier, NQuad)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    ws_bytes = lib.pd_workspace_bytes(ctypes.byref(cfg))
    workspace = torch.empty(max(ws_bytes // 8, 1), dtype=_F64, device=dev)
    solve_args = (_ptr(sol.taus), _ptr(sol.omega_s), _ptr(sol.wleg), _ptr(sol.s_s), _ptr(sol.colp), _ptr(bpos_s),
                  _ptr(bneg_s), _ptr(pmu0), _ptr(mu_d), _ptr(w_d), _ptr(ptab), _ptr(bdrf_q), _ptr(bdrf_q0),
                  _ptr(workspace), ws_bytes, _ptr(sol.K), _ptr(sol.G), _ptr(sol.Bv), _ptr(sol.dth), _ptr(sol.C),
                  _ptr(sol.Uif), _ptr(status), stream)
    _mark("begin", dev)
    sol._check(lib.pd_solve_stages(ctypes.byref(cfg), 1, *solve_args), "pd_solve (eigen stage)")
    _mark("solve_eigen", dev)
    sol._check(lib.pd_solve_stages(ctypes.byref(cfg), 2, *solve_args), "pd_solve (boundary-condition stage)")
    _mark("solve_bc", dev)
    del workspace
    sol.status = status
    sol.nt = nt_static
    # one read-back for the input checks and the per-column status, after everything has been enqueued
    sol._note("checks", checks)
    sol._note("status", status)
    if _defer is None:
        sol.raise_pending()

    mu_arr = np.concatenate([mu_h, -mu_h])
    if want_torch:
        # from the cached device copy of the nodes: a fresh upload from pageable host memory would make the host wait
        # for everything already queued on the stream (measured: the ensemble pipeline then runs in lock-step with
        # the GPU, one chunk at a time)
        mu_arr = torch.cat([mu_d, -mu_d])
    outs = _make_functions(sol)
    return (mu_arr,) + (outs[:3] if only_flux else outs)


def _host_any_nonzero(x):
    """Decide a structural flag from host data; a device tensor is never read back (-> "may be non-zero")."""
    if isinstance(x, DeviceInput):
        return any(_host_any_nonzero(p) for p in x.parts())
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x.numel() > 0
        x = x.detach().numpy()
    a = np.asarray(x)
    if a.size == 0:
        return False
    if a.size > 4096 and np.any(a.reshape(-1)[:4096] != 0):  # the usual case for a gigabyte-sized input: no full pass
        return True
    return bool(np.any(a != 0))


def _pending_flags(pending):
    """One int64 per queued device-side check (its maximum), as ONE device tensor (enqueued on the current stream)."""
    return torch.stack([f.reshape(-1).max().to(torch.int64) if f.numel() else torch.zeros((), dtype=torch.int64,
                                                                                           device=f.device)
                        for _, f in pending])


def _raise_for_pending(pending, flat=None):
    """Evaluate queued device-side checks with ONE device->host copy and raise / warn like the reference.  ``flat``:
    the values of ``_pending_flags(pending)`` if the caller already has them on the host (the ensemble pipeline brings
    them over on its download stream: reading them here would make the host wait for everything queued on the compute
    stream, the next ensemble's chunks included)."""
    if not pending:
        return
    if flat is None:
        flat = _pending_flags(pending).cpu().tolist()
    chk = bad = 0
    tau_range = False
    for (kind, f), v in zip(pending, flat):
        if kind == "checks":
            chk |= int(v)
        elif kind == "status":
            bad |= int(v)
        elif kind == "tau_range":
            tau_range = tau_range or bool(v)
    _raise_for_checks(chk)
    if bad:
        nbad = sum(int((f != 0).sum().item()) for kind, f in pending if kind == "status")
        ncol = sum(f.numel() for kind, f in pending if kind == "status")
        msg = []
        if bad & _lib.PD_ST_QR_NOCONV:
            msg.append("the shifted-QR eigen-solver did not converge")
        if bad & _lib.PD_ST_BAD_EIGEN:
            msg.append("a reduced eigenvalue k^2 was not positive (the reference would return NaN here)")
        if bad & _lib.PD_ST_ZERO_PIVOT:
            msg.append("an exactly singular pivot was met")
        warnings.warn(f"pydisort_b200: numerical trouble in {nbad} of {ncol} columns: " + "; ".join(msg) + ".")
    if tau_range:
        raise ValueError("tau input outside the tau range given for the atmosphere (check `tau_arr`).")


def _raise_for_checks(chk):
    C = _lib.CHK
    if chk & C["TAU_POS"]:
        raise ValueError("tau values cannot be non-positive.")
    if chk & C["THICK_POS"]:
        raise ValueError("Layer thicknesses cannot be non-positive.")
    if chk & C["OMEGA_RANGE"]:
        raise ValueError("Single-scattering albedo must be between 0 and 1, excluding 1.")
    if chk & C["LEG0_FIXED"]:
        warnings.warn("The zeroth index phase function Legendre coefficient must be, and has been corrected to, 1.")
    if chk & C["LEG_RANGE"]:
        raise ValueError("The phase function Legendre coefficients must all be between -1 and 1 exclusive (only the "
                         "zeroth coefficient can equal 1).")
    if chk & C["I0_NEG"]:
        raise ValueError("The intensity of the incident beam cannot be negative.")
    if chk & C["MU0_RANGE"]:
        raise ValueError("The cosine of the polar angle of the incident beam must be between 0 and 1, excluding 0.")
    if chk & C["PHI0_RANGE"]:
        raise ValueError("Provide the principal azimuthal angle for the incident beam (must be between 0 and 2pi, "
                         "excluding 2pi).")
    if chk & C["F_RANGE"]:
        raise ValueError("The fractional scattering must be between 0 and 1.")
    if chk & C["MU0_AT_NODE"]:
        raise ValueError("Some quadrature angles come too close to `mu0`. Perturb `NQuad` or `mu0` to rectify this "
                         "error.")
    if chk & C["OMEGA_NEAR1"]:
        warnings.warn("Some delta-scaled single-scattering albedos are very close to 1 which may cause numerical "
                      "instability.")
    if chk & C["LEG_NEAR1"]:
        warnings.warn("Some delta-scaled phase function Legendre coefficients have a magnitude that is very close "
                      "to 1 (this excludes the zeroth index coefficient which must be 1) which may cause numerical "
                      "instability.")


def _make_functions(sol):
    """The output functions, with the reference's signatures
    (_assemble_intensity_and_fluxes.py:170,334,446,527; pydisort.py:643)."""

    def flux_up(tau, is_antiderivative_wrt_tau=False, return_tau_arr=False):
        out, ntau = sol.eval_flux(tau, is_antiderivative_wrt_tau)
        res = sol._finish(out[0], (ntau,))
        return (res, sol.tau_arr_in) if return_tau_arr else res

    def flux_down(tau, is_antiderivative_wrt_tau=False, return_tau_arr=False):
        out, ntau = sol.eval_flux(tau, is_antiderivative_wrt_tau)
        res = (sol._finish(out[1], (ntau,)), sol._finish(out[2], (ntau,)))
        return res + (sol.tau_arr_in,) if return_tau_arr else res

    def u0(tau, is_antiderivative_wrt_tau=False, return_tau_arr=False, _return_act_dscale_for_reclass=False):
        val, recl, ntau = sol.eval_u0(tau, is_antiderivative_wrt_tau, _return_act_dscale_for_reclass)
        outs = (sol._finish(val, (sol.NQuad, ntau)),)
        if return_tau_arr:
            outs += (sol.tau_arr_in,)
        if _return_act_dscale_for_reclass:
            outs += (sol._finish(recl, (ntau,)),)
        return outs[0] if len(outs) == 1 else outs

    def u(tau, phi, is_antiderivative_wrt_tau=False, return_Fourier_error=False, return_tau_arr=False):
        val, ulast, ph, ntau, nphi = sol.eval_u(tau, phi, is_antiderivative_wrt_tau, sol.nt, False)
        outs = (sol._finish(val, (sol.NQuad, ntau, nphi)),)
        if return_Fourier_error:
            # Cauchy criterion on the un-corrected, un-rescaled series (:265-316)
            plain, ulast, _, _, _ = sol.eval_u(tau, phi, is_antiderivative_wrt_tau, False, True)
            resc = sol.colp[:, _lib.PD_COL_RESCALE][:, None, None, None]
            plain = torch.where(resc != 0, plain / resc, torch.zeros_like(plain)).abs()
            phi0 = sol.colp[:, _lib.PD_COL_PHI0][:, None]
            last = (ulast[:, :, :, None] * torch.cos((sol.NF - 1) * (phi0 - ph[None, :]))[:, None, None, :]).abs()
            ratio = torch.where(plain > 1e-8, last / plain, torch.zeros_like(plain))
            err = ratio.reshape(sol.B, -1).max(dim=1).values
            if not sol.batched:
                err = err[0]
            outs += (err if sol.want_torch else (err.cpu().numpy()[()]),)
        if return_tau_arr:
            outs += (sol.tau_arr_in,)
        return outs[0] if len(outs) == 1 else outs

    def u_at_mu(mu, tau, phi, is_antiderivative_wrt_tau=False):
        """u interpolated to user polar angles on the device (used by ``subroutines.interpolate``)."""
        val, _, _, ntau, nphi = sol.eval_u(tau, phi, is_antiderivative_wrt_tau, sol.nt, False)
        out, nmu = sol.interp_mu(val, mu)
        return sol._finish(out, (nmu, ntau, nphi))

    def u0_at_mu(mu, tau, is_antiderivative_wrt_tau=False):
        val, _, ntau = sol.eval_u0(tau, is_antiderivative_wrt_tau, False)
        out, nmu = sol.interp_mu(val, mu)
        return sol._finish(out, (nmu, ntau))

    u.at_mu = u_at_mu
    u0.at_mu = u0_at_mu
    for fn, kind in ((flux_up, "flux_up"), (flux_down, "flux_down"), (u0, "u0"), (u, "u")):
        fn.kind = kind
        fn.batched = sol.batched
        fn.solution = sol
    return flux_up, flux_down, u0, u
