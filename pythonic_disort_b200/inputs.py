"""Inputs of ``pydisort()`` given by a compact description and expanded on the device.

The reference takes ``Leg_coeffs_all`` [NLayers, NLeg_all] and ``s_poly_coeffs`` [NLayers, Nscoeffs] as arrays
(src/PythonicDISORT/pydisort.py:13-29); its users usually build them from a few numbers per layer --
``g ** np.arange(NLeg_all)`` for Henyey-Greenstein layers (pydisotest/4_test.py:27, 5_test.py:24),
``generate_s_poly_coeffs`` (subroutines.py:413-454) for a thermal source that is linear in tau between level values.
For an ensemble in host memory those arrays are most of the bytes that cross the host -> device link (LW ensemble:
660 of 844 doubles per column).  The objects below stand for such an input in a ``pydisort()`` / ``solve_ensemble()``
call; only their description is uploaded and a kernel of the library writes the array the solver reads
(``pd_hg_moments`` / ``pd_level_source``, include/pydisort_b200.h).  Passing the expanded arrays instead gives the same
results to the rounding of the moments (CUDA's ``pow`` is within 2 ulp of NumPy's).
"""
import numpy as np
import torch


class DeviceInput:
    """Base class: ``parts`` are the arrays of the description (NumPy or torch, with or without the column axis)."""

    argument = None  # name of the pydisort() argument the object stands for

    def parts(self):
        raise NotImplementedError

    def with_parts(self, parts):
        raise NotImplementedError

    def batched(self, B):
        """True if the description carries the leading column axis of a batch of B columns."""
        raise NotImplementedError

    def expand(self, lib, dev, to_device, tau):
        """The array the solver reads, as a device tensor (``tau``: the call's [B, L] optical depths on the device)."""
        raise NotImplementedError

    def __getitem__(self, sl):
        return self.with_parts(tuple(p[sl] for p in self.parts()))

    def on_device(self):
        return any(isinstance(p, torch.Tensor) and p.is_cuda for p in self.parts())


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"libpydisort_b200: {what} failed with code {rc}")


class HenyeyGreenstein(DeviceInput):
    """``Leg_coeffs_all`` of Henyey-Greenstein layers: ``g`` [NLayers] or [B, NLayers] asymmetry parameters ->
    moments ``g ** l``, l < ``NLeg_all``."""

    argument = "Leg_coeffs_all"

    def __init__(self, g, NLeg_all):
        self.g = g if isinstance(g, torch.Tensor) else np.asarray(g, dtype=np.float64)
        self.NLeg_all = int(NLeg_all)
        if self.NLeg_all < 1 or self.g.ndim not in (1, 2):
            raise ValueError("HenyeyGreenstein needs g of shape [NLayers] or [B, NLayers] and NLeg_all >= 1.")

    def parts(self):
        return (self.g,)

    def with_parts(self, parts):
        return HenyeyGreenstein(parts[0], self.NLeg_all)

    def batched(self, B):
        return self.g.ndim == 2 and self.g.shape[0] == B

    def expand(self, lib, dev, to_device, tau):
        g = to_device(self.g).contiguous()
        out = torch.empty(tuple(g.shape) + (self.NLeg_all,), dtype=torch.float64, device=dev)
        if g.numel():
            stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
            _check(lib.pd_hg_moments(g.numel(), self.NLeg_all, g.data_ptr(), out.data_ptr(), stream), "pd_hg_moments")
        return out


class LevelSource(DeviceInput):
    """``s_poly_coeffs`` of a thermal source that is linear in tau inside every layer: ``levels`` [NLayers + 1] or
    [B, NLayers + 1], the source (band-integrated emission) at the level optical depths 0, tau_0, .., tau_{L-1}."""

    argument = "s_poly_coeffs"

    def __init__(self, levels):
        self.levels = levels if isinstance(levels, torch.Tensor) else np.asarray(levels, dtype=np.float64)
        if self.levels.ndim not in (1, 2):
            raise ValueError("LevelSource needs levels of shape [NLayers + 1] or [B, NLayers + 1].")

    def parts(self):
        return (self.levels,)

    def with_parts(self, parts):
        return LevelSource(parts[0])

    def batched(self, B):
        return self.levels.ndim == 2 and self.levels.shape[0] == B

    def expand(self, lib, dev, to_device, tau):
        B, L = tau.shape
        lev = to_device(self.levels)
        single = lev.ndim == 1
        if lev.shape[-1] != L + 1 or (not single and lev.shape[0] != B):
            raise ValueError("Missing source specification at some boundaries / interfaces.")
        if single:
            lev = lev[None].expand(B, L + 1)
        lev, tau = lev.contiguous(), tau.contiguous()
        out = torch.empty((B, L, 2), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
        _check(lib.pd_level_source(B, L, tau.data_ptr(), lev.data_ptr(), out.data_ptr(), stream), "pd_level_source")
        return out[0] if (single and B == 1) else out
