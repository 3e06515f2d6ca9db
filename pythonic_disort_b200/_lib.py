"""ctypes binding of libpydisort_b200.so (C ABI: include/pydisort_b200.h) and
the in-tree build recipe (nvcc, sm_100a only)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("PD_LIB_PATH") or os.path.join(HERE, "libpydisort_b200.so")  # override: tuning experiments only
SOURCES = ["pd_api.cu", "pd_kernel_a.cu", "pd_kernel_b.cu", "pd_kernel_eval.cu"]


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + \
        [os.path.join(HERE, "..", "include", "pydisort_b200.h")]


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"] + os.environ.get("PD_NVCC_EXTRA", "").split()  # -D tuning switches for experiments
BUILD_DIR = os.environ.get("PD_BUILD_DIR") or os.path.join(HERE, "build")


class pd_config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("B", "L", "NQuad", "NLeg", "NLeg_all", "NFourier", "NBDRF", "Nscoeffs", "NFb", "flags")]


class pd_state(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("tau", "taus", "scale_tau", "colp", "K", "G", "Bv", "dth", "C", "mu_nodes", "w_nodes", "Uif")]


# constants of include/pydisort_b200.h
PD_FLAG_BEAM, PD_FLAG_ISO, PD_FLAG_DELTA_M, PD_FLAG_BDRF_PERCOL, PD_FLAG_GENERIC_KERNELS = 1, 2, 4, 8, 256
PD_ST_QR_NOCONV, PD_ST_BAD_EIGEN, PD_ST_ZERO_PIVOT = 1, 2, 4
PD_NCOLP = 8
PD_MAX_NQUAD = 92
PD_COL_MU0, PD_COL_I0, PD_COL_RESCALE, PD_COL_PHI0, PD_COL_I0_RAW, PD_COL_DM, PD_COL_NT = 0, 1, 2, 3, 4, 5, 6
CHK = dict(TAU_POS=1 << 0, THICK_POS=1 << 1, OMEGA_RANGE=1 << 2, LEG_RANGE=1 << 3, I0_NEG=1 << 4, MU0_RANGE=1 << 5,
           PHI0_RANGE=1 << 6, F_RANGE=1 << 7, LEG0_FIXED=1 << 8, OMEGA_NEAR1=1 << 9, LEG_NEAR1=1 << 10,
           MU0_AT_NODE=1 << 11)

EXPORTS = ["pd_abi_version", "pd_workspace_bytes", "pd_prologue", "pd_solve", "pd_solve_stages", "pd_eval_flux", "pd_eval_u0",
           "pd_eval_u", "pd_interp_mu", "pd_planck_band", "pd_s_poly_coeffs", "pd_hg_moments", "pd_level_source", "pd_hapke_modes",
           "pd_fp64_probe"]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in [os.path.join(CSRC, x) for x in SOURCES] + _headers())


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    extra = ["-Xptxas=-v"] if verbose else []
    procs = []
    for src in SOURCES:  # one nvcc per translation unit, in parallel
        obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        cmd = ["nvcc"] + NVCC_FLAGS + extra + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    for src, obj, proc in procs:
        out = proc.communicate()[0]
        log.append(out)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        objs.append(obj)
    res = subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(log))
    return LIB_PATH


def bind(path):
    """Load a library exporting the C ABI and declare argument types."""
    lib = ctypes.CDLL(path)
    vp, ci, cfgp, stp = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(pd_config), ctypes.POINTER(pd_state)
    lib.pd_abi_version.restype = ci
    lib.pd_workspace_bytes.restype = ctypes.c_size_t
    lib.pd_workspace_bytes.argtypes = [cfgp]
    lib.pd_prologue.restype = ci
    lib.pd_prologue.argtypes = [cfgp] + [vp] * 11 + [ci] + [vp] * 10 + [vp]
    lib.pd_solve.restype = ci
    lib.pd_solve.argtypes = [cfgp] + [vp] * 13 + [vp, ctypes.c_size_t] + [vp] * 7 + [vp]
    lib.pd_solve_stages.restype = ci
    lib.pd_solve_stages.argtypes = [cfgp, ci] + [vp] * 13 + [vp, ctypes.c_size_t] + [vp] * 7 + [vp]
    lib.pd_eval_flux.restype = ci
    lib.pd_eval_flux.argtypes = [cfgp, stp, vp, ci, ci, vp, vp, vp, vp]
    lib.pd_eval_u0.restype = ci
    lib.pd_eval_u0.argtypes = [cfgp, stp, vp, ci, ci, vp, vp, vp]
    lib.pd_eval_u.restype = ci
    lib.pd_eval_u.argtypes = [cfgp, stp, vp, ci, vp, ci, ci, ci] + [vp] * 5 + [vp, vp, vp]
    lib.pd_interp_mu.restype = ci
    lib.pd_interp_mu.argtypes = [ci, ci, ctypes.c_long, ci, vp, vp, vp, vp]
    lib.pd_planck_band.restype = ci
    lib.pd_planck_band.argtypes = [ctypes.c_long, vp, ctypes.c_double, ctypes.c_double, vp, vp, vp]
    lib.pd_s_poly_coeffs.restype = ci
    lib.pd_s_poly_coeffs.argtypes = [ci, ci, vp, vp, ctypes.c_double, ctypes.c_double, vp, vp, vp]
    lib.pd_hg_moments.restype = ci
    lib.pd_hg_moments.argtypes = [ctypes.c_long, ci, vp, vp, vp]
    lib.pd_level_source.restype = ci
    lib.pd_level_source.argtypes = [ci, ci, vp, vp, vp, vp]
    lib.pd_hapke_modes.restype = ci
    lib.pd_hapke_modes.argtypes = [ci, ctypes.c_long, ci, ci, vp, vp, vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp, vp]
    lib.pd_fp64_probe.restype = ctypes.c_double
    lib.pd_fp64_probe.argtypes = [vp, ci, vp]
    if lib.pd_abi_version() != 2:
        raise RuntimeError("libpydisort_b200 ABI mismatch")
    return lib


_cuda_lib = None


def cuda_lib():
    """The CUDA library; there is no other implementation to fall back to."""
    global _cuda_lib
    if _cuda_lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(pythonic_disort_b200 has no CPU or PyTorch fallback)")
        _cuda_lib = bind(LIB_PATH)
    return _cuda_lib
