"""ctypes binding of the C ABI in include/planedepth_b200.h (the only way Python reaches the CUDA path).

There is deliberately no CPU / PyTorch fallback: if the shared library is missing or a call fails the
error is raised, so a silent eager path can never stand in for the kernels."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("PLANEDEPTH_B200_LIB") or os.path.join(CSRC, "libplanedepth_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
BUILD_DIR = os.path.join(CSRC, "build")  # git-ignored object files (one per translation unit)

PD_WARP_DISP, PD_WARP_HOMOGRAPHY, PD_WARP_DEPTH = 0, 1, 2
PD_LOSS_L1, PD_LOSS_MIXTURE, PD_LOSS_SSIM_L1 = 0, 1, 2
PD_MASK_NONE, PD_MASK_F32, PD_MASK_U8 = 0, 1, 2
PD_FLAG_EXACT_COORDS = 1
PD_FLAG_NO_MASK_SUMMARY = 2
PD_FLAG_ACCUMULATE = 4
PD_FLAG_WORKSPACE_READY = 8
ABI_VERSION = 9  # PD_ABI_VERSION in include/planedepth_b200.h
PD_STATS_PLAIN, PD_STATS_MIXTURE = 2, 4
PD_DTYPE_F32, PD_DTYPE_BF16 = 0, 1

EXPORTS = [
    "pd_version", "pd_last_error", "pd_launch_count", "pd_reset_launch_count",
    "pd_warp_composite_workspace_bytes", "pd_warp_composite_stats_bytes", "pd_warp_composite_fwd", "pd_warp_composite_bwd",
    "pd_photometric_workspace_bytes", "pd_photometric_fwd", "pd_photometric_bwd", "pd_debug_roundtrip",
    "pd_occlusion_masks_workspace_bytes", "pd_occlusion_masks_fwd",
    "pd_smooth_loss_workspace_bytes", "pd_smooth_loss_fwd", "pd_smooth_loss_bwd",
    "pd_plane_tail_fwd", "pd_plane_tail_bwd",
    "pd_get_tuning", "pd_set_tuning", "pd_x_constant_check", "pd_resize_bicubic_u8", "pd_warp_composite_supports",
]


class Strides4(C.Structure):
    _fields_ = [("b", C.c_int64), ("n", C.c_int64), ("y", C.c_int64), ("x", C.c_int64)]


class WarpDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("warp_type", C.c_int32), ("mixture", C.c_int32), ("automask", C.c_int32), ("mask_dtype", C.c_int32),
        ("disp_sign", C.c_float), ("flags", C.c_int32), ("dtype", C.c_int32), ("reserved0", C.c_int32),
        ("disp_stride", Strides4), ("mask_stride", Strides4),
    ]


class WarpIn(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("src", "tgt", "logits", "sigma", "disp", "mask", "hmat", "cam")]


class WarpOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "rgb_rec", "stats", "nll", "nll_auto", "rgb_rec_layered", "logit_rec", "probability_rec", "sigma_rec", "pi_rec")]


class WarpGradOut(C.Structure):
    _fields_ = [("g_rgb_rec", C.c_void_p), ("g_nll", C.c_void_p), ("g_ph_sum", C.c_void_p), ("ph_scale", C.c_float),
                ("g_unit", C.c_void_p), ("g_unit_nll", C.c_void_p), ("g_pred", C.c_void_p), ("mask_novel", C.c_void_p)]


class Tuning(C.Structure):
    _fields_ = [("stream_ctas_per_sm", C.c_int32), ("stream_hs", C.c_int32), ("stream_nst", C.c_int32), ("stream_smem_kb", C.c_int32),
                ("stream_px8", C.c_int32), ("ssim_tiles", C.c_int32), ("homo_tiles", C.c_int32), ("stream_fwd_minb", C.c_int32),
                ("stream_no_l2_hint", C.c_int32), ("stream_bwd_minb", C.c_int32), ("tail_direct", C.c_int32), ("reserved", C.c_int32 * 5)]


class WarpGradIn(C.Structure):
    _fields_ = [("g_logits", C.c_void_p), ("g_sigma", C.c_void_p), ("g_disp", C.c_void_p),
                ("g_disp_stride", Strides4), ("g_hmat", C.c_void_p)]


class TailDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("mixture", C.c_int32), ("mask_dtype", C.c_int32),
                ("disp_stride", Strides4), ("mask_stride", Strides4)]


class TailIn(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("logits_raw", "sigma_raw", "disp_layered", "mask")]


class TailOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("logits", "sigma", "probability", "pi", "disp", "depth", "stats")]


class TailGradOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("g_logits", "g_sigma", "g_probability", "g_disp", "g_depth")]


class TailGradIn(C.Structure):
    _fields_ = [("g_logits_raw", C.c_void_p), ("g_sigma_raw", C.c_void_p), ("g_disp_layered", C.c_void_p), ("g_disp_stride", Strides4)]


class SmoothDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("x0", C.c_int32), ("gamma", C.c_float)]


class OcclDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("flags", C.c_int32), ("disp_stride", Strides4)]


class OcclIn(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("logits", "probability", "disp_layered", "disp")]


class OcclOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("o_l", "o_fr", "mask_novel", "disp_pp")]


class ResizeDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("B", "Hs", "Ws", "Hf", "Wf", "y0", "x0", "H", "W", "src_layout")]


class LossDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("loss_mode", C.c_int32),
                ("automask", C.c_int32), ("has_mask_novel", C.c_int32), ("out_scale", C.c_float)]


class LossIn(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("rgb_rec", "tgt", "src", "mask_novel", "nll", "nll_auto")]


class LossOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("pred", "ph_map", "ph_sum", "g_unit", "g_unit_nll")]


class LossGradOut(C.Structure):
    _fields_ = [("g_ph_sum", C.c_void_p), ("g_pred", C.c_void_p)]


class LossGradIn(C.Structure):
    _fields_ = [("g_rgb_rec", C.c_void_p), ("g_nll", C.c_void_p)]


class PlaneDepthLibraryError(RuntimeError):
    pass


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _includes(path, seen=None):
    """Transitive closure of the quoted #include s of one source file (the headers its object depends on)."""
    seen = set() if seen is None else seen
    try:
        text = open(path).read()
    except OSError:
        return seen
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("#include \""):
            dep = os.path.normpath(os.path.join(os.path.dirname(path), line.split('"')[1]))
            if dep not in seen and os.path.exists(dep):
                seen.add(dep)
                _includes(dep, seen)
    return seen


def _digest(src):
    """Content hash of one translation unit and every header it includes (mtimes do not survive a snapshot copy)."""
    import hashlib

    h = hashlib.sha1()
    for path in [src] + sorted(_includes(src)):
        h.update(os.path.basename(path).encode())
        h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _manifest_path():
    return os.path.join(BUILD_DIR, "manifest.json")


def _load_manifest():
    import json

    try:
        return json.load(open(_manifest_path()))
    except (OSError, ValueError):
        return {}


class _BuildLock:
    """Exclusive file lock around the build: one process per GPU means every rank reaches lib() at the same time; one
    builds, the others wait and then find a fresh library."""

    def __enter__(self):
        import fcntl

        os.makedirs(BUILD_DIR, exist_ok=True)
        self.f = open(os.path.join(BUILD_DIR, ".lock"), "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl

        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def build_library(verbose: bool = False, force: bool = False) -> str:
    """Compile every translation unit of csrc/ (pd_abi.cu + one pd_tu_*.cu per kernel family) for sm_100a, in parallel,
    and link libplanedepth_b200.so.  Objects whose sources / headers did not change are reused.  nvcc cross-compiles
    without a GPU.  The library is linked to a temporary name and renamed into place, under a file lock."""
    from concurrent.futures import ThreadPoolExecutor

    import json

    with _BuildLock():
        old = {} if force else _load_manifest()
        new, jobs, objs = {}, [], []
        for src in _sources():
            name = os.path.basename(src)
            obj = os.path.join(BUILD_DIR, name[:-3] + ".o")
            objs.append(obj)
            new[name] = _digest(src)
            if old.get(name) != new[name] or not os.path.exists(obj):
                jobs.append(["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-c", "-o", obj, src])

        def run(cmd):
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise PlaneDepthLibraryError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
            return res.stderr

        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
            for err in pool.map(run, jobs):
                if verbose:
                    sys.stderr.write(err)
        if jobs or not os.path.exists(LIB_PATH) or old.get("__lib__") != "linked":
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs)
            os.replace(tmp, LIB_PATH)
        new["__lib__"] = "linked"
        with open(_manifest_path() + ".tmp", "w") as f:
            json.dump(new, f, indent=1, sort_keys=True)
        os.replace(_manifest_path() + ".tmp", _manifest_path())
    return LIB_PATH


def _needs_rebuild() -> bool:
    if os.environ.get("PLANEDEPTH_B200_LIB"):
        return False  # explicit library (kernel-variant experiments): use as is
    if not os.path.exists(LIB_PATH):
        return True
    man = _load_manifest()
    if man.get("__lib__") != "linked":
        return True
    return any(man.get(os.path.basename(src)) != _digest(src) for src in _sources())


_lib = None


def lib() -> C.CDLL:
    """Load (building first if sources are newer and nvcc exists) and type the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    if _needs_rebuild():
        from shutil import which

        if which("nvcc") is not None:
            build_library()
        elif not os.path.exists(LIB_PATH):
            raise PlaneDepthLibraryError(
                "planedepth_b200: %s is missing and nvcc is not on PATH; run __graft_entry__.build()" % LIB_PATH)
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise PlaneDepthLibraryError("cannot load %s: %s" % (LIB_PATH, e))
    L.pd_version.restype = C.c_int
    L.pd_last_error.restype = C.c_char_p
    L.pd_launch_count.restype = C.c_int64
    L.pd_reset_launch_count.restype = None
    L.pd_warp_composite_workspace_bytes.restype = C.c_size_t
    L.pd_warp_composite_workspace_bytes.argtypes = [C.POINTER(WarpDesc)]
    L.pd_warp_composite_stats_bytes.restype = C.c_size_t
    L.pd_warp_composite_stats_bytes.argtypes = [C.POINTER(WarpDesc)]
    L.pd_warp_composite_supports.restype = C.c_int
    L.pd_warp_composite_supports.argtypes = [C.POINTER(WarpDesc), C.POINTER(WarpIn)]
    L.pd_warp_composite_fwd.restype = C.c_int
    L.pd_warp_composite_fwd.argtypes = [C.POINTER(WarpDesc), C.POINTER(WarpIn), C.POINTER(WarpOut), C.c_void_p, C.c_void_p]
    L.pd_warp_composite_bwd.restype = C.c_int
    L.pd_warp_composite_bwd.argtypes = [C.POINTER(WarpDesc), C.POINTER(WarpIn), C.POINTER(WarpOut), C.POINTER(WarpGradOut),
                                        C.POINTER(WarpGradIn), C.c_void_p, C.c_void_p]
    L.pd_photometric_workspace_bytes.restype = C.c_size_t
    L.pd_photometric_workspace_bytes.argtypes = [C.POINTER(LossDesc)]
    L.pd_photometric_fwd.restype = C.c_int
    L.pd_photometric_fwd.argtypes = [C.POINTER(LossDesc), C.POINTER(LossIn), C.POINTER(LossOut), C.c_void_p, C.c_void_p]
    L.pd_photometric_bwd.restype = C.c_int
    L.pd_photometric_bwd.argtypes = [C.POINTER(LossDesc), C.POINTER(LossIn), C.POINTER(LossOut), C.POINTER(LossGradOut), C.POINTER(LossGradIn),
                                     C.c_void_p, C.c_void_p]
    L.pd_plane_tail_fwd.restype = C.c_int
    L.pd_plane_tail_fwd.argtypes = [C.POINTER(TailDesc), C.POINTER(TailIn), C.POINTER(TailOut), C.c_void_p]
    L.pd_plane_tail_bwd.restype = C.c_int
    L.pd_plane_tail_bwd.argtypes = [C.POINTER(TailDesc), C.POINTER(TailIn), C.POINTER(TailOut), C.POINTER(TailGradOut), C.POINTER(TailGradIn), C.c_void_p]
    L.pd_smooth_loss_workspace_bytes.restype = C.c_size_t
    L.pd_smooth_loss_workspace_bytes.argtypes = [C.POINTER(SmoothDesc)]
    L.pd_smooth_loss_fwd.restype = C.c_int
    L.pd_smooth_loss_fwd.argtypes = [C.POINTER(SmoothDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pd_smooth_loss_bwd.restype = C.c_int
    L.pd_smooth_loss_bwd.argtypes = [C.POINTER(SmoothDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pd_occlusion_masks_workspace_bytes.restype = C.c_size_t
    L.pd_occlusion_masks_workspace_bytes.argtypes = [C.POINTER(OcclDesc)]
    L.pd_occlusion_masks_fwd.restype = C.c_int
    L.pd_occlusion_masks_fwd.argtypes = [C.POINTER(OcclDesc), C.POINTER(OcclIn), C.POINTER(OcclOut), C.c_void_p, C.c_void_p]
    L.pd_get_tuning.restype = None
    L.pd_get_tuning.argtypes = [C.POINTER(Tuning)]
    L.pd_set_tuning.restype = None
    L.pd_set_tuning.argtypes = [C.POINTER(Tuning)]
    L.pd_x_constant_check.restype = C.c_int
    L.pd_x_constant_check.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Strides4), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.pd_resize_bicubic_u8.restype = C.c_int
    L.pd_resize_bicubic_u8.argtypes = [C.POINTER(ResizeDesc), C.c_void_p, C.c_void_p, C.c_void_p]
    L.pd_debug_roundtrip.restype = C.c_int
    L.pd_debug_roundtrip.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    if L.pd_version() != ABI_VERSION:
        raise PlaneDepthLibraryError("ABI version mismatch: library %d, binding %d" % (L.pd_version(), ABI_VERSION))
    _lib = L
    return L


class tuned:
    """Context manager for tests / experiments: ``with tuned(stream_ctas_per_sm=1): ...`` (pd_set_tuning)."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        L = lib()
        self.old = Tuning()
        L.pd_get_tuning(C.byref(self.old))
        new = Tuning()
        C.memmove(C.byref(new), C.byref(self.old), C.sizeof(Tuning))
        for k, v in self.kw.items():
            setattr(new, k, int(v))
        L.pd_set_tuning(C.byref(new))
        return self

    def __exit__(self, *exc):
        lib().pd_set_tuning(C.byref(self.old))


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().pd_last_error().decode("utf-8", "replace")
        raise PlaneDepthLibraryError("%s failed (pd_status %d): %s" % (what, rc, msg))
