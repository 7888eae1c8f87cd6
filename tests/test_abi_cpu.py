"""CPU: the C-ABI shared library builds for sm_100a, loads without a GPU, exports every symbol that
include/planedepth_b200.h declares, and the ctypes mirror of each struct has the C layout."""
import ctypes as C
import os
import re
import subprocess

import pytest

from planedepth_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "planedepth_b200.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z_0-9]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    lib = L.lib()
    names = declared_functions()
    assert set(L.EXPORTS) <= set(names)
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert lib.pd_version() == L.ABI_VERSION
    assert lib.pd_last_error() is not None


def test_ctypes_structs_match_c_layout(tmp_path):
    structs = {
        "pd_strides4": L.Strides4, "pd_warp_desc": L.WarpDesc, "pd_warp_in": L.WarpIn, "pd_warp_out": L.WarpOut,
        "pd_warp_grad_out": L.WarpGradOut, "pd_warp_grad_in": L.WarpGradIn, "pd_loss_desc": L.LossDesc, "pd_loss_in": L.LossIn,
        "pd_loss_out": L.LossOut, "pd_loss_grad_out": L.LossGradOut, "pd_loss_grad_in": L.LossGradIn,
        "pd_occl_desc": L.OcclDesc, "pd_occl_in": L.OcclIn, "pd_occl_out": L.OcclOut, "pd_smooth_desc": L.SmoothDesc,
        "pd_tail_desc": L.TailDesc, "pd_tail_in": L.TailIn, "pd_tail_out": L.TailOut, "pd_tail_grad_out": L.TailGradOut,
        "pd_tail_grad_in": L.TailGradIn, "pd_tuning": L.Tuning,
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "planedepth_b200.h"', "int main(void){"]
    for cname, st in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f, _ in st._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, st in structs.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for f, _ in st._fields_:
            assert int(got["%s.%s" % (cname, f)]) == getattr(st, f).offset, "%s.%s" % (cname, f)


def test_argument_validation_needs_no_gpu():
    lib = L.lib()
    d = L.WarpDesc(B=1, N=1, H=1, W=8)
    rc = lib.pd_warp_composite_fwd(C.byref(d), C.byref(L.WarpIn()), C.byref(L.WarpOut()), None, None)
    assert rc == 2 and b"H,W >= 2" in lib.pd_last_error()
    d = L.WarpDesc(B=1, N=1, H=8, W=8)
    rc = lib.pd_warp_composite_fwd(C.byref(d), C.byref(L.WarpIn()), C.byref(L.WarpOut()), None, None)
    assert rc == 1 and b"NULL" in lib.pd_last_error()
    with pytest.raises(L.PlaneDepthLibraryError):
        L.check(rc, "pd_warp_composite_fwd")
    ld = L.LossDesc(B=1, H=8, W=8, loss_mode=7)
    assert lib.pd_photometric_fwd(C.byref(ld), C.byref(L.LossIn()), C.byref(L.LossOut()), None, None) == 1


def test_product_path_refuses_cpu_tensors():
    import torch

    from planedepth_b200.functional import WarpConfig, warp_composite

    cfg = WarpConfig(L.PD_WARP_DISP, False, False, 1.0, (1, 2, 8, 8))
    with pytest.raises(L.PlaneDepthLibraryError):
        warp_composite(cfg, torch.zeros(1, 3, 8, 8), None, torch.zeros(1, 2, 8, 8), None, torch.ones(1, 2, 8, 8), None)


def test_argument_validation_returns_codes_not_crashes():
    """Error behaviour of the C ABI that needs no device: NULL / inconsistent arguments come back as pd_status codes with a
    message in pd_last_error(); nothing aborts or throws across the boundary."""
    lib = L.lib()
    PD_ERR_ARG, PD_ERR_SHAPE = 1, 2
    assert lib.pd_warp_composite_fwd(None, None, None, None, None) == PD_ERR_ARG
    assert b"NULL" in lib.pd_last_error()
    d = L.WarpDesc(B=1, N=300, H=8, W=8, warp_type=L.PD_WARP_DISP)
    tin = L.WarpIn(src=1, logits=1, disp=1)  # non-NULL dummies: validation stops before anything is dereferenced
    assert lib.pd_warp_composite_fwd(C.byref(d), C.byref(tin), None, None, None) == PD_ERR_SHAPE
    assert b"PD_MAX_PLANES" in lib.pd_last_error()
    d = L.WarpDesc(B=1, N=4, H=8, W=8, warp_type=L.PD_WARP_HOMOGRAPHY)
    assert lib.pd_warp_composite_fwd(C.byref(d), C.byref(L.WarpIn(src=1, logits=1)), None, None, None) == PD_ERR_ARG
    assert b"hmat" in lib.pd_last_error()
    sd = L.SmoothDesc(B=1, H=1, W=8, x0=0, gamma=1.0)
    assert lib.pd_smooth_loss_fwd(C.byref(sd), 1, 1, 1, 1, None) == PD_ERR_SHAPE
    od = L.OcclDesc(B=1, N=2, H=4, W=4)
    assert lib.pd_occlusion_masks_fwd(C.byref(od), C.byref(L.OcclIn(logits=1, disp_layered=1)), C.byref(L.OcclOut()), None, None) == PD_ERR_ARG
    td = L.TailDesc(B=1, N=2, H=4, W=4, mixture=1)
    assert lib.pd_plane_tail_fwd(C.byref(td), C.byref(L.TailIn(logits_raw=1, disp_layered=1)), C.byref(L.TailOut()), None) == PD_ERR_ARG
    # sizes are pure functions of the descriptor
    wd = L.WarpDesc(B=2, N=5, H=16, W=32, warp_type=L.PD_WARP_HOMOGRAPHY)
    assert lib.pd_warp_composite_workspace_bytes(C.byref(wd)) == 2 * 16 * 32 * 16
    wd.warp_type = L.PD_WARP_DISP
    assert lib.pd_warp_composite_workspace_bytes(C.byref(wd)) == 0
    assert lib.pd_warp_composite_stats_bytes(C.byref(wd)) == 2 * 2 * 16 * 32 * 4 + 2 * 16 * 8
    assert lib.pd_occlusion_masks_workspace_bytes(C.byref(od)) == 1 * 2 * 4 * 4 * 4


def test_tuning_block_is_clamped_and_restorable():
    """pd_set_tuning replaces the per-call getenv knobs of round 1: values are clamped when they are set (a zero or negative
    ring depth can no longer reach a division), NULL restores what the environment said at load."""
    lib = L.lib()
    t = L.Tuning(stream_ctas_per_sm=-3, stream_hs=-1, stream_nst=99, stream_smem_kb=10 ** 6, stream_px8=7, ssim_tiles=-2, homo_tiles=5, stream_fwd_minb=7)
    lib.pd_set_tuning(C.byref(t))
    got = L.Tuning()
    lib.pd_get_tuning(C.byref(got))
    assert (got.stream_ctas_per_sm, got.stream_hs, got.stream_nst, got.stream_smem_kb, got.stream_px8, got.ssim_tiles, got.homo_tiles) == (0, 0, 8, 220, 1, 1, 1)
    lib.pd_set_tuning(None)
    lib.pd_get_tuning(C.byref(got))
    assert got.stream_nst == int(os.environ.get("PD_STREAM_NST", "0") or 0)
    with L.tuned(stream_ctas_per_sm=1):
        lib.pd_get_tuning(C.byref(got))
        assert got.stream_ctas_per_sm == 1
    lib.pd_get_tuning(C.byref(got))
    assert got.stream_ctas_per_sm == int(os.environ.get("PD_STREAM_CTAS", "0") or 0)


def test_no_per_call_environment_reads():
    """kernel selection must not depend on getenv() at call time.  The only getenv in the library sits in the
    load-time initialiser of the tuning block."""
    csrc = os.path.join(ROOT, "planedepth_b200", "csrc")
    hits = []
    for f in sorted(os.listdir(csrc)):
        if f.endswith((".cu", ".cuh", ".h")):
            for i, line in enumerate(open(os.path.join(csrc, f)), 1):
                if "getenv(" in line and not line.lstrip().startswith("//"):
                    hits.append((f, i))
    assert [f for f, _ in hits] == ["pd_abi.cu"], hits


def test_python_constants_match_the_header_enums():
    """The flag / status / dtype values the ctypes binding uses are the header's (a renumbered enum would otherwise pass unnoticed)."""
    txt = open(HEADER).read()
    found = dict((k, int(v)) for k, v in re.findall(r"\b(PD_[A-Z0-9_]+)\s*=\s*(\d+)", txt))
    for name in ("PD_FLAG_EXACT_COORDS", "PD_FLAG_NO_MASK_SUMMARY", "PD_FLAG_ACCUMULATE", "PD_FLAG_WORKSPACE_READY", "PD_WARP_DISP",
                 "PD_WARP_HOMOGRAPHY", "PD_WARP_DEPTH", "PD_LOSS_L1", "PD_LOSS_MIXTURE", "PD_LOSS_SSIM_L1", "PD_MASK_NONE", "PD_MASK_F32",
                 "PD_MASK_U8", "PD_DTYPE_F32", "PD_DTYPE_BF16"):
        assert name in found, name
        assert getattr(L, name) == found[name], name
