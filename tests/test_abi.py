"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/uvg266_b200.h declares, and refuses to work without a GPU (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import pytest

import uvg266_b200 as U


def test_library_loads_and_exports_all_declared_symbols():
    L = U.lib()
    names = U.exported_symbols()
    assert len(names) > 90
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_every_reference_strategy_type_has_a_dropin():
    """Type strings of STRATEGIES_PICTURE/DCT/QUANT_EXPORTS on the hot path
    (strategies-picture.h:210-243, strategies-dct.h:93-106, strategies-quant.h:104-110)."""
    L = U.lib()
    picture = (["reg_sad", "satd_any_size", "satd_any_size_vtm", "satd_any_size_quad", "pixels_calc_ssd",
                "get_optimized_sad", "ver_sad", "hor_sad", "generate_residual"] +
               ["%s_%dx%d" % (m, n, n) for m in ("sad", "satd") for n in (4, 8, 16, 32, 64)] +
               ["%s_%dx%d_dual" % (m, n, n) for m in ("sad", "satd") for n in (4, 8, 16, 32, 64)])
    for suf in ("8", "16"):
        for t in picture:
            assert hasattr(L, "uvg_b200_%s_%s" % (t, suf)), t
    dct = (["dct_%dx%d" % (n, n) for n in (4, 8, 16, 32)] + ["idct_%dx%d" % (n, n) for n in (4, 8, 16, 32)] +
           ["fast_forward_dst_4x4", "fast_inverse_dst_4x4", "mts_dct", "mts_idct"])
    quant = ["quant", "dequant", "coeff_abs_sum", "fast_coeff_cost"]
    for t in dct + quant:
        assert hasattr(L, "uvg_b200_" + t), t
    assert hasattr(L, "uvg_b200_quantize_residual_8") and hasattr(L, "uvg_b200_quantize_residual_16")
    assert hasattr(L, "uvg_b200_quant_cbcr_residual_8") and hasattr(L, "uvg_b200_quant_cbcr_residual_16")   # 6 of 6 quant types


def test_only_uvg_prefixed_symbols_are_exported():
    """tests/test_external_symbols.sh of the reference: exported symbols must be uvg_-prefixed."""
    out = subprocess.check_output(["nm", "-D", "--defined-only", U.LIB_PATH]).decode()
    bad = []
    for line in out.splitlines():
        parts = line.split()
        if len(parts) == 3 and parts[1] in "TDBR":
            if not parts[2].startswith("uvg_"):
                bad.append(parts[2])
    assert not bad, bad[:10]


def test_no_silent_cpu_fallback():
    """Without a CUDA device every entry point must fail loudly."""
    L = U.lib()
    if L.uvg_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    assert L.uvg_b200_init(0) != 0
    assert b"no CPU fallback" in L.uvg_b200_last_error()
    a = U.Blocks.dense(None, 8, 8)
    assert L.uvg_b200_distortion_dev(U.SAD, 8, 8, 8, C.byref(a), C.byref(a), None, 1, 1, None) != 0


def test_product_does_not_link_the_oracle():
    out = subprocess.check_output(["ldd", U.LIB_PATH]).decode()
    assert "oracle" not in out and "uvg266ref" not in out
    syms = subprocess.check_output(["nm", "-D", U.LIB_PATH]).decode()
    assert "orc_" not in syms and "refh_" not in syms


def test_ctypes_mirror_has_the_headers_struct_layout(tmp_path):
    """tests and bench.py reach the C ABI through the ctypes mirror in uvg266_b200/__init__.py: every public struct must
    have the size and field offsets a C compiler gives the header's."""
    import ctypes as C
    import shutil
    import subprocess
    import uvg266_b200 as U
    if not shutil.which("gcc"):
        pytest.skip("no C compiler")
    pairs = [("uvg_b200_blocks", U.Blocks), ("uvg_b200_quant_cfg", U.QuantCfg), ("uvg_b200_frame_level", U.FrameLevel),
             ("uvg_b200_frame_tu", U.FrameTu), ("uvg_b200_frame_cfg", U.FrameCfg), ("uvg_b200_frame_result", U.FrameResult),
             ("uvg_b200_frame_device", U.FrameDevice)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "uvg266_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {}
    for ln in out.splitlines():
        a, b, c = ln.split()
        got[(a, b)] = int(c)
    for cname, cls in pairs:
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)


def test_expand_coeffs_accepts_tus_in_any_order():
    """uvg_b200_frame_expand_coeffs is host code (no GPU needed): the frame pipeline guarantees that a TU's levels are
    contiguous and in index order from tu_level_off[t], NOT that the TUs follow each other in index order - the helper
    must rebuild the dense array from a shuffled layout."""
    import numpy as np
    L = U.lib()
    rng = np.random.default_rng(9)
    n_tus, cpt = 37, 64
    dense = np.where(rng.random(n_tus * cpt) < 0.2, rng.integers(-300, 301, n_tus * cpt), 0).astype(np.int16)
    dense[5 * cpt: 6 * cpt] = 0                                    # a TU without coefficients
    sig = np.zeros((n_tus * cpt + 31) // 32, np.uint32)
    for g in np.nonzero(dense)[0]:
        sig[g >> 5] |= np.uint32(1) << np.uint32(g & 31)
    order = rng.permutation(n_tus)
    off = np.zeros(n_tus, np.uint32)
    levels = []
    pos = 3                                                         # and not starting at zero either
    levels.extend([7, 7, 7])
    for t in order:
        nz = dense[t * cpt: (t + 1) * cpt]
        nz = nz[nz != 0]
        off[t] = pos
        levels.extend(int(v) for v in nz)
        pos += len(nz)
    levels = np.array(levels, np.int16)
    out = np.full(n_tus * cpt, 99, np.int16)
    L.uvg_b200_frame_expand_coeffs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
    L.uvg_b200_frame_expand_coeffs.restype = None
    L.uvg_b200_frame_expand_coeffs(sig.ctypes.data, off.ctypes.data, levels.ctypes.data, cpt, n_tus, out.ctypes.data)
    assert (out == dense).all()
