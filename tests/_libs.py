"""ctypes loaders for the test-side libraries (oracle restatement, compiled reference).

Nothing here is product code: the product is uvg266_b200 (libuvg266_b200.so).
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_i16p = C.POINTER(C.c_int16)
c_u32p = C.POINTER(C.c_uint32)


def pix_dtype(bd):
    return np.uint8 if bd == 8 else np.uint16


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _ensure_oracle_built():
    so = os.path.join(ORACLE_DIR, "build", "liboracle8.so")
    src = os.path.join(ORACLE_DIR, "uvg_oracle.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


class OrcQuantParams(C.Structure):
    _fields_ = [("qp", C.c_int32), ("bitdepth", C.c_int32), ("slice_intra", C.c_int32),
                ("signhide", C.c_int32), ("mts", C.c_int32), ("chroma_qp_map", C.c_void_p),
                ("isp_mode", C.c_int32)]


_oracles = {}


def load_oracle(bd=8):
    """The CPU restatement (oracle/uvg_oracle.c), built for `bd`-bit pixels."""
    if bd in _oracles:
        return _oracles[bd]
    _ensure_oracle_built()
    L = C.CDLL(os.path.join(ORACLE_DIR, "build", "liboracle%d.so" % bd))
    vp, i, u = C.c_void_p, C.c_int, C.c_uint
    L.orc_reg_sad.argtypes = [vp, vp, i, i, u, u]; L.orc_reg_sad.restype = u
    L.orc_sad_nxn.argtypes = [i, vp, vp]; L.orc_sad_nxn.restype = u
    L.orc_sad_nxn_dual.argtypes = [i, vp, vp, vp]
    L.orc_satd_4x4.argtypes = [vp, vp]; L.orc_satd_4x4.restype = u
    L.orc_satd_nxn.argtypes = [i, vp, vp]; L.orc_satd_nxn.restype = u
    L.orc_satd_nxn_dual.argtypes = [i, vp, vp, vp]
    L.orc_satd_any_size.argtypes = [i, i, vp, i, vp, i]; L.orc_satd_any_size.restype = u
    L.orc_satd_any_size_quad.argtypes = [i, i, vp, i, vp, i, vp]
    L.orc_satd_any_size_vtm.argtypes = [i, i, vp, i, vp, i]; L.orc_satd_any_size_vtm.restype = u
    L.orc_pixels_calc_ssd.argtypes = [vp, vp, i, i, i, i]; L.orc_pixels_calc_ssd.restype = u
    L.orc_ver_sad.argtypes = [vp, vp, i, i, C.c_uint32]; L.orc_ver_sad.restype = C.c_uint32
    L.orc_hor_sad.argtypes = [vp, vp, i, i] + [C.c_uint32] * 4; L.orc_hor_sad.restype = C.c_uint32
    L.orc_generate_residual.argtypes = [vp, vp, vp, i, i, i, i]
    L.orc_image_calc_sad.argtypes = [vp, vp] + [i] * 9; L.orc_image_calc_sad.restype = u
    L.orc_tr_matrix.argtypes = [i, i, vp]; L.orc_tr_matrix.restype = i
    L.orc_dct_nxn.argtypes = [i, i, vp, vp]
    L.orc_idct_nxn.argtypes = [i, i, vp, vp]
    L.orc_fast_forward_dst_4x4.argtypes = [i, vp, vp]
    L.orc_fast_inverse_dst_4x4.argtypes = [i, vp, vp]
    L.orc_get_tr_type.argtypes = [i] * 9 + [vp, vp]
    L.orc_mts_dct.argtypes = [i] * 9 + [vp, vp, i]
    L.orc_mts_idct.argtypes = [i] * 9 + [vp, vp, i]
    L.orc_scan_table.argtypes = [i, i, vp]
    L.orc_quant.argtypes = [vp, vp, vp, i, i, i, i, i, i]
    L.orc_dequant.argtypes = [vp, vp, vp, i, i, i, i, i]
    L.orc_quantize_residual.argtypes = [vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, i]
    L.orc_quantize_residual.restype = i
    L.orc_quant_cbcr_residual.argtypes = [vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, i]
    L.orc_quant_cbcr_residual.restype = i
    L.orc_coeff_abs_sum.argtypes = [vp, C.c_size_t]; L.orc_coeff_abs_sum.restype = C.c_uint32
    L.orc_fast_coeff_cost.argtypes = [vp, i, i, C.c_uint64]; L.orc_fast_coeff_cost.restype = C.c_uint32
    _oracles[bd] = L
    return L


_refs = {}


def ref_available(bd=8):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "bd%d" % bd, "libref_harness.so"))


class Ref:
    """The unmodified reference (oracle/_ref/bd*/libuvg266ref.so) behind ref_harness.c."""

    def __init__(self, bd, opts):
        d = os.path.join(ORACLE_DIR, "_ref", "bd%d" % bd)
        self.bd = bd
        self.H = H = C.CDLL(os.path.join(d, "libref_harness.so"))
        H.refh_get.restype = C.c_void_p
        H.refh_get.argtypes = [C.c_char_p, C.c_char_p]
        H.refh_open.argtypes = [C.c_char_p]
        H.refh_scan_table.restype = c_u32p
        H.refh_chroma_qp_map.restype = C.c_void_p
        H.refh_image_calc_sad.restype = C.c_uint
        H.refh_image_calc_sad.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 8
        H.refh_mts.argtypes = [C.c_char_p] + [C.c_int] * 10 + [C.c_void_p, C.c_void_p, C.c_int]
        H.refh_get_tr_type.argtypes = [C.c_int] * 9 + [C.c_void_p, C.c_void_p]
        H.refh_quant.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p] + [C.c_int] * 7
        H.refh_dequant.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5
        H.refh_quantize_residual.argtypes = [C.c_char_p] + [C.c_int] * 9 + [C.c_void_p] * 4 + [C.c_int]
        H.refh_quant_cbcr_residual.argtypes = [C.c_char_p] + [C.c_int] * 8 + [C.c_void_p] * 7 + [C.c_int]
        H.refh_quant_cbcr_residual.restype = C.c_int
        H.refh_ctu_pass.restype = C.c_double
        H.refh_pair_batch.restype = C.c_double
        self.opts = None
        self.open(opts)

    def open(self, opts):
        if opts != self.opts:
            # the reference prints its strategy report on stderr at every open; keep the
            # test log readable by silencing fd 2 for the call.
            saved = os.dup(2)
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 2)
            try:
                ok = self.H.refh_open(opts.encode())
            finally:
                os.dup2(saved, 2)
                os.close(devnull)
                os.close(saved)
            assert ok, "reference encoder_open failed for %r" % opts
            self.opts = opts

    def fn(self, type_, restype, argtypes, strategy="generic"):
        p = self.H.refh_get(type_.encode(), strategy.encode())
        if not p:
            return None
        return C.CFUNCTYPE(restype, *argtypes)(p)


def load_ref(bd=8, opts="preset=ultrafast,qp=27"):
    if not ref_available(bd):
        return None
    if bd not in _refs:
        _refs[bd] = Ref(bd, opts)
    else:
        _refs[bd].open(opts)
    return _refs[bd]
