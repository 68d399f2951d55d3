"""uvg266_b200 - B200-native distortion / transform / quant kernels for uvg266.

Host-side mirror (ctypes) of the C ABI in include/uvg266_b200.h. The functions carry the
names of the reference's strategy table (reg_sad, satd_any_size, mts_dct, quant, ...;
strategies-picture.h / strategies-dct.h / strategies-quant.h) so that the parity tests
read like the reference's own tests.

The product is libuvg266_b200.so (hand-written CUDA, sm_100a). This module only loads it:
there is no Python or CPU implementation of any operation here, and importing it on a
machine without the built library raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libuvg266_b200.so")

# enum uvg_b200_metric
SAD, SAD_NXN, SATD, SATD_4X4, SATD_VTM, SSD, SATD_QUAD = range(7)
DCT2, DCT8, DST7 = 0, 1, 2
CU_INTRA, CU_INTER = 1, 2
COLOR_Y, COLOR_U, COLOR_V = 0, 1, 2
MTS_OFF, MTS_INTRA, MTS_INTER, MTS_BOTH, MTS_IMPLICIT = range(5)


class Blocks(C.Structure):
    """uvg_b200_blocks: affine addressing of a batch of blocks."""
    _fields_ = [("base", C.c_void_p), ("stride", C.c_uint32), ("blocks_per_row", C.c_uint32),
                ("col_pitch", C.c_uint64), ("row_pitch", C.c_uint64)]

    @staticmethod
    def dense(base, w, h):
        return Blocks(base, w, 1, 0, w * h)

    @staticmethod
    def plane(base, stride, plane_w, w, h):
        """Tiling of a picture plane by w*h blocks in raster order."""
        return Blocks(base, stride, plane_w // w, w, h * stride)


class QuantCfg(C.Structure):
    """uvg_b200_quant_cfg: what quant/dequant read through encoder_state_t."""
    _fields_ = [("qp", C.c_int32), ("bitdepth", C.c_int32), ("slice_intra", C.c_int32),
                ("signhide", C.c_int32), ("mts", C.c_int32), ("isp_mode", C.c_int32),
                ("chroma_qp_map", C.c_int8 * 128), ("has_chroma_qp_map", C.c_int32),
                ("cu_lfnst_idx", C.c_int32), ("cu_cr_lfnst_idx", C.c_int32), ("lfnst_index", C.c_int32),
                ("lfnst_apply", C.c_int32), ("lfnst_mode", C.c_int32), ("lfnst_wa_log2_w", C.c_int32),
                ("lfnst_wa_log2_h", C.c_int32)]

    @staticmethod
    def make(qp, bitdepth=8, slice_intra=1, signhide=0, mts=0, chroma_map=None, isp_mode=0):
        c = QuantCfg(qp, bitdepth, slice_intra, signhide, mts, isp_mode)
        if chroma_map is not None:
            for i, v in enumerate(chroma_map[:128]):
                c.chroma_qp_map[i] = int(v)
            c.has_chroma_qp_map = 1
        return c


FRAME_MAX_LEVELS = 4
FRAME_NO_H2D, FRAME_NO_D2H, FRAME_DENSE_COEFF, FRAME_ALL_COSTS = 1, 2, 4, 8


class FrameLevel(C.Structure):
    _fields_ = [("n", C.c_int32), ("k", C.c_int32), ("modes", C.c_int8 * 68)]


class FrameTu(C.Structure):
    _fields_ = [("n", C.c_int32), ("per_ctu", C.c_int32)]


class FrameCfg(C.Structure):
    """uvg_b200_frame_cfg"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bitdepth", C.c_int32), ("n_levels", C.c_int32),
                ("level", FrameLevel * FRAME_MAX_LEVELS), ("n_tus", C.c_int32), ("tu", FrameTu * FRAME_MAX_LEVELS),
                ("n_best", C.c_int32), ("quant", QuantCfg), ("cost_weights", C.c_uint64)]

    @staticmethod
    def make(width, height, levels, tus, n_best, quant, cost_weights):
        """levels: [(n, [modes])], tus: [(n, per_ctu)]"""
        c = FrameCfg()
        c.width, c.height, c.bitdepth = width, height, 8
        c.n_levels = len(levels)
        for i, (n, modes) in enumerate(levels):
            c.level[i].n, c.level[i].k = n, len(modes)
            for j, m in enumerate(modes):
                c.level[i].modes[j] = m
        c.n_tus = len(tus)
        for i, (n, cnt) in enumerate(tus):
            c.tu[i].n, c.tu[i].per_ctu = n, cnt
        c.n_best = n_best
        c.quant = quant
        c.cost_weights = cost_weights
        return c


class FrameResult(C.Structure):
    """uvg_b200_frame_result: host pointers owned by the frame object"""
    _A = C.c_void_p * FRAME_MAX_LEVELS
    _fields_ = [("ctu_begin", C.c_int32), ("ctu_end", C.c_int32), ("rough_cost", _A), ("rough_best", _A),
                ("rough_best_cost", _A), ("tu_mode", _A), ("tu_has_coeffs", _A), ("tu_ssd", _A), ("tu_cost", _A), ("tu_sig", _A),
                ("tu_level_off", _A), ("tu_coeff_dense", _A), ("levels", C.c_void_p), ("n_levels", C.c_uint64),
                ("device_ms", C.c_float)]


class FrameDevice(C.Structure):
    _A = C.c_void_p * FRAME_MAX_LEVELS
    _fields_ = [("plane", C.c_void_p), ("rough_cost", _A), ("rough_best", _A), ("tu_pred", _A), ("tu_rec", _A),
                ("tu_coeff", _A), ("stream", C.c_void_p)]


def np_view(ptr, dtype, count):
    """numpy view of `count` items at a host pointer handed out by the library (no copy)."""
    if not ptr or count == 0:
        return np.zeros(0, dtype)
    n = count * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_uint8 * n).from_address(ptr), dtype=dtype, count=count)


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    """Load libuvg266_b200.so; raise loudly when it is missing (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "%s not found - build it with `make -C uvg266_b200` (or __graft_entry__.build()); "
            "uvg266_b200 has no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    BP = C.POINTER(Blocks)
    QP = C.POINTER(QuantCfg)
    L.uvg_b200_init.argtypes = [i]
    L.uvg_b200_last_error.restype = C.c_char_p
    L.uvg_b200_version.restype = C.c_char_p
    L.uvg_b200_launch_count.restype = C.c_uint64
    L.uvg_b200_rough_tma_launches.restype = C.c_uint64
    L.uvg_b200_host_alloc.restype = vp
    L.uvg_b200_host_alloc.argtypes = [sz]
    L.uvg_b200_host_free.argtypes = [vp]
    L.uvg_b200_distortion_dev.argtypes = [i, i, i, i, BP, BP, vp, sz, i, vp]
    L.uvg_b200_rough_cost_dev.argtypes = [i, i, BP, BP, vp, sz, i, vp]
    L.uvg_b200_sad_search_dev.argtypes = [i, i, i, vp, vp, i, i, i, vp, vp, vp, vp, i, vp, sz, vp]
    L.uvg_b200_residual_dev.argtypes = [i, i, i, BP, BP, vp, sz, vp]
    L.uvg_b200_transform_dev.argtypes = [i] * 8 + [vp, vp, sz, vp]
    L.uvg_b200_quant_dev.argtypes = [QP, vp, vp, i, i, i, i, i, sz, vp]
    L.uvg_b200_dequant_dev.argtypes = [QP, vp, vp, i, i, i, i, sz, vp]
    L.uvg_b200_quantize_residual_dev.argtypes = [QP, i, i, i, i, i, i, i, BP, BP, BP, vp, vp, vp, vp,
                                                 C.c_uint64, sz, vp]
    L.uvg_b200_quant_cbcr_residual_dev.argtypes = [QP, i, i, i, i, i, i, i, BP, BP, BP, BP, BP, BP, vp, vp, sz, vp]
    L.uvg_b200_coeff_abs_sum_dev.argtypes = [vp, sz, vp, sz, vp]
    L.uvg_b200_fast_coeff_cost_dev.argtypes = [vp, i, i, C.c_uint64, vp, sz, vp]
    L.uvg_b200_distortion_host.argtypes = [i, i, i, i, BP, sz, BP, sz, vp, sz, i]
    L.uvg_b200_rough_cost_host.argtypes = [i, i, BP, sz, BP, sz, vp, sz, i]
    L.uvg_b200_transform_host.argtypes = [i] * 8 + [vp, vp, sz]
    L.uvg_b200_quantize_residual_host.argtypes = [QP, i, i, i, i, i, i, i, BP, sz, BP, sz, BP, sz, vp, vp,
                                                  vp, vp, C.c_uint64, sz]
    L.uvg_b200_crc32c_dev.argtypes = [i, i, BP, vp, sz, vp]
    L.uvg_b200_bipred_average_dev.argtypes = [i, i, i, BP, vp, i, vp, i, sz, vp]
    L.uvg_b200_ipol_dev.argtypes = [i, i, i, vp, i, i, i, i, i, vp, vp, vp, vp, vp, sz, vp]
    L.uvg_b200_frac_search_dev.argtypes = [i, i, i, vp, vp, i, i, i, vp, vp, vp, vp, vp, i, vp, sz, vp]
    L.uvg_b200_intra_predict_dev.argtypes = [i, i, vp, vp, i, vp, sz, vp]
    L.uvg_b200_intra_rough_dev.argtypes = [i, i, BP, vp, vp, i, vp, sz, vp]
    L.uvg_b200_intra_predict_host.argtypes = [i, i, vp, vp, i, vp, sz]
    L.uvg_b200_intra_rough_host.argtypes = [i, i, BP, sz, vp, vp, i, vp, sz]
    FP = C.POINTER(FrameCfg)
    L.uvg_b200_frame_create.argtypes = [FP]; L.uvg_b200_frame_create.restype = vp
    L.uvg_b200_frame_destroy.argtypes = [vp]; L.uvg_b200_frame_destroy.restype = None
    L.uvg_b200_frame_submit.argtypes = [vp, vp, sz, i, i, i]
    L.uvg_b200_frame_wait.argtypes = [vp, C.POINTER(FrameResult)]
    L.uvg_b200_frame_device_buffers.argtypes = [vp, C.POINTER(FrameDevice)]
    L.uvg_b200_device_read.argtypes = [vp, vp, sz]
    L.uvg_b200_pixel_var_dev.argtypes = [i, vp, sz, C.c_uint32, vp, sz, vp]
    for suf in ("8", "16"):
        f = getattr(L, "uvg_b200_pixel_var_" + suf); f.argtypes = [vp, C.c_uint32]; f.restype = C.c_double
    L.uvg_b200_set_lfnst_tables.argtypes = [vp, vp]
    L.uvg_b200_band_rows.argtypes = [i, i, i, vp, vp]; L.uvg_b200_band_rows.restype = None
    L.uvg_b200_comm_unique_id.argtypes = [vp]
    L.uvg_b200_comm_create.argtypes = [vp, i, i]; L.uvg_b200_comm_create.restype = vp
    L.uvg_b200_comm_destroy.argtypes = [vp]; L.uvg_b200_comm_destroy.restype = None
    L.uvg_b200_comm_nccl_version.restype = i
    L.uvg_b200_refpic_exchange.argtypes = [vp, vp, sz, i, vp]
    L.uvg_b200_frame_expand_coeffs.argtypes = [vp, vp, vp, i, sz, vp]; L.uvg_b200_frame_expand_coeffs.restype = None
    # drop-ins
    for suf in ("8", "16"):
        for n in (4, 8, 16, 32, 64):
            for m in ("sad", "satd"):
                f = getattr(L, "uvg_b200_%s_%dx%d_%s" % (m, n, n, suf))
                f.argtypes = [vp, vp]; f.restype = C.c_uint
                g = getattr(L, "uvg_b200_%s_%dx%d_dual_%s" % (m, n, n, suf))
                g.argtypes = [vp, vp, C.c_uint, vp]; g.restype = None
        f = getattr(L, "uvg_b200_reg_sad_" + suf); f.argtypes = [vp, vp, i, i, C.c_uint, C.c_uint]; f.restype = C.c_uint
        for nm in ("satd_any_size_", "satd_any_size_vtm_"):
            f = getattr(L, "uvg_b200_" + nm + suf); f.argtypes = [i, i, vp, i, vp, i]; f.restype = C.c_uint
        f = getattr(L, "uvg_b200_satd_any_size_quad_" + suf)
        f.argtypes = [i, i, vp, i, vp, i, C.c_uint, vp, vp]; f.restype = None
        f = getattr(L, "uvg_b200_pixels_calc_ssd_" + suf); f.argtypes = [vp, vp, i, i, i, i]; f.restype = C.c_uint
        f = getattr(L, "uvg_b200_ver_sad_" + suf); f.argtypes = [vp, vp, i, i, C.c_uint32]; f.restype = C.c_uint32
        f = getattr(L, "uvg_b200_hor_sad_" + suf); f.argtypes = [vp, vp, i, i] + [C.c_uint32] * 4; f.restype = C.c_uint32
        f = getattr(L, "uvg_b200_generate_residual_" + suf); f.argtypes = [vp, vp, vp, i, i, i, i]; f.restype = None
        f = getattr(L, "uvg_b200_get_optimized_sad_" + suf); f.argtypes = [C.c_int32]; f.restype = vp
        f = getattr(L, "uvg_b200_fme_filter_" + suf); f.argtypes = [i, vp, i, i, i, vp, i, i, i, i]; f.restype = None
        for nm in ("crc32c_4x4_", "crc32c_8x8_"):
            f = getattr(L, "uvg_b200_" + nm + suf); f.argtypes = [vp, C.c_uint32]; f.restype = C.c_uint32
        f = getattr(L, "uvg_b200_bipred_average_" + suf); f.argtypes = [vp, i, vp, i, vp, i, i, i]; f.restype = None
        for nm in ("sample_quarterpel_luma_", "sample_quarterpel_luma_hi_", "sample_octpel_chroma_", "sample_octpel_chroma_hi_"):
            f = getattr(L, "uvg_b200_" + nm + suf); f.argtypes = [vp, i, i, i, vp, i, i, i]; f.restype = None
        f = getattr(L, "uvg_b200_angular_pred_" + suf); f.argtypes = [i, i, i, i, vp, vp, vp, i, i]; f.restype = None
        f = getattr(L, "uvg_b200_intra_pred_planar_" + suf); f.argtypes = [i, i, vp, vp, vp]; f.restype = None
        f = getattr(L, "uvg_b200_pdpc_planar_dc_" + suf); f.argtypes = [i, i, i, vp, vp, vp]; f.restype = None
        f = getattr(L, "uvg_b200_quantize_residual_" + suf)
        f.argtypes = [QP, i, i, i, i, i, i, i, i, vp, vp, vp, vp, i]; f.restype = i
        f = getattr(L, "uvg_b200_quant_cbcr_residual_" + suf)
        f.argtypes = [QP, i, i, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, i]; f.restype = i
    for n in (4, 8, 16, 32):
        for m in ("dct", "idct"):
            f = getattr(L, "uvg_b200_%s_%dx%d" % (m, n, n)); f.argtypes = [C.c_int8, vp, vp]; f.restype = None
    for nm in ("fast_forward_dst_4x4", "fast_inverse_dst_4x4"):
        f = getattr(L, "uvg_b200_" + nm); f.argtypes = [C.c_int8, vp, vp]; f.restype = None
    for nm in ("mts_dct", "mts_idct"):
        f = getattr(L, "uvg_b200_" + nm)
        f.argtypes = [C.c_int8, i, i, i, i, i, i, C.c_int8, C.c_int8, vp, vp, C.c_int8]; f.restype = None
    L.uvg_b200_get_tr_type.argtypes = [i] * 9 + [vp, vp]
    L.uvg_b200_quant.argtypes = [QP, vp, vp, C.c_int32, C.c_int32, i, C.c_int8, C.c_int8, C.c_int8, C.c_uint8]
    L.uvg_b200_quant.restype = None
    L.uvg_b200_dequant.argtypes = [QP, vp, vp, C.c_int32, C.c_int32, i, C.c_int8, C.c_int8]
    L.uvg_b200_dequant.restype = None
    L.uvg_b200_coeff_abs_sum.argtypes = [vp, sz]; L.uvg_b200_coeff_abs_sum.restype = C.c_uint32
    L.uvg_b200_fast_coeff_cost.argtypes = [vp, C.c_int32, C.c_int32, C.c_uint64]
    L.uvg_b200_fast_coeff_cost.restype = C.c_uint32
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError("uvg266_b200: " + lib().uvg_b200_last_error().decode())


def init(device=0):
    check(lib().uvg_b200_init(device))


def suffix(bitdepth):
    return "8" if bitdepth == 8 else "16"


def dropin(name, bitdepth=8):
    """The drop-in leaf function for a reference strategy type string, e.g. 'satd_8x8'."""
    L = lib()
    for cand in ("uvg_b200_%s_%s" % (name, suffix(bitdepth)), "uvg_b200_%s" % name):
        if hasattr(L, cand):
            return getattr(L, cand)
    raise AttributeError(name)


def exported_symbols():
    """Every function include/uvg266_b200.h declares (used by the CPU-side ABI test)."""
    hdr = open(os.path.join(os.path.dirname(_HERE), "include", "uvg266_b200.h")).read()
    import re
    names = set(re.findall(r"\b(uvg_b200_[a-z0-9_]+)\s*\(", hdr))
    out = set()
    for n in names:
        if n.endswith("_##SUF") or "##" in n:
            continue
        out.add(n)
    # macro-declared picture set
    for m in re.findall(r"uvg_b200_([a-z0-9_x]+)_##SUF\(", hdr):
        for suf in ("8", "16"):
            out.add("uvg_b200_%s_%s" % (m, suf))
    return sorted(out)
