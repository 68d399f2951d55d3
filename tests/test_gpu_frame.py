"""GPU parity of the device-resident frame pipeline (uvg_b200_frame_*, include/uvg266_b200.h section 4):
every output of a pass - rough costs, candidate lists, the mode each TU was predicted with, has_coeffs, SSD,
coefficient cost, the quantised coefficients (compact and dense) and the reconstruction - against the UNMODIFIED
reference doing the same per-CTU work with its own functions (oracle/ref_harness.c::refh_frame_pass:
uvg_intra_predict, satd/sad duals, quantize_residual, pixels_calc_ssd, fast_coeff_cost), and against the oracle
restatement on sampled blocks. Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import uvg266_b200 as U
from _libs import load_oracle, load_ref, ptr, OrcQuantParams

pytestmark = pytest.mark.gpu

MODES8 = list(range(0, 67, 3))                 # 23 candidates
MODES16 = list(range(0, 67, 3)) + [1]          # 24
WEIGHTS = (28 << 0) | (300 << 16) | (700 << 32) | (1100 << 48)
LEVELS = [(8, MODES8), (16, MODES16)]
TUS = [(4, 240), (8, 120), (16, 24), (32, 5)]


class RefLevel(C.Structure):
    _fields_ = [("n", C.c_int), ("count", C.c_int), ("modes", C.c_int)]


class RefFrameLevel(C.Structure):
    _fields_ = [("n", C.c_int), ("k", C.c_int), ("modes", C.c_int8 * 68)]


def synth(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    base = 128 + 60 * np.sin(x * 0.013 + 1) * np.cos(y * 0.011) + 40 * np.sin((x + y) * 0.047)
    return np.clip(base + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)


def run_gpu(plane, levels, tus, n_best, qp, rows=None, flags=U.FRAME_DENSE_COEFF | U.FRAME_ALL_COSTS):
    L = U.lib()
    h, w = plane.shape
    cfg = U.FrameCfg.make(w, h, levels, tus, n_best, U.QuantCfg.make(qp, 8, 1, 0, 0), WEIGHTS)
    f = L.uvg_b200_frame_create(C.byref(cfg))
    assert f, L.uvg_b200_last_error()
    try:
        r0, r1 = rows if rows else (0, 0)
        U.check(L.uvg_b200_frame_submit(f, plane.ctypes.data, w, r0, r1, flags))
        res = U.FrameResult()
        U.check(L.uvg_b200_frame_wait(f, C.byref(res)))
        n_ctu = res.ctu_end - res.ctu_begin
        out = {"ctu_begin": res.ctu_begin, "n_ctu": n_ctu, "n_levels": int(res.n_levels), "device_ms": res.device_ms}
        for i, (n, modes) in enumerate(levels):
            nb = n_ctu * (64 // n) ** 2
            out["cost%d" % i] = U.np_view(res.rough_cost[i], np.uint32, nb * len(modes)).copy()
            out["best%d" % i] = U.np_view(res.rough_best[i], np.uint8, nb * n_best).copy()
            out["bestcost%d" % i] = U.np_view(res.rough_best_cost[i], np.uint32, nb * n_best).copy()
        levels_arr = U.np_view(res.levels, np.int16, int(res.n_levels)).copy()
        out["levels"] = levels_arr
        for i, (n, cnt) in enumerate(tus):
            nt = n_ctu * cnt
            out["mode%d" % i] = U.np_view(res.tu_mode[i], np.uint8, nt).copy()
            out["has%d" % i] = U.np_view(res.tu_has_coeffs[i], np.uint8, nt).copy()
            out["ssd%d" % i] = U.np_view(res.tu_ssd[i], np.uint32, nt).copy()
            out["tcost%d" % i] = U.np_view(res.tu_cost[i], np.uint32, nt).copy()
            sig = U.np_view(res.tu_sig[i], np.uint32, (nt * n * n + 31) // 32).copy()
            off = U.np_view(res.tu_level_off[i], np.uint32, nt + 1).copy()
            dense = np.zeros(nt * n * n, np.int16)
            L.uvg_b200_frame_expand_coeffs(sig.ctypes.data, off.ctypes.data, levels_arr.ctypes.data, n * n, nt, dense.ctypes.data)
            out["coeff%d" % i] = dense
            out["sig%d" % i] = sig
            if flags & U.FRAME_DENSE_COEFF:
                out["dense%d" % i] = U.np_view(res.tu_coeff_dense[i], np.int16, nt * n * n).copy()
        # the reconstruction stays on the device: read it through the device view
        dv = U.FrameDevice()
        U.check(L.uvg_b200_frame_device_buffers(f, C.byref(dv)))
        for i, (n, cnt) in enumerate(tus):
            rec = np.zeros(n_ctu * cnt * n * n, np.uint8)
            U.check(L.uvg_b200_device_read(rec.ctypes.data, dv.tu_rec[i], rec.nbytes))
            out["rec%d" % i] = rec
        return out
    finally:
        L.uvg_b200_frame_destroy(f)


def run_ref(plane, levels, tus, n_best, qp, rows=None, threads=8):
    R = load_ref(8, "preset=ultrafast,qp=%d,signhide=0,rdoq=0" % qp)
    if R is None:
        pytest.skip("oracle/_ref missing")
    R.H.refh_set_state(qp, 1)
    h, w = plane.shape
    cols, nrows = (w + 63) // 64, (h + 63) // 64
    r0, r1 = rows if rows else (0, nrows)
    if r1 <= 0:
        r1 = nrows
    c0, c1 = r0 * cols, r1 * cols
    n_ctu = c1 - c0
    lv = (RefFrameLevel * len(levels))()
    for i, (n, modes) in enumerate(levels):
        lv[i].n, lv[i].k = n, len(modes)
        for j, m in enumerate(modes):
            lv[i].modes[j] = m
    tu = (RefLevel * len(tus))(*[RefLevel(n, c, 0) for (n, c) in tus])
    out = {"n_ctu": n_ctu}
    keep = []

    def parr(arrs):
        a = (C.c_void_p * len(arrs))(*[x.ctypes.data for x in arrs])
        keep.append(a)
        return a
    cost = [np.zeros(n_ctu * (64 // n) ** 2 * len(m), np.uint32) for (n, m) in levels]
    best = [np.zeros(n_ctu * (64 // n) ** 2 * n_best, np.uint8) for (n, m) in levels]
    mode = [np.zeros(n_ctu * c, np.uint8) for (n, c) in tus]
    has = [np.zeros(n_ctu * c, np.uint8) for (n, c) in tus]
    ssd = [np.zeros(n_ctu * c, np.uint32) for (n, c) in tus]
    tcost = [np.zeros(n_ctu * c, np.uint32) for (n, c) in tus]
    coeff = [np.zeros(n_ctu * c * n * n, np.int16) for (n, c) in tus]
    rec = [np.zeros(n_ctu * c * n * n, np.uint8) for (n, c) in tus]
    R.H.refh_frame_pass.restype = C.c_double
    R.H.refh_frame_pass.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_uint64]
    secs = R.H.refh_frame_pass(threads, plane.ctypes.data, w, h, w, c0, c1, len(levels), lv, parr(cost), parr(best), len(tus), tu,
                               n_best, parr(mode), parr(has), parr(ssd), parr(tcost), parr(coeff), parr(rec), WEIGHTS)
    out["secs"] = secs
    for i in range(len(levels)):
        out["cost%d" % i], out["best%d" % i] = cost[i], best[i]
    for i in range(len(tus)):
        out["mode%d" % i], out["has%d" % i], out["ssd%d" % i], out["tcost%d" % i] = mode[i], has[i], ssd[i], tcost[i]
        out["coeff%d" % i], out["rec%d" % i] = coeff[i], rec[i]
    return out


def compare(g, r, levels, tus):
    for i in range(len(levels)):
        assert (g["cost%d" % i] == r["cost%d" % i]).all(), "rough costs, level %d" % i
        assert (g["best%d" % i] == r["best%d" % i]).all(), "candidate lists, level %d" % i
        # the costs that travel with the candidate lists = the full table at the listed modes
        k = len(levels[i][1])
        modes = np.array(levels[i][1]); full = r["cost%d" % i].reshape(-1, k); bm = r["best%d" % i].reshape(full.shape[0], -1)
        idx = np.array([[int(np.where(modes == m)[0][0]) for m in row] for row in bm])
        assert (g["bestcost%d" % i].reshape(bm.shape) == np.take_along_axis(full, idx, axis=1)).all(), "candidate costs, level %d" % i
    for i, (n, cnt) in enumerate(tus):
        for key in ("mode", "has", "ssd", "tcost", "coeff", "rec"):
            a, b = g["%s%d" % (key, i)], r["%s%d" % (key, i)]
            assert a.shape == b.shape and (a == b).all(), "%s of the %dx%d TUs" % (key, n, n)
        if "dense%d" % i in g:
            assert (g["dense%d" % i] == r["coeff%d" % i]).all()
    nnz = sum(int(np.count_nonzero(r["coeff%d" % i])) for i in range(len(tus)))
    assert g["n_levels"] == nnz


@pytest.fixture(scope="module", autouse=True)
def _init():
    U.init(0)


@pytest.mark.parametrize("w,h,qp,n_best", [(256, 128, 27, 2), (200, 136, 22, 3), (136, 72, 37, 1)])
def test_frame_pass_equals_reference(w, h, qp, n_best):
    """Whole pictures, incl. sizes that are not multiples of the CTU / of 8 (border replication)."""
    plane = synth(w, h, 11 + w)
    g = run_gpu(plane, LEVELS, TUS, n_best, qp)
    r = run_ref(plane, LEVELS, TUS, n_best, qp)
    assert g["n_ctu"] == r["n_ctu"] == ((w + 63) // 64) * ((h + 63) // 64)
    compare(g, r, LEVELS, TUS)
    assert any(g["has%d" % i].any() for i in range(len(TUS)))


def test_frame_band_equals_slice_of_picture():
    """CTU-row bands (the multi-GPU shard and the in-call pipelining unit) give exactly the picture's results."""
    plane = synth(192, 256, 5)
    r = run_ref(plane, LEVELS, TUS, 2, 27, rows=(1, 3))
    g = run_gpu(plane, LEVELS, TUS, 2, 27, rows=(1, 3))
    assert g["ctu_begin"] == 3 and g["n_ctu"] == 6
    compare(g, r, LEVELS, TUS)


def test_frame_other_level_sets():
    """One 32x32 level with all 67 modes feeding 32x32 and 16x16 TUs; flat picture (all-zero coefficient output)."""
    plane = synth(128, 128, 3)
    lv = [(32, list(range(67)))]
    tus = [(32, 4), (16, 20)]
    compare(run_gpu(plane, lv, tus, 2, 30), run_ref(plane, lv, tus, 2, 30), lv, tus)
    flat = np.full((64, 128), 77, np.uint8)
    g = run_gpu(flat, LEVELS, TUS, 2, 27)
    assert g["n_levels"] == 0 and not any(g["has%d" % i].any() for i in range(len(TUS)))
    compare(g, run_ref(flat, LEVELS, TUS, 2, 27), LEVELS, TUS)


def test_frames_in_flight_from_two_threads_repeat_their_results():
    """Two frame objects driven from two host threads (the bench's e2e driver, an encoder's frame-level parallelism): the
    library orders their kernels in one compute stream and replays per-stage CUDA graphs from the third pass on. Every
    pass of either thread must give what a pass of that picture gives on its own - checked on the expanded
    coefficients, the per-TU results and the candidate lists (the order of the TUs inside `levels` is unspecified)."""
    import threading
    L = U.lib()
    w, h = 320, 192
    planes = [synth(w, h, 21), synth(w, h, 22)]
    alone = [run_gpu(p, LEVELS, TUS, 2, 27, flags=0) for p in planes]
    cfg = U.FrameCfg.make(w, h, LEVELS, TUS, 2, U.QuantCfg.make(27, 8, 1, 0, 0), WEIGHTS)
    frames = [L.uvg_b200_frame_create(C.byref(cfg)) for _ in planes]
    assert all(frames)
    errors = []

    def drive(i):
        try:
            res = U.FrameResult()
            for it in range(6):
                U.check(L.uvg_b200_frame_submit(frames[i], planes[i].ctypes.data, w, 0, 0, 0))
                U.check(L.uvg_b200_frame_wait(frames[i], C.byref(res)))
                n_ctu = res.ctu_end - res.ctu_begin
                assert int(res.n_levels) == alone[i]["n_levels"], (i, it)
                lv = U.np_view(res.levels, np.int16, int(res.n_levels)).copy()
                for k, (n, modes) in enumerate(LEVELS):
                    nb = n_ctu * (64 // n) ** 2
                    assert (U.np_view(res.rough_best[k], np.uint8, nb * 2) == alone[i]["best%d" % k]).all(), (i, it, k)
                    assert (U.np_view(res.rough_best_cost[k], np.uint32, nb * 2) == alone[i]["bestcost%d" % k]).all(), (i, it, k)
                for k, (n, cnt) in enumerate(TUS):
                    nt = n_ctu * cnt
                    for key, field, dt in (("mode", res.tu_mode, np.uint8), ("has", res.tu_has_coeffs, np.uint8),
                                           ("ssd", res.tu_ssd, np.uint32), ("tcost", res.tu_cost, np.uint32)):
                        assert (U.np_view(field[k], dt, nt) == alone[i]["%s%d" % (key, k)]).all(), (i, it, key, k)
                    sig = U.np_view(res.tu_sig[k], np.uint32, (nt * n * n + 31) // 32).copy()
                    off = U.np_view(res.tu_level_off[k], np.uint32, nt).copy()
                    dense = np.zeros(nt * n * n, np.int16)
                    L.uvg_b200_frame_expand_coeffs(sig.ctypes.data, off.ctypes.data, lv.ctypes.data, n * n, nt, dense.ctypes.data)
                    assert (dense == alone[i]["coeff%d" % k]).all(), (i, it, k)
        except Exception as exc:                                  # surfaced in the main thread
            errors.append(exc)

    th = [threading.Thread(target=drive, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for f in frames:
        L.uvg_b200_frame_destroy(f)
    assert not errors, errors[0]


def test_frame_samples_against_oracle():
    """The same pass against the oracle restatement (no reference library needed) on sampled blocks."""
    O = load_oracle(8)
    vp, ci = C.c_void_p, C.c_int
    O.orc_intra_predict.argtypes = [ci, ci, ci, vp, vp, vp]
    w, h, qp = 192, 128, 27
    plane = synth(w, h, 99)
    g = run_gpu(plane, LEVELS, TUS, 2, qp)
    cols = w // 64

    def refs_of(X, Y, n):
        i = np.minimum(np.arange(3 * n + 3), 2 * n)
        top = plane[np.clip(Y - 1, 0, h - 1), np.clip(X - 1 + i, 0, w - 1)]
        left = plane[np.clip(Y - 1 + i, 0, h - 1), np.clip(X - 1, 0, w - 1)]
        return np.ascontiguousarray(top), np.ascontiguousarray(left)
    rng = np.random.default_rng(1)
    for li, (n, modes) in enumerate(LEVELS):
        ppc = (64 // n) ** 2
        for b in rng.integers(0, g["n_ctu"] * ppc, 12):
            ctu, j = divmod(int(b), ppc)
            X, Y = (ctu % cols) * 64 + (j % (64 // n)) * n, (ctu // cols) * 64 + (j // (64 // n)) * n
            top, left = refs_of(X, Y, n)
            orig = np.ascontiguousarray(plane[Y:Y + n, X:X + n]).ravel()
            for m, mode in enumerate(modes):
                pred = np.zeros(n * n, np.uint8)
                O.orc_intra_predict(n, n, mode, ptr(top), ptr(left), ptr(pred))
                e = min(O.orc_satd_nxn(n, ptr(orig), ptr(pred)), 2 * O.orc_sad_nxn(n, ptr(orig), ptr(pred)))
                assert g["cost%d" % li][b * len(modes) + m] == e, (n, b, mode)
    p = OrcQuantParams(qp, 8, 1, 0, 0, None)
    for ti, (n, cnt) in enumerate(TUS):
        ppc = (64 // n) ** 2
        for t in rng.integers(0, g["n_ctu"] * cnt, 10):
            ctu, j = divmod(int(t), cnt)
            pos = j % ppc
            X, Y = (ctu % cols) * 64 + (pos % (64 // n)) * n, (ctu // cols) * 64 + (pos // (64 // n)) * n
            top, left = refs_of(X, Y, n)
            orig = np.ascontiguousarray(plane[Y:Y + n, X:X + n]).ravel()
            pred = np.zeros(n * n, np.uint8)
            O.orc_intra_predict(n, n, int(g["mode%d" % ti][t]), ptr(top), ptr(left), ptr(pred))
            rec = np.zeros(n * n, np.uint8); co = np.zeros(n * n, np.int16)
            hc = O.orc_quantize_residual(C.byref(p), 1, 0, n, n, 0, 0, n, n, ptr(orig), ptr(pred), ptr(rec), ptr(co), 0)
            sl = slice(t * n * n, (t + 1) * n * n)
            assert g["has%d" % ti][t] == hc
            assert (g["coeff%d" % ti][sl] == co).all() and (g["rec%d" % ti][sl] == rec).all(), (n, t)
            assert g["ssd%d" % ti][t] == O.orc_pixels_calc_ssd(ptr(orig), ptr(rec), n, n, n, n)
            assert g["tcost%d" % ti][t] == O.orc_fast_coeff_cost(ptr(co), n, n, WEIGHTS)


def test_refpic_exchange_single_rank_and_nccl_binding():
    """The NCCL-backed exchange behind the C ABI on one rank: NCCL binds at run time, a 1-rank communicator comes up
    and the exchange leaves the picture untouched (the N > 1 path is checked inside bench.py --gpus N: every rank
    compares the gathered picture with the full one)."""
    import torch
    L = U.lib()
    assert L.uvg_b200_comm_nccl_version() > 20000
    uid = (C.c_uint8 * 128)()
    U.check(L.uvg_b200_comm_unique_id(uid))
    comm = L.uvg_b200_comm_create(uid, 0, 1)
    assert comm, L.uvg_b200_last_error()
    plane = torch.randint(0, 256, (130, 256), dtype=torch.uint8, device="cuda")
    before = plane.clone()
    U.check(L.uvg_b200_refpic_exchange(comm, plane.data_ptr(), 256, 130, None))
    torch.cuda.synchronize()
    assert torch.equal(plane, before)
    L.uvg_b200_comm_destroy(comm)
