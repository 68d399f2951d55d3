"""End-to-end drop-in check: the UNMODIFIED reference encoder (oracle/_ref/bd*/uvg266) run with
the cuda strategies installed (LD_PRELOAD of the glue) must produce the same bitstream as with
its generic-C strategies (--cpuid=0). SURVEY.md section 0: bitstream identity is a valid
end-to-end parity check (md5 equal for --cpuid=0/1 and any thread count)."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_yuv(path, w, h, frames, bd=8, seed=1234):
    rng = np.random.default_rng(seed)
    dt = np.uint8 if bd == 8 else np.dtype("<u2")
    mx = (1 << bd) - 1
    with open(path, "wb") as f:
        for i in range(frames):
            y, x = np.mgrid[0:h, 0:w]
            Y = 0.5 + 0.23 * np.sin((x + 3 * i) * 0.1) * np.cos(y * 0.07) + 0.12 * np.sin((x + y) * 0.31)
            Y = np.clip((Y + rng.normal(0, 0.02, Y.shape)) * mx, 0, mx).astype(dt)
            f.write(Y.tobytes())
            c = np.clip((0.5 + 0.1 * np.sin(np.mgrid[0:h // 2, 0:w // 2][1] * 0.2 + i)) * mx, 0, mx).astype(dt)
            f.write(c.tobytes()); f.write(c[::-1].copy().tobytes())


def encode(bd, yuv, w, h, frames, opts, cuda, threads=0, groups=None, extra_env=None, patched=False):
    d = os.path.join(ROOT, "oracle", "_ref", "bd%d%s" % (bd, "p" if patched else ""))
    exe = os.path.join(d, "uvg266")
    glue = os.path.join(ROOT, "uvg266_b200", "lib", "libuvg266_b200_glue%d.so" % bd)
    if not (os.path.exists(exe) and os.path.exists(glue)):
        pytest.skip("prebuilt reference CLI / glue missing")
    out = yuv + (".cuda" if cuda else ".gen") + ".vvc"
    # --cpuid=0 in BOTH runs: every strategy that is not "cuda" is the generic one, so the comparison
    # isolates the cuda strategies (the reference's own avx2 strategies are not bit-exact with its
    # generic ones for every tool, e.g. --isp --rd=2 gives a different bitstream with --cpuid=1)
    cmd = [exe, "-i", yuv, "--input-res", "%dx%d" % (w, h), "-o", out, "--frames", str(frames),
           "--threads=%d" % threads, "--cpuid=0"] + opts
    if bd != 8:
        cmd += ["--input-bitdepth=%d" % bd]
    env = dict(os.environ)
    if cuda:
        env["LD_PRELOAD"] = glue
        env["UVG266_B200_VERBOSE"] = "1"
        if groups:
            env["UVG266_B200_GROUPS"] = groups
        if extra_env:
            env.update(extra_env)
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    launches = None
    m = re.search(r"kernel launches = (\d+)", p.stderr)
    if m:
        launches = int(m.group(1))
    return hashlib.md5(open(out, "rb").read()).hexdigest(), launches, p.stderr


CASES = [
    ("intra_ultrafast", 8, 128, 128, 2, ["--preset=ultrafast", "-p", "1"]),
    ("intra_ultrafast_signhide_mts", 8, 128, 64, 2, ["--preset=ultrafast", "-p", "1", "--signhide", "--mts=intra", "--rd=2"]),
    ("ldp_ultrafast", 8, 128, 64, 3, ["--preset=ultrafast", "--gop=lp-g4d3t1", "--me=hexbs"]),
    ("ra_bipred", 8, 64, 64, 4, ["--preset=ultrafast", "--gop=8", "--bipred"]),          # bipred_average
    ("intra_ibc", 8, 64, 64, 1, ["--preset=ultrafast", "-p", "1", "--ibc=1"]),           # crc32c_4x4 / 8x8 hashes
    ("intra_medium_rdoq", 8, 64, 64, 1, ["--preset=medium", "-p", "1"]),
    ("intra_ultrafast_10bit", 10, 128, 64, 2, ["--preset=ultrafast", "-p", "1"]),
    # JCCR: quant_cbcr_residual as one fused cuda kernel (plain chain) ...
    ("intra_jccr", 8, 128, 64, 2, ["--preset=ultrafast", "-p", "1", "--rd=2", "--jccr"]),
    ("intra_jccr_10bit_signhide", 10, 64, 64, 1, ["--preset=ultrafast", "-p", "1", "--rd=2", "--jccr", "--signhide"]),
    # ... and forwarded to the reference's composite (which runs on the cuda leaves) when LFNST is in play
    ("intra_jccr_lfnst_mts", 8, 128, 64, 2, ["--preset=ultrafast", "-p", "1", "--rd=2", "--jccr", "--lfnst", "--mts=intra"]),
    # the tool set of BASELINE config 4 (10-bit, MTS + LFNST)
    ("intra_mts_lfnst_10bit", 10, 128, 64, 2, ["--preset=ultrafast", "-p", "1", "--rd=2", "--mts=intra", "--lfnst"]),
    # transform skip: use_trskip = 1 / tr_idx == MTS_SKIP inside the fused kernels, luma and chroma
    ("intra_trskip", 8, 128, 64, 2, ["--preset=ultrafast", "-p", "1", "--transform-skip", "--rd=2"]),
    ("intra_trskip_chroma_mts", 8, 128, 64, 1, ["--preset=ultrafast", "-p", "1", "--transform-skip", "--chroma-transform-skip",
                                                "--rd=2", "--mts=intra"]),
    # LMCS chroma residual scaling: quantize_residual forwarded to the reference's composite, leaves on the GPU
    ("intra_lmcs", 8, 128, 64, 2, ["--preset=ultrafast", "-p", "1", "--lmcs"]),
    # ISP CUs with explicit MTS and no RDOQ: the fused kernel must pick DST-7 for them (uvg_get_tr_type :2522)
    ("intra_isp_mts", 8, 128, 64, 2, ["--preset=ultrafast", "-p", "1", "--isp", "--mts=intra", "--rd=2"]),
    # motion search patterns of BASELINE config 3 (TZ) and the exhaustive one
    ("ldp_tz", 8, 192, 128, 3, ["--preset=ultrafast", "--gop=lp-g4d3t1", "--me=tz"]),
    ("ldp_full8", 8, 128, 64, 2, ["--preset=ultrafast", "--gop=lp-g4d3t1", "--me=full8"]),
    # variance-adaptive quantisation: pixel_var (doubles, index-order accumulation) on the picture and on every CTU
    ("intra_vaq", 8, 128, 128, 2, ["--preset=ultrafast", "-p", "1", "--vaq=5"]),
    ("intra_vaq_10bit", 10, 128, 64, 1, ["--preset=ultrafast", "-p", "1", "--vaq=5"]),
]

# One encode per BASELINE.json config at a frame size that has all four frame edges, interior CTUs and partial
# CTUs at the right / bottom (832x480 = 13 x 7.5 CTUs), run with the encoder's worker threads so that the leaf
# calls overlap. Scaling lists cannot get a case: the reference CLI itself dies with SIGFPE for
# --scaling-list=default in every configuration tried (generic and avx2 alike), before any strategy is called.
CONFIG_CASES = [
    ("cfg1_intra_ultrafast_832x480", 8, 832, 480, 1, ["--preset=ultrafast", "-p", "1"], 8),
    ("cfg2_intra_medium_832x480", 8, 832, 480, 1, ["--preset=medium", "-p", "1"], 8),
    ("cfg3_ldp_tz_832x480", 8, 832, 480, 2, ["--preset=ultrafast", "--gop=lp-g4d3t1", "--me=tz"], 8),
    ("cfg4_intra_10bit_mts_lfnst_832x480", 10, 832, 480, 1, ["--preset=ultrafast", "-p", "1", "--mts=intra", "--lfnst", "--rd=2"], 8),
    ("cfg5_ra_gop16_tiles_832x480", 8, 832, 480, 3, ["--preset=ultrafast", "--gop=16", "--bipred", "--tiles=2x2", "--owf=1"], 8),
    ("tiles_wpp_threads", 8, 256, 128, 2, ["--preset=ultrafast", "-p", "1", "--tiles=2x2", "--owf=1"], 4),
]


@pytest.mark.parametrize("name,bd,w,h,frames,opts", CASES, ids=[c[0] for c in CASES])
def test_bitstream_identity(tmp_path, name, bd, w, h, frames, opts):
    yuv = str(tmp_path / (name + ".yuv"))
    make_yuv(yuv, w, h, frames, bd)
    md5_gen, _, _ = encode(bd, yuv, w, h, frames, opts, cuda=False)
    md5_cuda, launches, err = encode(bd, yuv, w, h, frames, opts, cuda=True)
    assert launches and launches > 100, err[-1500:]
    assert md5_cuda == md5_gen, "bitstream differs with the cuda strategies (%s)" % name


@pytest.mark.parametrize("name,bd,w,h,frames,opts,threads", CONFIG_CASES, ids=[c[0] for c in CONFIG_CASES])
def test_bitstream_identity_baseline_configs(tmp_path, name, bd, w, h, frames, opts, threads):
    yuv = str(tmp_path / (name + ".yuv"))
    make_yuv(yuv, w, h, frames, bd)
    md5_gen, _, _ = encode(bd, yuv, w, h, frames, opts, cuda=False, threads=threads)
    md5_cuda, launches, err = encode(bd, yuv, w, h, frames, opts, cuda=True, threads=threads)
    assert launches and launches > 1000, err[-1500:]
    assert md5_cuda == md5_gen, "bitstream differs with the cuda strategies (%s)" % name


@pytest.mark.parametrize("bd", [8, 10])
def test_lfnst_blocks_stay_in_the_fused_kernel(tmp_path, bd):
    """Row 8(f)-3 stage 1 at the encoder level (the tool set of BASELINE config 4): with the encoder's LFNST kernels
    handed to the device, an LFNST TU is ONE launch (residual .. forward LFNST .. quant .. inverse LFNST ..
    reconstruction); with UVG266_B200_LFNST=host the same TU walks the reference's composite leaf by leaf. Same
    bitstream all three ways, far fewer launches."""
    yuv = str(tmp_path / "lfnst.yuv")
    make_yuv(yuv, 128, 64, 2, bd)
    opts = ["--preset=ultrafast", "-p", "1", "--rd=2", "--mts=intra", "--lfnst"]
    md5_gen, _, _ = encode(bd, yuv, 128, 64, 2, opts, cuda=False)
    md5_dev, launches_dev, err = encode(bd, yuv, 128, 64, 2, opts, cuda=True)
    md5_host, launches_host, _ = encode(bd, yuv, 128, 64, 2, opts, cuda=True, extra_env={"UVG266_B200_LFNST": "host"})
    assert md5_dev == md5_gen and md5_host == md5_gen, err[-1500:]
    assert launches_dev and launches_host and launches_dev < 0.8 * launches_host, (launches_dev, launches_host)


@pytest.mark.parametrize("groups,threads", [("hook", 0), ("hook", 4), ("picture,dct,quant,hook", 4)])
def test_patched_call_site_is_bitstream_identical(tmp_path, groups, threads):
    """Row 8(f)-4: the reference built with glue/patches/search_intra_rough.patch scores all 67 modes of every square
    8 / 16 / 32 luma CU of search_intra_rough through ONE uvg_b200_intra_rough_host call; the bitstream must be the
    one the unmodified encoder writes."""
    yuv = str(tmp_path / "hook.yuv")
    make_yuv(yuv, 320, 192, 2)
    opts = ["--preset=ultrafast", "-p", "1"]
    md5_gen, _, _ = encode(8, yuv, 320, 192, 2, opts, cuda=False)
    md5_patched_off, _, _ = encode(8, yuv, 320, 192, 2, opts, cuda=False, patched=True)
    md5_hook, launches, err = encode(8, yuv, 320, 192, 2, opts, cuda=True, threads=threads, groups=groups, patched=True)
    assert md5_patched_off == md5_gen
    assert launches and launches > 50, err[-1500:]
    assert md5_hook == md5_gen, "bitstream differs with the batched call site (%s)" % groups
    # a slower preset walks more CU sizes and the full rough-search refinement
    opts = ["--preset=medium", "-p", "1"]
    md5_gen, _, _ = encode(8, yuv, 320, 192, 1, opts, cuda=False)
    md5_hook, launches, err = encode(8, yuv, 320, 192, 1, opts, cuda=True, threads=threads, groups=groups, patched=True)
    assert launches and md5_hook == md5_gen, err[-1500:]


def test_multithreaded_encode_is_identical(tmp_path):
    """Leaf functions are called from several worker threads at once (threadqueue.c:275)."""
    yuv = str(tmp_path / "mt.yuv")
    make_yuv(yuv, 192, 128, 2)
    opts = ["--preset=ultrafast", "-p", "1"]
    md5_gen, _, _ = encode(8, yuv, 192, 128, 2, opts, cuda=False)
    md5_cuda, launches, err = encode(8, yuv, 192, 128, 2, opts, cuda=True, threads=4)
    assert launches and launches > 100
    assert md5_cuda == md5_gen


INTRA_CASES = [
    ("intra_only_group", 8, 64, 64, 1, ["--preset=ultrafast", "-p", "1"], "intra"),
    ("all_groups_medium", 8, 64, 64, 1, ["--preset=medium", "-p", "1"], "picture,dct,quant,intra"),
    ("intra_mrl_isp_10bit", 10, 64, 64, 1, ["--preset=ultrafast", "-p", "1", "--mrl", "--isp", "--rd=2"], "intra"),
    # row 8(f)-1: the MC sample-interpolation leaves (uni- and bi-prediction, luma and chroma)
    ("ipol_ldp", 8, 128, 64, 3, ["--preset=ultrafast", "--gop=lp-g4d3t1", "--me=hexbs"], "ipol"),
    ("ipol_ra_bipred_10bit", 10, 64, 64, 4, ["--preset=ultrafast", "--gop=8", "--bipred"], "ipol"),
    ("all_five_groups_ldp", 8, 64, 64, 2, ["--preset=ultrafast", "--gop=lp-g4d3t1"], "picture,dct,quant,intra,ipol"),
]


@pytest.mark.parametrize("name,bd,w,h,frames,opts,groups", INTRA_CASES, ids=[c[0] for c in INTRA_CASES])
def test_bitstream_identity_with_intra_leaves(tmp_path, name, bd, w, h, frames, opts, groups):
    """Row 8(f)-2 at the drop-in boundary: angular_pred / intra_pred_planar / pdpc_planar_dc served by the
    GPU for every block the encoder predicts (luma and chroma, all shapes, wide angles, MRL / ISP)."""
    yuv = str(tmp_path / (name + ".yuv"))
    make_yuv(yuv, w, h, frames, bd)
    md5_gen, _, _ = encode(bd, yuv, w, h, frames, opts, cuda=False)
    md5_cuda, launches, err = encode(bd, yuv, w, h, frames, opts, cuda=True, groups=groups)
    assert launches and launches > 100, err[-1500:]
    assert md5_cuda == md5_gen, "bitstream differs with the cuda intra strategies (%s)" % name
