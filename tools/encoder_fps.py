"""Encoder-level numbers: the reference CLI end to end on a synthetic
all-intra clip, four ways on the same host threads -
  generic        unmodified reference, --cpuid=0
  avx2           unmodified reference, its own SIMD strategies (--cpuid=1)
  cuda-dropin    unmodified reference + LD_PRELOAD of the strategy glue: every leaf call (SAD/SATD/DCT/quant ...) is
                 one kernel launch + two PCIe copies + a sync - the parity vehicle, not a throughput path
  cuda-patched   reference built with glue/patches/search_intra_rough.patch + glue in hook mode: the rough intra search
                 of every square 8/16/32 luma CU is ONE launch (all 67 modes), everything else stays on the CPU
                 (avx2 strategies, or generic with --cpuid=0)
Bitstreams of all four are identical for --cpuid=0 (tests/test_gpu_encoder_e2e.py); fps = the CLI's own "FPS:" line.

  python tools/encoder_fps.py [--frames 2] [--size 3840x2160] [--no-gpu]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from test_gpu_encoder_e2e import make_yuv

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--size", default="3840x2160")
ap.add_argument("--no-gpu", action="store_true")
ap.add_argument("--dropin-frames", type=int, default=1, help="frames for the (slow) drop-in leg")
args = ap.parse_args()
w, h = (int(v) for v in args.size.split("x"))
exe = os.path.join(ROOT, "oracle", "_ref", "bd8", "uvg266")
exe_p = os.path.join(ROOT, "oracle", "_ref", "bd8p", "uvg266")
glue = os.path.join(ROOT, "uvg266_b200", "lib", "libuvg266_b200_glue8.so")
if not os.path.exists(exe):
    print(json.dumps({"unavailable": "oracle/_ref/bd8/uvg266 not built"})); sys.exit(0)
threads = bench.host_cores()
out = {"threads": threads, "clip": "%dx%d all-intra ultrafast, synthetic" % (w, h), "runs": []}
d = tempfile.mkdtemp()
yuv = os.path.join(d, "clip.yuv")
make_yuv(yuv, w, h, args.frames, 8)
ctus = ((w + 63) // 64) * ((h + 63) // 64)
opts = ["--preset=ultrafast", "-p", "1"]


def run(name, binary, cpuid, frames, env_extra=None):
    env = dict(os.environ)
    if env_extra:
        env.update(env_extra)
    p = subprocess.run([binary, "-i", yuv, "--input-res", "%dx%d" % (w, h), "-o", os.path.join(d, name + ".vvc"), "--frames", str(frames),
                        "--threads=%d" % threads, "--cpuid=%d" % cpuid] + opts, capture_output=True, text=True, timeout=3000, env=env)
    m = re.search(r"FPS:\s*([0-9.]+)", p.stderr + p.stdout)
    fps = float(m.group(1)) if m else None
    lm = re.search(r"kernel launches = (\d+)", p.stderr)
    import hashlib
    md5 = hashlib.md5(open(os.path.join(d, name + ".vvc"), "rb").read()).hexdigest() if p.returncode == 0 else None
    out["runs"].append({"variant": name, "frames": frames, "cpuid": cpuid, "fps": fps, "ctus_per_s": round(fps * ctus, 1) if fps else None,
                        "kernel_launches": int(lm.group(1)) if lm else 0, "md5": md5, "rc": p.returncode})


run("generic", exe, 0, args.frames)
run("avx2", exe, 1, args.frames)
if not args.no_gpu and os.path.exists(glue):
    pre = {"LD_PRELOAD": glue, "UVG266_B200_VERBOSE": "1"}
    run("cuda-dropin (picture,dct,quant leaves; cpuid=0)", exe, 0, args.dropin_frames, pre)
    if os.path.exists(exe_p):
        run("cuda-patched (rough-search hook only; rest generic)", exe_p, 0, args.frames, dict(pre, UVG266_B200_GROUPS="hook"))
        run("cuda-patched (rough-search hook only; rest avx2)", exe_p, 1, args.frames, dict(pre, UVG266_B200_GROUPS="hook"))
g = [r for r in out["runs"] if r["variant"] == "generic"]
if g and g[0]["md5"]:
    for r in out["runs"]:
        if r["cpuid"] == 0 and r["frames"] == g[0]["frames"]:
            r["same_bitstream_as_generic"] = r["md5"] == g[0]["md5"]
print(json.dumps(out))
