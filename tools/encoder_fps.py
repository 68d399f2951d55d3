"""Encoder-level context numbers of SURVEY.md 8(d)(iii)/(b): the unmodified reference CLI (oracle/_ref) on a
synthetic clip, generic strategies (--cpuid=0) and its own SIMD strategies (--cpuid=1), all the host threads
the cgroup allows. CPU only (runs on the GPU box's host cores; no kernel of this repo is involved):
this is the reference's own end-to-end speed the CTU/s figures are set against.

  python tools/encoder_fps.py [--frames 10]
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
ap.add_argument("--frames", type=int, default=10)
args = ap.parse_args()
exe = os.path.join(ROOT, "oracle", "_ref", "bd8", "uvg266")
if not os.path.exists(exe):
    print(json.dumps({"unavailable": "oracle/_ref/bd8/uvg266 not built"})); sys.exit(0)
threads = bench.host_cores()
out = {"threads": threads, "runs": []}
d = tempfile.mkdtemp()
for (w, h, opts, name) in ((832, 480, ["--preset=ultrafast", "-p", "1"], "832x480 all-intra ultrafast"),
                           (1920, 1080, ["--preset=ultrafast", "-p", "1"], "1920x1080 all-intra ultrafast")):
    yuv = os.path.join(d, "%dx%d.yuv" % (w, h))
    make_yuv(yuv, w, h, args.frames, 8)
    ctus = ((w + 63) // 64) * ((h + 63) // 64)
    for cpuid in (0, 1):
        p = subprocess.run([exe, "-i", yuv, "--input-res", "%dx%d" % (w, h), "-o", os.path.join(d, "o.vvc"), "--frames", str(args.frames),
                            "--threads=%d" % threads, "--cpuid=%d" % cpuid] + opts, capture_output=True, text=True, timeout=1200)
        m = re.search(r"FPS:\s*([0-9.]+)", p.stderr + p.stdout)
        fps = float(m.group(1)) if m else None
        out["runs"].append({"clip": name, "frames": args.frames, "cpuid": cpuid, "fps": fps,
                            "ctus_per_s": round(fps * ctus, 1) if fps else None})
print(json.dumps(out))
