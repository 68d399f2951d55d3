"""Per-source-line instruction counts of one kernel launch of an ncu report (needs -lineinfo and --import-source on):
  python tools/ncu_lines.py report.ncu-rep <launch index> [top N]
aggregates `ncu --page source --print-source cuda,sass --csv` to (file, line) -> warp instructions executed, samples."""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", idx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
cur = None
hdr = None
agg = {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[2] == "-":          # the per-line summary row (SASS rows carry an address)
        agg[(cur, int(r[0]))] = (int(r[ie]), int(r[si]), r[1].strip()[:110])
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(fn[:120]); print("warp instructions %d, samples %d" % (tot, tots))
for (f, ln), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% %5.2f%%s %-14s %4d  %s" % (100.0 * n / tot, 100.0 * s / tots, f, ln, src))
