"""SASS census: per kernel of libuvg266_b200.so, how many instructions of the families that tell
the hardware path apart - 5th-gen tensor core (UTC*MMA, LDTM/STTM = tcgen05 + TMEM), TMA (UTMALDG/UTMASTG tensor-map
copies, UBLKCP bulk copies), legacy tensor core (IMMA/HMMA = mma.sync), packed ALU (IDP.4A/2A, VABSDIFF4, PRMT),
async copies (LDGSTS) and mbarrier traffic (SYNCS). Runs on the CPU box (cuobjdump only needs the objects).

  python tools/sass_census.py > profiles/r2_sass_census.txt
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = [("UTCMMA", r"\bUTC[A-Z]*MMA"), ("LDTM/STTM", r"\b(LDTM|STTM)"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
            ("UBLKCP", r"\bUBLKCP"), ("IMMA", r"\bIMMA"), ("HMMA", r"\bHMMA"), ("IDP", r"\bIDP\."), ("VABSDIFF", r"\bVABSDIFF"),
            ("PRMT", r"\bPRMT"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("REDUX", r"\bREDUX"), ("DADD/DMUL", r"\b(DADD|DMUL|DFMA)")]
rows = []
for obj in sorted(glob.glob(os.path.join(ROOT, "uvg266_b200", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts, total = None, None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                rows.append((os.path.basename(obj), cur, total, counts))
            cur, counts, total = m.group(1), collections.Counter(), 0
            continue
        if cur and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
            total += 1
            for name, pat in FAMILIES:
                if re.search(pat, line):
                    counts[name] += 1
    if cur:
        rows.append((os.path.basename(obj), cur, total, counts))


def demangle(n):
    try:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        d = d.replace("(anonymous namespace)::", "").replace("uvgb::", "")
        return d.split("(")[0]
    except OSError:
        return n


print("# per-kernel SASS instruction counts (static), sm_100a; columns: total | " + " | ".join(n for n, _ in FAMILIES))
print("object | kernel | total | " + " | ".join(n for n, _ in FAMILIES))
for obj, fn, total, c in rows:
    print("%s | %s | %d | %s" % (obj, demangle(fn)[:90], total, " | ".join(str(c.get(n, 0)) for n, _ in FAMILIES)))
print()
tot = collections.Counter()
for _, _, _, c in rows:
    tot.update(c)
print("# library totals: " + ", ".join("%s %d" % (n, tot.get(n, 0)) for n, _ in FAMILIES))
