"""Copy the round-2 evidence from gpurun_out/ into profiles/ (tracked) and derive the small summaries."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for src, dst in (("r2_bench.json", "r2_bench.json"), ("r2_bench_reference_arm.json", "r2_bench_reference_arm.json"),
                 ("r2_pytest_gpu.txt", "r2_pytest_gpu.txt"), ("r2_launches.csv", "r2_launches.csv"),
                 ("bench_r2_n2.json", "r2_bench_2gpu.json"), ("r2_frame_device_ms.txt", "r2_frame_device_ms.txt"),
                 ("r2_frame_kernel_times.txt", "r2_frame_kernel_times.txt"), ("r2_kernel_sweep_transform.json", "r2_kernel_sweep_transform.json"),
                 ("r2_kernel_sweep_transform_fast.json", "r2_kernel_sweep_transform_fast.json"), ("r2_e2e_timeline.txt", "r2_e2e_timeline.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
py = sys.executable
if os.path.exists(os.path.join(G, "r2_launches.csv")):
    open(os.path.join(P, "r2_launch_shares.txt"), "w").write(
        subprocess.run([py, os.path.join(ROOT, "tools", "launch_shares.py"), os.path.join(G, "r2_launches.csv")], capture_output=True, text=True).stdout)
traffic = {}
for rep, out in (("r2_prof.ncu-rep", "r2_ncu_full_summary.txt"), ("r2_prof_frame.ncu-rep", "r2_ncu_frame_summary.txt")):
    rp = os.path.join(G, rep)
    if not os.path.exists(rp):
        continue
    open(os.path.join(P, out), "w").write(
        subprocess.run([py, os.path.join(ROOT, "tools", "ncu_summary.py"), rp], capture_output=True, text=True).stdout)
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    k = 34.0 / 12.0 if rep == "r2_prof.ncu-rep" else 1.0      # prof_kernels.py ran on 12 of the picture's 34 CTU rows
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void uvgb::", "").replace("uvgb::", "")
        b = float(r[rd].replace(",", "")) * scale.get(units[rd], 1) + float(r[wr].replace(",", "")) * scale.get(units[wr], 1)
        traffic.setdefault(name, []).append(b * k)
if traffic:
    json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, bench-sized inputs; tools/r2_capture.sh); "
                       "the rough / fused kernels were captured on 12 of the 34 CTU rows of a picture and are scaled by 34/12 to a whole picture (their traffic is linear in the number of blocks), the frame-pass kernels on a whole picture",
               "bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}}, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
print("profiles updated:", sorted(f for f in os.listdir(P) if f.startswith("r2_")))
