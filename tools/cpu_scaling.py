"""CPU reference arm scaling sweep (threads x sample size) on the host; no GPU needed."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2040
arm = bench.CpuArm(n)
cores = bench.host_cores()
print("cores", cores, "ctus", n)
for th in [1, 8, 16, 32, 64, 96, cores, 2 * cores]:
    if th > 2 * cores:
        continue
    arm.run(th)
    ts = [arm.run(th) for _ in range(5)]
    print(th, "threads: CTU/s", [round(n / t) for t in ts])
