"""CPU reference arm scaling sweep (threads x pinning) on the host; no GPU needed."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2040
arm = bench.CpuArm(n)
cores = bench.host_cores()
print("cores", cores, "ctus", n)
for pin, static in (("1", "1"), ("1", "0"), ("0", "1")):
    os.environ["REFH_PIN"] = pin
    os.environ["REFH_STATIC"] = static
    print("pin", pin, "static", static)
    for th in [32, 64, cores]:
        arm.run(th)
        ts = [arm.run(th) for _ in range(6)]
        print("pin", pin, th, "threads: CTU/s", [round(n / t) for t in ts], "mean", round(n * len(ts) / sum(ts)))
