python -m pytest tests/test_gpu_frame.py tests/test_gpu_intra.py -x -q 2>&1 | tail -2; python tools/frame_prof.py --iters 5; ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/frame_launches.csv python tools/frame_prof.py --iters 1 > /dev/null 2>&1; python - <<EOF
import csv
rows=[r for r in csv.reader(open("gpurun_out/frame_launches.csv")) if len(r)>5 and r[0].isdigit()]
half=rows[len(rows)//2:]
tot=0
for r in half:
    t=float(r[-1]); tot+=t
    if "intra_rough" in r[4] or "tu_predict" in r[4]: print("%-60s %8.1f" % (r[4][:58], t/1000))
print("total us", tot/1000)
EOF
