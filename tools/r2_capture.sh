#!/bin/bash
# Round-2 evidence run (one gpurun call): full GPU test suite, both bench arms, ncu launch list of the bench command,
# (the ncu --set full captures are tools/r2_capture_ncu.sh, run before this one).
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_pytest_gpu.txt
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-v1 --e2e-steps 3 > gpurun_out/r2_ncu_bench.log 2>&1
python tools/frame_prof.py --iters 5 > gpurun_out/r2_frame_device_ms.txt 2>&1
bash tools/frame_times.sh > gpurun_out/r2_frame_kernel_times.txt 2>&1
# (transform sweep: tools/kernel_sweep.py --only transform, with and without UVG266_B200_TRANSFORM=fast; profiles/r2_kernel_sweep_transform*.json)
for n in 1 2 3; do python tools/e2e_timeline.py --inflight $n --steps 10 2>/dev/null | head -1; done > gpurun_out/r2_e2e_timeline.txt
UVG266_B200_FRAME_PUSH=1 python tools/e2e_timeline.py --inflight 2 --steps 10 2>/dev/null | head -1 | sed 's/^/push kernel: /' >> gpurun_out/r2_e2e_timeline.txt
echo captured
