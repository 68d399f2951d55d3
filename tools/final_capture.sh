#!/bin/bash
# Round-end evidence run (one gpurun call): full GPU test suite, bench line, ncu launch list of the bench command,
# ncu --set full captures of the step's kernels on bench-sized inputs.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 3 > gpurun_out/final_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fused_imma|fused_small|rough_staged|rough_mma' -c 12 \
    -f -o gpurun_out/final_prof python tools/prof_kernels.py 6 > gpurun_out/final_ncu_full.log 2>&1
echo captured
