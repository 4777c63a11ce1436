#!/bin/bash
# Round-2 closing measurements on one B200 (run under gpurun from the repo root): GPU test suite, the bench line, the ncu launch
# list of the bench command and ncu --set full captures of the three shipped tcgen05 kernel classes.  Outputs -> gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 600 gpurun_out/bench_r2_final.err
NB200_BENCH_MINIMAL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench_step.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
for k in swin_attn_tc; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 10 -f -o gpurun_out/r2_$k python profiles/one_frame.py 4k > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -15
