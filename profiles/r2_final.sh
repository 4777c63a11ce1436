#!/bin/bash
# Round-2 closing measurements on one B200 (run under gpurun from the repo root): GPU test suite, smoke, the bench line, the
# eager-PyTorch-on-the-same-GPU arm.  Outputs -> gpurun_out/.  (The ncu launch list of the bench step and the ncu --set full
# captures of the shipped kernels were taken earlier in the round with the same kernels: profiles/r2/final/, profiles/r2/.)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; tail -4 gpurun_out/r2_smoke.log
timeout 600 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 600 gpurun_out/bench_r2_final.err
timeout 300 python bench.py --impl torch_gpu --steps 2 --warmup 1 > gpurun_out/bench_torch_gpu.json 2> gpurun_out/bench_torch_gpu.err; tail -c 400 gpurun_out/bench_torch_gpu.err
ls -la gpurun_out | tail -8
