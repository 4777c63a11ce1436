#!/bin/bash
# ZoeD_N workload (BASELINE configs[4]) evidence on one B200: pipeline tests, both iw3 bench lines, the ncu launch list of one
# 4K batch and ncu --set full of its two dominant kernels.  Outputs -> gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_video_pipeline.py tests/test_gpu_zoedepth.py -q > gpurun_out/zoe_pytest4.log 2>&1; tail -3 gpurun_out/zoe_pytest4.log
timeout 400 python bench.py --workload iw3_4k_zoe --steps 20 --warmup 3 > gpurun_out/bench_iw3_4k_zoe.json 2> gpurun_out/bench_iw3_4k_zoe.err; tail -c 800 gpurun_out/bench_iw3_4k_zoe.err
timeout 300 python bench.py --workload iw3_1080p --steps 20 --warmup 3 > gpurun_out/bench_iw3_1080p.json 2> gpurun_out/bench_iw3_1080p.err; tail -c 800 gpurun_out/bench_iw3_1080p.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_zoe_4k.csv python profiles/one_frame.py 4k zoe > gpurun_out/ncu_zoe_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_conv_persistent|flash_attention' -s 330 -c 24 -f -o gpurun_out/r2_zoe python profiles/one_frame.py 4k zoe > gpurun_out/ncu_zoe_full.log 2>&1
ls -la gpurun_out | tail -12
