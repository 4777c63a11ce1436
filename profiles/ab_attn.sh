export NB200_BENCH_MINIMAL=1
python bench.py --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
NB200_TUNE=12=1 python bench.py --no-cpu-baseline > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err
python - <<PY
import json
for n in ["tc","old"]:
    try:
        d=json.load(open("gpurun_out/bench_%s.json"%n))
        print(n, round(d["value"],2), round(d["ms_per_step"],2), d["kernel_classes_ms"], d.get("clocks"))
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/bench_%s.err"%n).read()[-800:])
PY
