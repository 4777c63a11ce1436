"""Turn an `ncu --set full` capture of the shipped kernels into the numbers bench.py reports as `roofline.traffic`.

  gpurun:  ncu --set full --clock-control none --import-source on -k regex:'swin_attn_tc|swin_mlp_fused|gemm_conv_persistent' \
               -s <skip> -c <n> -o gpurun_out/r2_fused python profiles/one_frame.py 4k
  here:    ncu -i gpurun_out/r2_fused.ncu-rep --page raw --csv > profiles/r2/fused_ncu_raw.csv
           python profiles/ncu_traffic.py profiles/r2/fused_ncu_raw.csv profiles/r2/ncu_traffic.json

Per kernel class the LARGEST launch (by duration) is kept: dram__bytes_read.sum + dram__bytes_write.sum, duration, the tensor
pipe and DRAM utilisation ncu saw (cold cache, serialised - shares, not absolutes)."""
import csv
import json
import re
import sys

CLASSES = {"fused_attn": r"swin_attn_tc_kernel|swin_attn_fused_kernel", "fused_mlp": r"swin_mlp_fused2?_kernel", "gemm": r"gemm_conv_persistent"}
WANT = {
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes", "gpu__time_duration.sum": "duration_ns",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
}
UNIT = {"us": 1e3, "ns": 1.0, "ms": 1e6, "s": 1e9, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    rows = [r for r in rows if r and not r[0].startswith("==")]
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    kname = col.get("Kernel Name")
    out = {}
    for r in rows[2:]:
        if kname is None or len(r) <= kname:
            continue
        cls = next((c for c, pat in CLASSES.items() if re.search(pat, r[kname])), None)
        if cls is None:
            continue
        rec = {"kernel_name": r[kname][:160]}
        for metric, key in WANT.items():
            if metric in col and col[metric] < len(r):
                v = num(r[col[metric]])
                if v is not None:
                    rec[key] = v * UNIT.get(units[col[metric]], 1.0) if key.endswith("bytes") or key == "duration_ns" else v
        if "duration_ns" not in rec:
            continue
        if cls not in out or rec["duration_ns"] > out[cls]["duration_ns"]:
            out[cls] = rec
    for cls, rec in out.items():
        rec["dram_bytes_per_launch"] = rec.get("dram_read_bytes", 0.0) + rec.get("dram_write_bytes", 0.0)
        rec["launch"] = f"largest {cls} launch of one 4K frame under ncu --set full: grid {int(rec.get('grid', 0))} x {int(rec.get('block', 0))} threads, " \
                        f"{rec['duration_ns'] / 1e3:.0f} us (cold cache, serialised)"
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
