"""Dump a fixed subset of `ncu --set full` metrics (value + unit) of the first kernel in each report to JSON.
Usage: python tools/ncu_metrics.py out.json name=report.ncu-rep[:note] ..."""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
out = {}
for arg in sys.argv[2:]:
    name, rest = arg.split("=", 1)
    path, _, note = rest.partition(":")
    if path.endswith(".csv"):   # a `--page raw --csv` dump made on the GPU box
        txt = open(path).read()
    else:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    out[name] = {"kernel": vals[ix["Kernel Name"]][:80], "capture": note,
                 "metrics": {k: {"value": vals[ix[k]], "unit": units[ix[k]]} for k in WANT if k in ix}}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps({k: v["metrics"].get("gpu__time_duration.sum") for k, v in out.items()}))
