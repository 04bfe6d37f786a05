"""Summarise an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
of bench.py into (a) a markdown table per kernel for ONE whole step - the launches between two consecutive mc_reduce_kernel
launches - and (b) profiles/traffic.json (mean DRAM bytes per launch of the tcgen05 kernels: bench.py's roofline.traffic).
   python tests/tools/launch_list_summary.py gpurun_out/r2_launches_bench.csv profiles/r2_bench_launch_list.csv.gz profiles/traffic.json
"""
import csv
import gzip
import json
import re
import sys
from collections import OrderedDict, defaultdict

src, out_gz, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(src, errors="replace") if l.startswith('"')]
rows = list(csv.DictReader(lines))
launches = OrderedDict()
for r in rows:
    d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    unit = r["Metric Unit"]
    if "time" in r["Metric Name"]:
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r["Metric Name"]] = v
ids = list(launches)
ends = [i for i in ids if "mc_reduce_kernel" in launches[i]["name"]]
if len(ends) < 2:
    raise SystemExit(f"need two mc_reduce_kernel launches to delimit a step, found {len(ends)} in {len(ids)} launches")
step = [i for i in ids if ends[-2] < i <= ends[-1]]


def short(n):
    n = re.sub(r"^void ", "", n)
    n = n.replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_:]+(?:<[^(]*>)?)", n)
    return (m.group(1) if m else n)[:60]


agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i in step:
    d = launches[i]
    a = agg[short(d["name"])]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"one step = {len(step)} launches, {tot:.2f} ms of serialised kernel time, "
      f"{sum(a[2] for a in agg.values()) / 1e12:.3f} TB read + {sum(a[3] for a in agg.values()) / 1e12:.3f} TB written\n")
print("| kernel | launches | total ms | share | DRAM read GB | DRAM write GB | GB/s |\n|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e9:.2f} | {a[3] / 1e9:.2f} | {(a[2] + a[3]) / max(a[1], 1e-9) / 1e6:.0f} |")
tc = [a for k, a in agg.items() if "gemm_f16_tc_kernel" in k or "conv3x3_c64_stream" in k or "stem_conv_pool" in k]
n_tc = sum(a[0] for a in tc)
traffic = sum(a[2] + a[3] for a in tc) / max(n_tc, 1)
share = sum(a[1] for a in tc) / tot
print(f"\ntcgen05 kernels: {n_tc} launches, {100 * share:.1f}% of the captured time, mean DRAM traffic per launch {traffic / 1e9:.3f} GB")
json.dump({"tcgen05_dram_bytes_per_launch": traffic, "tcgen05_launches_per_step": n_tc, "tcgen05_share_of_kernel_time": share,
           "launches_per_step": len(step), "kernel_ms_serialised": tot, "source": out_gz.split("/")[-1]}, open(out_json, "w"), indent=1)
with gzip.open(out_gz, "wt") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "time_ms", "dram_read_bytes", "dram_write_bytes"])
    for i in step:
        d = launches[i]
        w.writerow([i, short(d["name"]), d["grid"], f"{d.get('gpu__time_duration.sum', 0.0):.6f}",
                    int(d.get("dram__bytes_read.sum", 0)), int(d.get("dram__bytes_write.sum", 0))])
