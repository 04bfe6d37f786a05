"""Summarise a .ncu-rep (one captured launch) into the markdown tables kept under profiles/.
   python tests/tools/ncu_summary.py gpurun_out/r1_x.ncu-rep "command line" > profiles/r1_ncu_x.md"""
import csv
import io
import re
import subprocess
import sys

rep, cmd = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|gpu__dram_throughput\.avg\.pct|"
                  r"gpu__time_duration\.sum|l1tex__throughput\.avg\.pct|launch__(block_size|grid_size|registers_per_thread$|"
                  r"shared_mem_per_block_dynamic|occupancy_limit)|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct|"
                  r"sm__cycles_elapsed\.max|sm__inst_executed_pipe_tc\.avg|sm__pipe_tc_cycles_active\.avg\.pct|sm__inst_issued\.avg\.pct|"
                  r"sm__issue_active\.avg\.pct|sm__warps_active\.avg\.pct|sm__throughput\.avg\.pct|smsp__cycles_active\.avg\.pct|"
                  r"smsp__warp_issue_stalled_(long_scoreboard|math_pipe|barrier|membar|short_scoreboard|wait)_per_warp_active|"
                  r"sm__pipe_(alu|fma|xu)_cycles_active\.avg\.pct)")
name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
print(f"# {rep.split('/')[-1].replace('.ncu-rep', '')}\n")
print(f"`ncu --set full --clock-control none --import-source on -c 1` of `{cmd}`")
print(f"kernel: `{name}`\n")
print("| metric | unit | value |\n|---|---|---|")
for h, u, v in sorted(zip(hdr, units, vals)):
    if KEEP.search(h):
        print(f"| {h} | {u} | {v} |")
