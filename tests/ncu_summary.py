"""Summarise `ncu --set full` captures (read here, on the CPU box):  python tests/ncu_summary.py <tag>=<file.ncu-rep> ...
Writes a markdown table of selected counters per capture to stdout and merges dram traffic per launch into
profiles/ncu_traffic.json under the key given as tag ("c2/parity", "c4/parity", ...), which bench.py reports
as roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def main():
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for arg in sys.argv[1:]:
        tag, path = arg.split("=", 1)
        if path.endswith(".csv"):                     # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv
            raw = open(path).read()
        else:
            raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units, data = rows[0], rows[1], rows[2:]
        last = data[-1]                               # the last captured launch (warm)
        print(f"## {tag}  ({os.path.basename(path)}, launch {len(data)} of {len(data)} captured)")
        vals = {}
        for name, unit, v in zip(head, units, last):
            short = name.split("TPC.TriageCompute.")[-1]
            if short in KEEP or name in KEEP:
                print(f"| {short} | {v} | {unit} |")
                vals[short] = (v, unit)
        rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        if rd and wr:
            b = float(rd[0].replace(",", "")) * UNIT.get(rd[1], 1.0) + float(wr[0].replace(",", "")) * UNIT.get(wr[1], 1.0)
            traffic[tag] = {"bytes": b, "source": f"profiles: ncu --set full, {os.path.basename(path)}, dram__bytes_read.sum + dram__bytes_write.sum of one launch"}
    json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
