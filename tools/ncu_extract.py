#!/usr/bin/env python
"""Selected metrics of an ncu report -> profiles/<name>_metrics.csv (one column per captured launch) and, for the chain
kernels, profiles/ncu_traffic.json (DRAM bytes per row, read by bench.py for `roofline.traffic`).

    ncu --set full --clock-control none --import-source on -k regex:chain_tc4 -s 6 -c 2 -o gpurun_out/tc4 \\
        python bench.py --rows 20000000 --steps 1 --no-e2e --no-cpu --no-train --no-cfd
    python tools/ncu_extract.py gpurun_out/tc4.ncu-rep profiles/r01_chain_tc4_metrics.csv --rows 20000000 --traffic
"""
import argparse
import csv
import io
import json
import re
import subprocess

KEEP = re.compile(r"^(Kernel Name|.*sm__pipe_tensor_cycles_active.*|dram__bytes_(read|write)\.sum(\.pct_of_peak_sustained_elapsed|"
                  r"\.per_second)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|launch__(block_size|grid_size|"
                  r"registers_per_thread|registers_per_thread_allocated|shared_mem_per_block_dynamic)|"
                  r"sm__inst_executed_pipe_(alu|fma|tmem|uniform)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
                  r"smsp__cycles_active\.avg|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active)$")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_csv")
    ap.add_argument("--rows", type=int, default=0, help="rows per captured launch (for --traffic)")
    ap.add_argument("--traffic", action="store_true", help="also write profiles/ncu_traffic.json (launch 0 = encode, 1 = decode)")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    names, units, launches = rows[0], rows[1], rows[2:]
    with open(a.out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(launches))])
        for c, n in enumerate(names):
            if KEEP.match(n):
                w.writerow([n, units[c]] + [l[c] for l in launches])
    if a.traffic:
        col = {n: c for c, n in enumerate(names)}

        def gbytes(l, n):
            v, u = float(l[col[n]].replace(",", "")), units[col[n]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]

        per_row = [(gbytes(l, "dram__bytes_read.sum") + gbytes(l, "dram__bytes_write.sum")) / a.rows for l in launches[:2]]
        json.dump({"encode_dram_bytes_per_row": round(per_row[0], 4), "decode_dram_bytes_per_row": round(per_row[1], 4),
                   "source": "%s (ncu --set full --clock-control none, %d-row launches of chain_tc4_kernel; dram__bytes_read.sum + "
                             "dram__bytes_write.sum)" % (a.out_csv, a.rows), "rows": a.rows},
                  open("profiles/ncu_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main()
