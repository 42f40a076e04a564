"""Summarise an `ncu --page raw --csv` dump: one line per kernel launch with the counters the roofline discussion uses.
    ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv; python tools/ncu_summary.py gpurun_out/prof_all_raw.csv
"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "rdMB", 1e-6), ("dram__bytes_write.sum", "wrMB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%", 1), ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1), ("launch__registers_per_thread", "regs", 1),
        ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1), ("launch__cluster_size", "cluster", 1)]


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}
    print(f"{'kernel':58s} " + " ".join(f"{c[1]:>8s}" for c in COLS))
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        k = r[idx["Kernel Name"]]
        k = k.split("(")[0].replace("apx::(anonymous namespace)::", "").replace("void ", "")[:58]
        vals = []
        for metric, _, scale in COLS:
            i = idx.get(metric)
            try:
                v = float(r[i].replace(",", "")) if i is not None and r[i] != "" else float("nan")
                u = units[i] if i is not None else ""
                if metric.startswith("gpu__time_duration"):
                    v *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}.get(u, 1e-3)
                elif metric.startswith("dram__bytes"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                vals.append(f"{v:8.2f}")
            except (ValueError, TypeError):
                vals.append(f"{'-':>8s}")
        print(f"{k:58s} " + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
