"""Summarise an `ncu --set full` report (read here with `ncu -i rep --page raw --csv`) into one CSV row per kernel
launch: duration, DRAM bytes, tensor-pipe and DRAM utilisation, occupancy, registers -- the file kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv && python tools/ncu_summary.py /tmp/raw.csv > profiles/xyz.csv
"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    out = csv.writer(sys.stdout)
    present = [(m, label) for m, label in WANT if m in col]
    out.writerow(["kernel"] + [f"{label} [{units[col[m]]}]" for m, label in present])
    for r in data:
        name = r[col["Kernel Name"]]
        name = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        out.writerow([name] + [r[col[m]] for m, _ in present])


if __name__ == "__main__":
    main()
