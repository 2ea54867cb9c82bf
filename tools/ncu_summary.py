"""Key metrics of `ncu --set full` captures (.ncu-rep, read here with `ncu -i ... --page raw --csv`) -> one markdown table.

    python tools/ncu_summary.py profiles/r02f_ncu_summary.md gpurun_out/r02f_ncu_*.ncu-rep"""
import csv
import io
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main(out, reps):
    rows_out = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, vals = rows[0], rows[1], rows[-1]
        ix = {h: i for i, h in enumerate(hdr)}
        name = vals[ix["Kernel Name"]]
        cells = []
        for m, _label in METRICS:
            if m in ix:
                v, u = vals[ix[m]], units[ix[m]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("")
        rows_out.append((os.path.basename(rep), name, cells))
    with open(out, "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on: key metrics (one launch each; cold caches, ncu clocks)\n\n")
        fh.write("| capture | kernel | " + " | ".join(l for _m, l in METRICS) + " |\n|---|---|" + "---|" * len(METRICS) + "\n")
        for rep, name, cells in rows_out:
            fh.write(f"| `{rep}` | `{name}` | " + " | ".join(cells) + " |\n")
    print(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
