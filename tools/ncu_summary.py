"""Summarises an .ncu-rep (read here, no GPU needed) into a small markdown table under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.md "title"
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {m: hdr.index(m) for m, _ in METRICS if m in hdr}
    name_i = hdr.index("Kernel Name")
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none; cold-cache, serialised launches)", "",
             "| kernel | " + " | ".join(f"{lab}" for m, lab in METRICS if m in idx) + " |",
             "|---|" + "---|" * len(idx)]
    for r in rows[2:]:
        name = r[name_i].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("jaxib::", "")
        vals = []
        for m, lab in METRICS:
            if m not in idx:
                continue
            v, u = r[idx[m]], units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                if u == "byte":
                    f /= 1e6
                elif u == "Kbyte":
                    f /= 1e3
                elif u == "Gbyte":
                    f *= 1e3
                elif u in ("ns", "nsecond"):
                    f /= 1e3
                elif u in ("ms", "msecond"):
                    f *= 1e3
                vals.append(f"{f:.4g}")
            except ValueError:
                vals.append(v)
        lines.append(f"| {name[:48]} | " + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
