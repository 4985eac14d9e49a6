"""Per-kernel stall breakdown and hottest SASS lines from an .ncu-rep captured with --import-source on.

    python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep <kernel-regex> [min_fraction]
"""
import csv
import io
import subprocess
import sys


def I(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.012
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) < len(hdr) - 2 or r[0] in seen:
            continue
        seen.add(r[0])
        data.append(r)
    iS, isrc = hdr.index("# Samples"), hdr.index("Source")
    names = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_lg", "stall_wait", "stall_math",
             "stall_selected", "stall_not_selected", "stall_no_inst", "stall_branch_resolving", "stall_dispatch"]
    tot = max(1, sum(I(r[iS]) for r in data))
    print(rows[0][1][:90])
    print("samples", tot, {n.replace("stall_", ""): round(sum(I(r[hdr.index(n)]) for r in data) / tot, 2) for n in names})
    extra = {k: hdr.index(k) for k in ("L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared", "L2 Theoretical Sectors Global Excessive",
                                      "L2 Theoretical Sectors Global", "Instructions Executed") if k in hdr}
    print({k: sum(I(r[i]) for r in data) for k, i in extra.items()})
    il, ib = hdr.index("stall_long_sb"), hdr.index("stall_barrier")
    for idx, r in enumerate(data):
        if I(r[iS]) >= frac * tot:
            prev = [x[isrc].strip()[:36] for x in data[max(0, idx - 8):idx]
                    if any(t in x[isrc] for t in ("LDG", "LDS", "LDGSTS", "BAR", "DEPBAR", "STL", "LDL", "STS", "STG"))][-2:]
            st = {n.replace("stall_", ""): I(r[hdr.index(n)]) for n in names if I(r[hdr.index(n)]) > 0.2 * I(r[iS])}
            print(f"  {idx:5d} {I(r[iS]):5d} {r[isrc].strip()[:44]:44s} {st} <- {prev}")
    # worst shared-memory conflict lines
    if "L1 Wavefronts Shared Excessive" in extra:
        ie = extra["L1 Wavefronts Shared Excessive"]
        worst = sorted(data, key=lambda r: -I(r[ie]))[:8]
        print("  smem excessive wavefronts:", [(r[isrc].strip()[:30], I(r[ie])) for r in worst if I(r[ie]) > 0])
    if "L2 Theoretical Sectors Global Excessive" in extra:
        ie = extra["L2 Theoretical Sectors Global Excessive"]
        worst = sorted(data, key=lambda r: -I(r[ie]))[:8]
        print("  global excessive sectors:", [(r[isrc].strip()[:34], I(r[ie])) for r in worst if I(r[ie]) > 0])


if __name__ == "__main__":
    main()
