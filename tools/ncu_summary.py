#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full --import-source on) into a small markdown summary for profiles/."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]

def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main(rep, title, launch=0):
    rows = ncu(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2 + launch]      # `launch`: which captured launch of the report
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"# {title}\n\nkernel: `{kname[:140]}`\n\nsource: `{rep.split('/')[-1]}` (ncu --set full --clock-control none --import-source on; cold-cache single replayed launch)\n")
    print("| metric | value | unit |\n|---|---|---|")
    m = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
    for k in KEYS:
        if k in m: print(f"| {k} | {m[k][0]} | {m[k][1]} |")
    try:
        rd = float(m["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_read.sum"][1]]
        wr = float(m["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m["dram__bytes_write.sum"][1]]
        print(f"\nDRAM traffic (read+write) per launch: **{(rd + wr) / 1e9:.3f} GB**")
    except Exception:
        pass
    if launch != 0:     # the source page aggregates per kernel; only printed for the first launch of a report
        return
    src = ncu(rep, "source")
    h = src[1]; ix = {n: i for i, n in enumerate(h)}
    data = [r for r in src[2:] if len(r) == len(h) and (r[ix["# Samples"]] or "0").isdigit()]   # multi-kernel reports repeat the header
    S = lambda r: int(r[ix["# Samples"]] or 0)
    tot = sum(S(r) for r in data) or 1
    agg = {n: sum(int(r[ix[n]] or 0) for r in data) for n in h if n.startswith("stall_") and "Not Issued" not in n}
    print("\nwarp-stall samples (all warps): " + ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    print("\nhottest SASS instructions (share of samples):\n")
    for r in sorted(data, key=lambda r: -S(r))[:12]:
        print(f"- {100 * S(r) / tot:.1f}%  `{r[ix['Source']].strip()[:90]}`")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
