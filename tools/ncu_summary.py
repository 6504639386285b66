"""Summarise an .ncu-rep (read here, without a GPU) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt --units 20000000 --unit-name particle-steps
"""
import argparse
import collections
import csv
import io
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg",
    "smsp__inst_executed_op_tma_ld.sum", "smsp__inst_executed_op_shared_atom.sum",
    "smsp__inst_executed_op_global_red.sum",
]


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--units", type=float, default=0, help="work units processed per launch")
    ap.add_argument("--unit-name", default="particle-steps")
    ap.add_argument("--bytes-per-unit", type=float, default=136.0)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    rows = ncu(a.rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# ncu summary of {a.rep}", f"# {a.note}" if a.note else "#"]
    for k, d in enumerate(data):
        name = d[hdr.index("Kernel Name")]
        lines.append(f"\n## launch {k}: {name}")
        vals = {}
        for key in KEYS + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("ratio")]:
            if key in hdr:
                v = d[hdr.index(key)]
                vals[key] = v
                lines.append(f"{key:85s} {v:>18s} {units[hdr.index(key)]}")
        try:
            dur_s = float(vals["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[
                units[hdr.index("gpu__time_duration.sum")]]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(vals["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(vals["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]]
            lines.append(f"derived: dram traffic {rd + wr:.4e} B per launch -> {(rd + wr) / dur_s / 1e9:.1f} GB/s "
                         f"(under ncu, cold cache; compare shares, not absolutes)")
            if a.units:
                lines.append(f"derived: algorithmic bytes {a.units * a.bytes_per_unit:.4e} B "
                             f"({a.bytes_per_unit:g} B x {a.units:g} {a.unit_name}); traffic/algorithmic = "
                             f"{(rd + wr) / (a.units * a.bytes_per_unit):.3f}")
                inst = float(vals["smsp__inst_executed.sum"])
                lines.append(f"derived: {inst / (a.units / 32):.1f} warp instructions per 32 {a.unit_name}")
        except Exception as e:  # noqa: BLE001
            lines.append(f"derived: n/a ({e})")
    # opcode mix of the first launch
    src = ncu(a.rep, "source")
    if len(src) > 2:
        h = src[1]
        i_s, i_e = h.index("Source"), h.index("Instructions Executed")
        ops = collections.Counter()
        for r in src[2:]:
            try:
                e = int(r[i_e])
            except Exception:  # noqa: BLE001
                continue
            s = r[i_s].strip()
            if s.startswith("@"):
                s = s.split(None, 1)[1]
            ops[s.split()[0].split(".")[0]] += e
        tot = sum(ops.values())
        lines.append("\n## SASS opcode mix (executed warp instructions, all captured launches)")
        for op, c in ops.most_common(24):
            lines.append(f"{op:10s} {c:>14d} {100.0 * c / tot:5.1f}%")
    open(a.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
