#!/usr/bin/env python
"""Condense ncu outputs into the small text files kept under profiles/.
  python tools/ncu_summary.py full  gpurun_out/prof_X.ncu-rep  profiles/X_full.md
  python tools/ncu_summary.py list  gpurun_out/launches_X.csv  profiles/X_launches.md
"""
import collections, csv, io, json, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        # the persistent variants (K2H / K2R): shared-memory side and where the warps wait
        "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_ldgsts.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def full(rep, out):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of `{rep}`", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines += [f"## {d.get('Kernel Name', '?')}  (launch id {d.get('ID')})", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
        try:
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            lines.append(f"| **DRAM traffic per launch** | {(rd * mult[ur] + wr * mult[uw]) / 1e6:.1f} | MB |")
        except Exception:
            pass
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    if len(src) > 3:
        h = src[1]
        iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        op, sm = collections.Counter(), collections.Counter()
        tot = tots = 0
        for r in src[2:]:
            if len(r) < len(h) or r[0] in ("Kernel Name", "Address"):
                if r and r[0] == "Kernel Name":
                    break
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
            o = m.group(2).split(".")[0] if m else "?"
            op[o] += int(r[iE]); sm[o] += int(r[iN]); tot += int(r[iE]); tots += int(r[iN])
        lines += ["## SASS instruction mix of the first captured launch", "", "| opcode | % of executed warp instructions | % of stall samples |", "|---|---|---|"]
        for o, c in op.most_common(14):
            lines.append(f"| {o} | {100 * c / max(tot, 1):.1f} | {100 * sm[o] / max(tots, 1):.1f} |")
        lines.append(f"\nexecuted warp instructions: {tot}")
    open(out, "w").write("\n".join(lines) + "\n")


def launches(path, out):
    per = collections.OrderedDict()
    for r in csv.DictReader(l for l in open(path) if l.startswith('"')):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<unnamed>::", "", r["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:110]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0)
        n, t = per.get(name, (0, 0.0))
        per[name] = (n + 1, t + v)
    total = sum(t for _, t in per.values())
    lines = [f"# ncu launch list `{path}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)", "",
             "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / total:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
