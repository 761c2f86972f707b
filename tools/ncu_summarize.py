"""Condense an .ncu-rep into a small text summary (raw metrics, opcode mix, stall mix, hottest
SASS lines).  Run where ncu is available; the summary is what gets committed under profiles/."""
import collections, csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fp64",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct", "smsp__average_warps_issue_stalled",
        "local_load", "local_store", "smsp__sass_inst_executed_op_local", "sm__cycles_active.avg",
        "launch__shared_mem_per_block", "smsp__inst_executed_op_shared"]


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main(rep, out):
    lines = []
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    if len(raw) >= 3:
        hdr, units = raw[0], raw[1]
        for row in raw[2:]:
            lines.append("== kernel: " + row[hdr.index("Kernel Name")][:150])
            for h, u, v in zip(hdr, units, row):
                if any(k in h for k in KEYS):
                    lines.append(f"  {h} [{u}] = {v}")
    src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hi = next((i for i, r in enumerate(src) if "Source" in r and "# Samples" in r), None)
    if hi is not None:
        hdr = src[hi]
        idx = {h: i for i, h in enumerate(hdr)}
        rows = [r for r in src[hi + 1:] if len(r) == len(hdr)]
        tot_s = sum(int(r[idx["# Samples"]]) for r in rows) or 1
        tot_i = sum(int(r[idx["Instructions Executed"]]) for r in rows) or 1
        ops, ops_s = collections.Counter(), collections.Counter()
        for r in rows:
            t = r[idx["Source"]].split()
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[op] += int(r[idx["Instructions Executed"]])
            ops_s[op] += int(r[idx["# Samples"]])
        lines.append(f"== SASS: {len(rows)} lines, {tot_i} warp-instructions executed, {tot_s} stall samples")
        lines.append("  opcode mix by executed instructions (%): " +
                     ", ".join(f"{k} {100 * v / tot_i:.1f}" for k, v in ops.most_common(24)))
        lines.append("  opcode mix by stall samples (%): " +
                     ", ".join(f"{k} {100 * v / tot_s:.1f}" for k, v in ops_s.most_common(16)))
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[idx[h]] or 0) for r in rows) for h in stalls}
        tot = sum(agg.values()) or 1
        lines.append("  stall reasons (% of samples): " +
                     ", ".join(f"{k[6:]} {100 * v / tot:.1f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
        top = sorted(rows, key=lambda r: -int(r[idx["# Samples"]]))[:25]
        lines.append("  hottest SASS lines (samples, executed, instruction):")
        for r in top:
            lines.append(f"    {r[idx['# Samples']]:>7} {r[idx['Instructions Executed']]:>10}  {r[idx['Source']].strip()[:110]}")
        tma = sum(int(r[idx["Instructions Executed"]]) for r in rows if "UBLKCP" in r[idx["Source"]])
        lines.append(f"  UBLKCP (TMA bulk copy) instructions executed: {tma}")
    # source-level view (needs -lineinfo and --import-source on): per-source-line aggregates
    cu = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    agg, fname, hdr = [], "?", None
    for r in cu:
        if len(r) >= 2 and r[0] in ("File Path", "File Name"):
            fname = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr is not None and len(r) == len(hdr) and r[2] == "-":
            i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            agg.append((fname, r[0], r[1].strip(), int(r[i_inst] or 0), int(r[i_samp] or 0)))
    if agg:
        tot_i = sum(a[3] for a in agg) or 1
        tot_s = sum(a[4] for a in agg) or 1
        lines.append(f"== CUDA source lines with metrics: {len(agg)}")
        lines.append("  top source lines by executed warp-instructions (% instr, % stall samples, file:line, text):")
        for a in sorted(agg, key=lambda a: -a[3])[:80]:
            lines.append(f"    {100 * a[3] / tot_i:5.2f} {100 * a[4] / tot_s:5.2f}  {a[0]}:{a[1]:<5} {a[2][:110]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
