"""Text summary of an .ncu-rep capture (one kernel launch, `ncu --set full --import-source on`) for profiles/:
python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x_ncu_full.txt ["free-text note"]
Prints the counters the DESIGN / VERDICT arguments rest on, the warp-stall breakdown, and -- from the SASS page -- where the
issued instructions and the stall samples sit (straight-line blocks of equal execution count)."""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed"]


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [f"# {rep.split('/')[-1]}: ncu --set full --clock-control none --import-source on (one launch)"]
    if note:
        lines.append(f"# {note}")
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"kernel: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in RAW:
            if k in d and d[k] != "":
                lines.append(f"  {k:86s} {d[k]} {u[k]}")
        lines.append("  warp stalls (warps per issue-active cycle):")
        st = sorted(((float(d[k]), k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k] not in ("", None)), reverse=True)
        for v, k in st[:9]:
            lines.append(f"    {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.3f}")
    rows = ncu(rep, "source", ["--print-source", "sass"])
    if len(rows) > 2:
        hdr = rows[1]
        try:
            isrc, ie, it, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
            data = []
            for r in rows[2:]:
                try:
                    data.append((r[isrc], int(r[ie]), float(r[it]), int(r[ism])))
                except (ValueError, IndexError):
                    pass
            tot, ts = sum(x[1] for x in data), max(sum(x[3] for x in data), 1)
            lines.append(f"  SASS: {len(data)} instructions, {tot / 1e6:.1f} M warp-instructions executed, {ts} stall samples")
            lines.append("  blocks of equal execution count with >= 1.5 % of the issued instructions or of the samples:")
            blk, cur = [], None
            for k, (s, e, t, m) in enumerate(data):
                if cur is None or abs(e - cur["e"]) > 0.02 * max(e, cur["e"], 1):
                    if cur:
                        blk.append(cur)
                    cur = {"start": k, "e": e, "n": 0, "t": 0.0, "m": 0, "first": s.strip()}
                cur["n"] += 1
                cur["t"] += t
                cur["m"] += m
            blk.append(cur)
            for b in blk:
                si, ss = b["e"] * b["n"] / tot * 100, b["m"] / ts * 100
                if si >= 1.5 or ss >= 1.5:
                    lines.append(f"    @{b['start']:5d} {b['n']:4d} instr x {b['e'] / 1e6:8.2f} M  lanes {b['t'] / b['n']:4.1f}  issued {si:5.1f} %  samples {ss:5.1f} %   {b['first'][:48]}")
            top = sorted(((m, k, s) for k, (s, e, t, m) in enumerate(data)), reverse=True)[:6]
            lines.append("  most-sampled instructions:")
            for m, k, s in top:
                lines.append(f"    @{k:5d} {m / ts * 100:5.1f} %  {s.strip()[:80]}")
        except ValueError:
            pass
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
