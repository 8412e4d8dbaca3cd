"""Write the text summary of an ncu report that profiles/ keeps (development aid).
usage: python tools/ncu_summary.py report.ncu-rep out.txt "header line" """
import csv, io, subprocess, sys, collections

rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
ki = hdr.index("Kernel Name")
lines = [header, "", "Kernel Name".ljust(86) + " | ".join(r[ki].split("(")[0] for r in data)]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append(f"{w:76s} {units[i]:9s} " + " | ".join(r[i] for r in data))
for r in data:
    name = r[ki].split("(")[0]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + name + "$"], capture_output=True, text=True).stdout
    cur = None; h = None; agg = {}
    for row in csv.reader(io.StringIO(src)):
        if not row: continue
        if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
        if row[0] == "Line No":
            h = row; ia = h.index("Address"); ii = h.index("Instructions Executed"); it = h.index("Thread Instructions Executed")
            isamp = h.index("# Samples"); st = [k for k, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
            continue
        if h is None: continue
        try: ln = int(row[0])
        except ValueError: continue
        if row[ia] != "-": continue
        try: inst = int(row[ii]); thr = int(row[it]); samp = int(row[isamp])
        except ValueError: continue
        a = agg.setdefault((cur, ln), [0, 0, 0, collections.Counter(), row[1]])
        a[0] += inst; a[1] += thr; a[2] += samp
        for k in st:
            try: a[3][h[k]] += int(row[k])
            except ValueError: pass
    ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[2] for a in agg.values()) or 1
    allst = collections.Counter()
    for a in agg.values(): allst.update(a[3])
    lines += ["", f"== {name}: source lines by executed warp instructions (share, active lanes, share of stall samples)",
              "   stall samples: " + ", ".join(f"{k[6:]} {100*v/ts:.1f}%" for k, v in allst.most_common(6))]
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        lines.append(f"   {f}:{l:<5d} {100*a[0]/ti:5.1f}%  lanes {a[1]/max(a[0],1):4.1f}  samples {100*a[2]/ts:4.1f}%  | {a[4].strip()[:100]}")
    lines.append("   -- by stall samples")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:8]:
        top = ", ".join(f"{k[6:]}" for k, v in a[3].most_common(2))
        lines.append(f"   {f}:{l:<5d} samples {100*a[2]/ts:4.1f}%  ({top})  | {a[4].strip()[:100]}")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
