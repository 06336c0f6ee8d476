"""Key metrics + warp-stall breakdown of a .ncu-rep (run where ncu is installed)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_st.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "lts__t_requests_srcunit_tex_op_write.sum"]
for k in keys:
    for h in hdr:
        if h == k:
            print("%-70s %s %s" % (h, d[h], u.get(h, "")))
stalls = [(h, float(d[h].replace(",", ""))) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")] or \
         [(h, float(d[h].replace(",", ""))) for h in hdr if "warp_issue_stalled" in h and h.endswith(".pct")]
for h, v in sorted(stalls, key=lambda x: -x[1])[:10]:
    print("  stall %-66s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", ""), v))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h, body = rows[1], rows[2:]
    isrc, ismp = h.index("Source"), h.index("# Samples")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    body = [r for r in body if len(r) > ismp and r[ismp].isdigit()]
    tot = sum(int(r[ismp]) for r in body) or 1
    print("  top SASS instructions by warp-stall samples:")
    for r in sorted(body, key=lambda r: -int(r[ismp]))[:int(sys.argv[2])]:
        st = sorted([(int(r[i]), h[i]) for i in stall_cols if r[i].isdigit()], reverse=True)[:2]
        print("  %5.1f%%  %-58s %s" % (100.0 * int(r[ismp]) / tot, r[isrc].strip()[:58], st))
