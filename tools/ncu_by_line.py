"""Warp-stall samples and executed instructions of an .ncu-rep aggregated by CUDA source line
(needs -lineinfo and --import-source on).  Usage: python tools/ncu_by_line.py rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
isamp, iexec = h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_")]
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= isamp or not r[isamp].isdigit():
        continue
    key = (r[0], r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0, collections.Counter()])
    a[0] += int(r[isamp]); a[1] += int(r[iexec]) if r[iexec].isdigit() else 0
    for i in stall_cols:
        if r[i].isdigit() and int(r[i]):
            a[2][h[i]] += int(r[i])
tot = sum(a[0] for a in agg.values()) or 1
tote = sum(a[1] for a in agg.values()) or 1
print("samples %d, warp instructions %d" % (tot, tote))
for (ln, src), (s, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  L%-4s %-90s %s" % (100.0 * s / tot, 100.0 * e / tote, ln, src,
                                                        ",".join("%s:%d" % (k[6:], v) for k, v in st.most_common(2))))
