"""Top stalled SASS instructions of one kernel in an ncu report (source page).  python tools/ncu_hot.py rep regex [N]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
lines = out.splitlines()
# first kernel only
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(lines[start:end]))
hdr = rows[0]
S = hdr.index("# Samples"); SRC = hdr.index("Source"); EX = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in rows[1:] if r[S].isdigit())
print("total samples", tot, "instructions", len(rows) - 1)
idx = sorted(range(1, len(rows)), key=lambda i: -int(rows[i][S] or 0))[:n]
for i in sorted(idx):
    r = rows[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%% ex=%8s  %-60s %s" % (i, 100.0 * int(r[S]) / tot, r[EX], r[SRC].strip()[:60], " ".join("%s=%d" % (b, a) for a, b in st if a)))
# coarse histogram over the instruction stream (bins of 40 SASS instructions) with the dominant opcode families
print("--- sample histogram by SASS position")
bins = {}
for i in range(1, len(rows)):
    s = int(rows[i][S] or 0)
    bins.setdefault(i // 40, [0, []])
    bins[i // 40][0] += s
    bins[i // 40][1].append(rows[i][SRC].strip().split()[0 if not rows[i][SRC].strip().startswith("@") else 1])
for b in sorted(bins):
    if bins[b][0] * 100 >= tot:
        ops = {}
        for o in bins[b][1]:
            ops[o.split(".")[0]] = ops.get(o.split(".")[0], 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
        print("%5d-%5d %5.1f%%  %s" % (b * 40, b * 40 + 39, 100.0 * bins[b][0] / tot, " ".join("%s:%d" % kv for kv in top)))
