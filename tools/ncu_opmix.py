"""Opcode mix and per-tile static instruction count of one kernel from an ncu report's source page:
    python tools/ncu_opmix.py report.ncu-rep [dump.txt]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ia, ie, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= ie:
        continue
    try:
        data.append((int(r[ie]), int(r[iS]), r[ia].strip()))
    except ValueError:
        pass
tot = sum(c for c, _, _ in data); ts = max(1, sum(s for _, s, _ in data))
ops, samp = collections.Counter(), collections.Counter()
for c, s, src in data:
    t = src.split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += c; samp[op] += s
print(rows[0][1][:90] if len(rows[0]) > 1 else "", "\ntotal warp instructions %d, static %d" % (tot, len(data)))
for op, c in ops.most_common(24):
    print("%-10s %6.2f%% instr  %6.2f%% samples" % (op, 100 * c / tot, 100 * samp[op] / ts))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        for c, s, src in data:
            f.write("%9d %6d  %s\n" % (c, s, src))
