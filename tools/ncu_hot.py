"""Top stall-sampled instructions of one kernel launch in an ncu report (needs the source page):
    python tools/ncu_hot.py report.ncu-rep <kernel regex> [launch index] [rows]"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ia, ie, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[rows.index(hdr) + 1:]:
    try:
        data.append((int(r[ie]), int(r[iS]), r[ia].strip(), [int(r[i]) if r[i] else 0 for i in stall]))
    except (ValueError, IndexError):
        pass
tot, ts = sum(d[0] for d in data), max(1, sum(d[1] for d in data))
print(rows[0][1][:100] if len(rows[0]) > 1 else "", "\ninstr", tot, "samples", ts, "static", len(data))
st = collections.Counter()
for _, _, _, sl in data:
    for i, v in zip(stall, sl):
        st[hdr[i]] += v
print(" ".join("%s %.1f%%" % (k[6:], 100 * v / ts) for k, v in st.most_common(8)))
for n, (c, s, src, sl) in enumerate(data):
    data[n] = (c, s, src, sl, n)
for c, s, src, sl, n in sorted(data, key=lambda d: -d[1])[:top]:
    why = max(zip(sl, [hdr[i][6:] for i in stall]))[1]
    print("%5d %9d %6d (%4.1f%%) %-64s %s" % (n, c, s, 100 * s / ts, src[:64], why))
