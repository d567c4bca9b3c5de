"""Turn an ncu report / launch-list CSV from gpurun_out/ into the markdown summaries committed under profiles/.

  python tools/summarize_ncu.py full   gpurun_out/prof_r1k.ncu-rep  profiles/r1k_ncu_full_summary.md "<title>"
  python tools/summarize_ncu.py launch gpurun_out/launches_r1k.csv  profiles/r1k_launches.md         "<title>"
"""
import collections, csv, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size']


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').replace('pile::', '') for r in data]
    lines = ["# " + title, "# raw report %s is scratch (not committed)" % rep, "",
             "| metric | " + " | ".join(names) + " |", "|---|" + "---:|" * len(names)]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append("| %s [%s] | " % (w, units[i]) + " | ".join(r[i][:12] for r in data) + " |")
    open(out, "w").write("\n".join(lines) + "\n")


def launch(path, out, title):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "").replace("pile::", "")})[r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["ns"] for a in agg.values())
    lines = ["# " + title, "# per-launch times are cold-cache / serialised under ncu: compare SHARES", "",
             "| kernel | launches | mean us | share | DRAM read MB/launch | DRAM write MB/launch |", "|---|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        lines.append("| %s | %d | %.1f | %.1f%% | %.0f | %.0f |" % (k, a["n"], a["ns"] / a["n"] / 1e3, 100 * a["ns"] / tot,
                                                                   a["rd"] / a["n"] / 1e6, a["wr"] / a["n"] / 1e6))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
