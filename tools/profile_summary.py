"""Summarise gpurun_out/ ncu artefacts into profiles/ (tracked).  Usage: python tools/profile_summary.py r01"""
import csv, io, subprocess, sys, collections, json, os

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []

# ---- launch list --------------------------------------------------------------------
rows = []
with open("gpurun_out/launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(io.StringIO("".join(lines)))
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
agg = collections.OrderedDict()
for name, ns in rows:
    short = name.split("(")[0]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(v[1] for v in agg.values())
out.append("# ncu launch list ({}): `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 2 --warmup 1`".format(tag))
out.append("")
out.append("{} launches captured (warm-up + 2 steps + setup), {:.2f} ms of kernel time; per-launch times are cold-cache and serialised, so read the SHARE.".format(len(rows), total / 1e6))
out.append("")
out.append("| kernel | launches | total ms | share | avg us |")
out.append("|---|---|---|---|---|")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    out.append("| `{}` | {} | {:.3f} | {:.1f}% | {:.1f} |".format(name[:110], n, ns / 1e6, 100 * ns / total, ns / n / 1e3))
open("profiles/{}_launches.md".format(tag), "w").write("\n".join(out) + "\n")

# ---- full captures -----------------------------------------------------------------------
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__cluster_size",
           "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "launch__shared_mem_per_block_dynamic", "dram__bytes.sum.per_second"]
summary = {}
doc = ["# ncu --set full captures ({})".format(tag), ""]
for rep in ("prof_tc_gemm", "prof_gru", "prof_att"):
    path = "gpurun_out/{}.ncu-rep".format(rep)
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    header, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(header)}
    doc.append("## {}".format(rep))
    doc.append("")
    cols = [m for m in METRICS if m in idx]
    for row in data:
        name = row[idx["Kernel Name"]].split("(")[0][:100]
        doc.append("### `{}`".format(name))
        doc.append("")
        doc.append("| metric | value | unit |")
        doc.append("|---|---|---|")
        rec = {}
        for m in cols:
            doc.append("| {} | {} | {} |".format(m, row[idx[m]], units[idx[m]]))
            rec[m] = (row[idx[m]], units[idx[m]])
        doc.append("")
        summary.setdefault(rep, []).append((name, rec))
    # any metric mentioning umma / tensor
    tens = [h for h in header if "tensor" in h and "pct" in h][:12]
    if tens and data:
        doc.append("tensor-pipe metrics of the first launch: " + ", ".join("{}={}".format(h, data[0][idx[h]]) for h in tens))
        doc.append("")
open("profiles/{}_ncu_full.md".format(tag), "w").write("\n".join(doc) + "\n")
print("\n".join(out[:40]))
