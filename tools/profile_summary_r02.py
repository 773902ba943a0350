"""Summarise gpurun_out/ncu_r02/ into profiles/ (tracked): launch shares, per-kernel roofline table with tensor-pipe
utilisation and DRAM traffic, and profiles/r02_traffic.json (the `traffic` figures bench.py reports).

    python tools/profile_summary_r02.py"""
import collections
import csv
import io
import json
import os
import re
import subprocess

SRC = "gpurun_out/ncu_r02"
TAG = "r02"
PEAK_TF, PEAK_HBM = 1507.1, 6505.0
try:
    _p = json.load(open("MEASURED_PEAKS.json"))
    PEAK_TF, PEAK_HBM = float(_p["bf16_tflops_sustained"]), float(_p["hbm_gbs"])
except Exception:  # pylint: disable=broad-except
    pass


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def to_us(val, unit):
    return val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1.0)


def to_bytes(val, unit):
    return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


# ---- launch list ---------------------------------------------------------------------------------
out = []
path = os.path.join(SRC, "launches.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], to_us(fnum(r["Metric Value"]), r["Metric Unit"])))
    agg = collections.OrderedDict()
    for name, us in rows:
        a = agg.setdefault(name.split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    out += ["# ncu launch list ({}): `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py "
            "--steps 2 --warmup 1 --no-cpu-baseline --no-extras`".format(TAG), "",
            "{} launches captured (model build, warm-up, 2 timed steps, profile steps), {:.2f} ms of kernel time; per-launch "
            "times are cold-cache and serialised, so read the SHARE.".format(len(rows), total / 1e3), "",
            "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        out.append("| `{}` | {} | {:.3f} | {:.1f}% | {:.1f} |".format(name[:110], n, us / 1e3, 100 * us / total, us / n))
    open("profiles/{}_launches.md".format(TAG), "w").write("\n".join(out) + "\n")

# ---- launch list of the Transformer workload (newest call directory that has one) ---------------------------
import glob
cands = sorted(glob.glob("gpurun_out/*/transformer_launches.csv"), key=os.path.getmtime)
if cands:
    lines = [l for l in open(cands[-1]) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], to_us(fnum(r["Metric Value"]), r["Metric Unit"])))
    agg = collections.OrderedDict()
    for name, us in rows:
        a = agg.setdefault(re.sub(r"\(.*", "", name)[:110], [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    env = ""
    try:
        env = open(os.path.join(os.path.dirname(cands[-1]), "transformer_env.txt")).read().strip()
    except OSError:
        pass
    tdoc = ["# ncu launch list ({}): Transformer-base training steps (`bench_workloads.py transformer`) {}".format(TAG, env), "",
            "{} launches captured (model build, warm-up and timed steps; the capture stops at its launch cap), {:.1f} ms of "
            "kernel time; per-launch times are cold-cache and serialised, so read the SHARE.".format(len(rows), total / 1e3), "",
            "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
        tdoc.append("| `{}` | {} | {:.3f} | {:.1f}% | {:.1f} |".format(name, n, us / 1e3, 100 * us / total, us / n))
    open("profiles/{}_transformer_launches.md".format(TAG), "w").write("\n".join(tdoc) + "\n")

# ---- full captures ------------------------------------------------------------------------------------
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic"]
doc = ["# ncu --set full captures ({}): `--clock-control none --import-source on`".format(TAG), "",
       "Peaks (MEASURED_PEAKS.json): HBM {:.0f} GB/s, dense bf16 sustained {:.1f} TFLOP/s.".format(PEAK_HBM, PEAK_TF), ""]
traffic = {}
for rep, title in (("prof_tc_gemm", "tcgen05 GEMM, the kind::f16 instances (vocabulary projection and its gradients)"),
                   ("prof_tc_tf32", "tcgen05 GEMM, kind::tf32 instances of one training step (dense projections, weight gradients)"),
                   ("prof_rnn", "GRU recurrences and Bahdanau attention (training)"),
                   ("prof_decode", "decoding: fused step, vocabulary GEMM combine, beam top-k")):
    path, csv_path = os.path.join(SRC, rep + ".ncu-rep"), os.path.join(SRC, rep + ".raw.csv")
    if os.path.exists(csv_path):        # exported on the GPU box (the reports themselves exceed gpurun's 64 MiB)
        txt = open(csv_path).read()
    elif os.path.exists(path):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        continue
    rd = list(csv.reader(io.StringIO(txt)))
    header, units, data = rd[0], rd[1], rd[2:]
    class _Idx(dict):          # this ncu prefixes some metrics with their unit ("FBSP.TriageCompute.dram__throughput...")
        def __missing__(self, key):
            for h, i in self.items():
                if h.endswith("." + key):
                    return i
            for h, i in self.items():
                if key in h:
                    return i
            raise KeyError(key)
    idx = _Idx({h: i for i, h in enumerate(header)})
    tens = [h for h in header if ("umma" in h or "tensor" in h) and "pct" in h]
    doc += ["## {} (`{}`)".format(title, rep), "",
            "| kernel | us | DRAM rd MB | DRAM wr MB | DRAM % of measured peak | SM % | tensor pipe % | grid x block | regs |",
            "|---|---|---|---|---|---|---|---|---|"]
    groups = collections.OrderedDict()
    for row in data:
        name = row[idx["Kernel Name"]]
        us = to_us(fnum(row[idx["gpu__time_duration.sum"]]), units[idx["gpu__time_duration.sum"]])
        rd_b = to_bytes(fnum(row[idx["dram__bytes_read.sum"]]), units[idx["dram__bytes_read.sum"]])
        wr_b = to_bytes(fnum(row[idx["dram__bytes_write.sum"]]), units[idx["dram__bytes_write.sum"]])
        tp = max([fnum(row[idx[h]]) for h in tens if row[idx[h]] not in ("", "n/a")] or [float("nan")])
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"\(int\)|\(bool\)", "", short)[:90]
        doc.append("| `{}` | {:.1f} | {:.1f} | {:.1f} | {} | {} | {:.1f} | {} x {} | {} |".format(
            short, us, rd_b / 1e6, wr_b / 1e6, "{:.1f}".format(100.0 * (rd_b + wr_b) / max(us, 1e-9) / 1e3 / PEAK_HBM),
            "{:.1f}".format(fnum(row[idx["sm__throughput.avg.pct_of_peak_sustained_elapsed"]])), tp, row[idx["launch__grid_size"]],
            row[idx["launch__block_size"]], row[idx["launch__registers_per_thread"]]))
        g = groups.setdefault(short, [0, 0.0, 0.0, 0.0])
        g[0] += 1; g[1] += us; g[2] += rd_b + wr_b; g[3] = max(g[3], tp if tp == tp else 0.0)
        # map the fp16 vocabulary instances to the entry points bench.py names
        m = re.search(r"tc_gemm_kernel<(\d+), (\w+), (\w+), (\d+), (\d+)(?:, \d+)?>", short)
        if m:
            bn, a_mn, b_mn, mode, esz = m.groups()
            call = None
            if esz == "2" and mode == "1":
                call = "nm_logits_xent_fwd16"
            elif esz == "2" and mode == "3":
                call = "nm_logits_xent_bwd16"
            elif esz == "2" and mode == "0" and a_mn in ("true", "1"):
                call = "nm_gemm_f16_tn"
            elif esz == "2" and mode == "0":
                call = "nm_gemm_f16"
            if call and us > 100.0:
                traffic[call] = rd_b + wr_b
    doc += ["", "Per kernel (sum over its captured launches):", "", "| kernel | launches | total us | DRAM MB | GB/s | max tensor % |",
            "|---|---|---|---|---|---|"]
    for short, (n, us, byt, tp) in groups.items():
        doc.append("| `{}` | {} | {:.1f} | {:.1f} | {:.0f} | {:.1f} |".format(short, n, us, byt / 1e6, byt / max(us, 1e-9) / 1e3, tp))
    doc.append("")
open("profiles/{}_ncu_full.md".format(TAG), "w").write("\n".join(doc) + "\n")
if traffic:
    json.dump(traffic, open("profiles/{}_traffic.json".format(TAG), "w"), indent=1)
print("\n".join(out[:36]))
print(json.dumps(traffic))
