#!/usr/bin/env python3
"""Benchmark of the hot path BASELINE.json names: training throughput of the Bahdanau en-de
configuration (examples/translation.ini dims: E = He = H = O = 300, C = A = 600, Tx = Ty = 50,
batch 256 per GPU, synthetic V = 32000) in non-pad target tokens per second over full
optimizer steps (forward + backward + [all-reduce] + clip + Adam).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  `value` is device-resident throughput (ids already in
HBM); `e2e` feeds every step from pinned host memory and reads the loss back.  The
`roofline` object describes the dominant kernel (timed with CUDA events inside the timed
steps), `cpu_baseline` times the CPU oracle restatement of the same step on a bounded sample.
`--impl reference` runs ONLY that CPU restatement (the reference's TF-1.12 path cannot run
here: SURVEY.md 8(c)).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train_target_tokens_per_sec"
UNIT = "tokens/s"
DIMS = dict(vs=32000, vt=32000, es=300, he=300, et=300, hd=300, out=300, maxout=False, max_len=50,
            supress_unk=False)
BATCH, TX, TY = 256, 50, 50
CPU_SAMPLE_SENTENCES = 16


# DRAM bytes per launch of the fused vocabulary kernels at the bench shape, from the committed
# ncu --set full capture (profiles/r01_ncu_full.md): read + write
XENT_TRAFFIC = {"nm_logits_xent_fwd": 0.054071e9 + 0.017578e9, "nm_logits_xent_bwd": 0.072671e9 + 1.584649e9}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true",
                    help="issue every kernel of a step from Python instead of replaying the captured step")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary workloads (transformer / beam-8 / captioning) at N=1")
    ap.add_argument("--breakdown", action="store_true", help="print the per-entry-point time table")
    return ap.parse_args()


def workload_config(n_gpus, batch):
    return {"workload": "examples/translation.ini GRU+Bahdanau en-de, synthetic ids",
            "per_gpu_batch": batch, "global_batch": batch * n_gpus, "src_len": TX, "tgt_len": TY,
            "vocab": DIMS["vt"], "emb": 300, "rnn": 300, "optimizer": "Adam 1e-4, clip 1.0 per tensor, l2 1e-8",
            "lengths": "fixed (no padding)", "parallelism": "dp{}".format(n_gpus),
            "step_submission": "CrossEntropyTrainer(use_cuda_graph=True): the step is captured once per batch shape and replayed (N>1: backward graph, NCCL all-reduce, clip+Adam graph)",
            "gemm": "tcgen05 kind::tf32 (fp32 storage, fp32 accumulate); GRU recurrences on tcgen05 with weights resident in tensor memory (fp16 operands forward, tf32 backward, fp32 accumulate)",
            "l2_between_iters": "working set per step (>1.6 GB dlogits) exceeds the 126 MB L2"}


def synthetic_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(4, DIMS["vs"], (batch, TX), generator=g)
    tgt = torch.randint(4, DIMS["vt"], (batch, TY), generator=g)
    tgt[:, TY - 1] = 2  # </s>
    return src, tgt


# ---------------------------------------------------------------------------
# CPU arm: the oracle restatement, timed
# ---------------------------------------------------------------------------
def cpu_train_tokens_per_sec(n_sent, steps, warmup, seed=2574600):
    from oracle import nm_oracle as O
    p = O.init_bahdanau_params(DIMS["vs"], DIMS["vt"], DIMS["es"], DIMS["he"], DIMS["et"], DIMS["hd"],
                               None, DIMS["out"], False, seed=seed)
    spec = O.RNNDecoderSpec("decoder", "attention", DIMS["max_len"], "tanh", False)
    st = O.AdamState(p)
    src, tgt = synthetic_batch(n_sent, seed)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(p, spec, "sentence_encoder", src, tgt.t(), st, l2=1e-8, clip_norm=1.0)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return n_sent * TY / per_step, per_step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    cores = torch.get_num_threads()
    value, per_step = cpu_train_tokens_per_sec(CPU_SAMPLE_SENTENCES, steps, warmup)
    sample = ("{} sentences x {} target tokens per step of the same workload, {} timed steps"
              .format(CPU_SAMPLE_SENTENCES, TY, steps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, args.batch),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample,
                             "note": "restated-reference CPU baseline (TF 1.12 unavailable on this box)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:  # pylint: disable=broad-except
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for s in self.samples:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), s[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])),
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args):
    from neuralmonkey_b200 import distributed, lib
    from tests.helpers import build_bahdanau, feed
    distributed.init_from_env()
    rank, world = distributed.rank(), distributed.world_size()
    model = build_bahdanau(**DIMS, clip=1.0, l2=1e-8, lr=1e-4, cuda_graph=not args.no_cuda_graph)
    trainer = model["trainer"]
    dev = model["arena"].params.device
    batch = args.batch
    tokens_per_step_rank = batch * TY

    # a few distinct synthetic batches, resident on the device (value) and pinned on host (e2e)
    host_batches = [synthetic_batch(batch, 2574600 + rank + 1000 * i) for i in range(4)]
    pinned = [(s.pin_memory(), t.pin_memory()) for s, t in host_batches]

    def step_resident(i):
        src, tgt = host_batches[i % len(host_batches)]
        feed(model, src, tgt, train=True)  # ids tiny; see e2e for the counted copy
        return trainer.train_step()

    dev_src = [s.to(dev) for s, _ in host_batches]

    def step_device_inputs(i):
        j = i % len(host_batches)
        enc, att, dec = model["enc"], model["att"], model["dec"]
        enc.input_sequence.feed_ids([dev_src[j]], train=True)
        for part in (enc, att):
            part.reset_batch()
            part.train_mode = True
            part.batch_size = batch
        dec.feed_ids(host_batches[j][1], batch, train=True)
        return trainer.train_step()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(step_fn, steps, read_loss):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        last = None
        t_host = time.perf_counter()
        for i in range(steps):
            out = step_fn(i)
            if read_loss:
                last = float(out["losses"][0])  # device -> host read of the step's loss
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / steps   # time the host needs to ENQUEUE a step
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / steps, last

    for i in range(args.warmup):
        step_device_inputs(i)
    with ClockSampler(dev.index or 0) as clocks:
        ms_step, _ = timed(step_device_inputs, args.steps, read_loss=False)
    host_enqueue_ms = host_ms[0]

    # per-entry-point device time (CUDA events around every C-ABI call) and the kernel count of a
    # step: these steps issue every launch from Python - a replayed graph contains the same kernel
    # nodes but does not pass through the library's launch counter
    graphed = trainer.use_cuda_graph
    trainer.use_cuda_graph = False
    launches0 = lib.launch_count()
    lib.profile_start()
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        step_device_inputs(i)
    prof = lib.profile_stop()
    launches = (lib.launch_count() - launches0) // prof_steps * args.steps
    trainer.use_cuda_graph = graphed

    def step_e2e(i):
        src, tgt = pinned[i % len(pinned)]
        feed(model, src, tgt, train=True)  # H2D of the pinned id tensors inside the timed region
        return trainer.train_step()

    for i in range(2):
        step_e2e(i)
    ms_e2e, loss = timed(step_e2e, args.steps, read_loss=True)

    if rank != 0:
        return
    value = world * tokens_per_step_rank / (ms_step * 1e-3)
    e2e_value = world * tokens_per_step_rank / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel -------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # pylint: disable=broad-except
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md sustained)"
    m, k, v = batch * TY, DIMS["out"], DIMS["vt"]
    table = sorted(((n, d["ms"] / prof_steps, d["calls"] // prof_steps) for n, d in prof.items()),
                   key=lambda x: -x[1])
    total_ms = sum(t for _, t, _ in table)
    # The dominant kernel is tc_gemm_kernel (gemm_tc.cu): every dense projection, weight-gradient
    # product and the fused vocabulary forward/backward are instances of it.  Algorithmic flops per
    # launch = 2*M*N*K of the product; time = CUDA events around the launches
    # inside the profiled steps.
    import re
    fam_flops, fam_ms, instances = 0.0, 0.0, []
    for name, d in prof.items():
        fl = None
        mm = re.match(r"nm_gemm\[(\w\w) (\d+)x(\d+)x(\d+)\]", name)
        if mm:
            fl = 2.0 * int(mm.group(2)) * int(mm.group(3)) * int(mm.group(4))
        elif name in ("nm_logits_xent_fwd", "nm_logits_xent_bwd", "nm_logits_xent_fwd16", "nm_logits_xent_bwd16"):
            fl = 2.0 * m * k * v
        else:
            mm = re.match(r"nm_gemm_f16\[(\d+)x(\d+)x(\d+)\]", name)     # NMB200_XENT16=1: kind::f16 instances
            if mm:
                fl = 2.0 * int(mm.group(1)) * int(mm.group(2)) * int(mm.group(3))
        if fl is None:
            continue
        calls = d["calls"]
        fam_flops += fl * calls
        fam_ms += d["ms"]
        instances.append({"call": name, "launches_per_step": calls // prof_steps,
                          "ms_per_launch": d["ms"] / calls, "tflops": fl / (d["ms"] / calls * 1e-3) / 1e12,
                          "traffic": XENT_TRAFFIC.get(name),
                          "kind": "f16" if ("16" in name.split("[")[0]) else "tf32"})
    instances.sort(key=lambda e: -e["ms_per_launch"] * e["launches_per_step"])
    roof = None
    if fam_ms > 0:
        ach = fam_flops / (fam_ms * 1e-3) / 1e12
        best = max(instances, key=lambda e: e["tflops"])
        roof = {"kernel": "tc_gemm_kernel (tcgen05 kind::tf32, all instances of the step)", "bound": "tensor",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": XENT_TRAFFIC["nm_logits_xent_fwd"] + XENT_TRAFFIC["nm_logits_xent_bwd"],
                "traffic_note": "DRAM read+write bytes per launch of the two fused vocabulary instances "
                                "(profiles/r01_ncu_full.md, ncu --set full); their algorithmic bytes are "
                                "{:.0f} MB (fwd) and {:.0f} MB (bwd, incl. the fp32 dlogits it must write)"
                                .format((4.0 * (m * k + k * v + v) + 32.0 * m * ((v + 255) // 256)) / 1e6,
                                        4.0 * (m * k + k * v + v + m * v) / 1e6),
                "peak_source": peak_src,
                "note": "kind::tf32 issues at half the bf16 rate, so 0.5 is the ceiling of frac for this "
                        "kernel; frac of the tf32 ceiling = {:.3f} (best instance {}: {:.0f} TFLOP/s = {:.3f})"
                        .format(ach / (peak_tf / 2.0), best["call"], best["tflops"],
                                best["tflops"] / (peak_tf / 2.0)),
                "share_of_step": fam_ms / prof_steps / max(total_ms, 1e-9),
                "instances": instances[:6]}
    if args.breakdown:
        for n, t, c in table:
            print("# {:34s} {:8.3f} ms/step  {:4d} calls/step".format(n, t, c), file=sys.stderr)
        print("# sum {:.3f} ms of device time vs {:.3f} ms per step".format(total_ms, ms_step),
              file=sys.stderr)

    cpu = None
    if not args.no_cpu_baseline:
        cores = torch.get_num_threads()
        cpu_value, cpu_step = cpu_train_tokens_per_sec(CPU_SAMPLE_SENTENCES, 2, 1)
        cpu = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "{} sentences x {} target tokens per step, 2 timed steps ({:.1f} s each)"
                         .format(CPU_SAMPLE_SENTENCES, TY, cpu_step)}

    # secondary workloads of BASELINE.json (configs[2..4]); N=1 only, a few seconds each
    extras = None
    if world == 1 and not args.no_extras:
        import bench_workloads
        extras = {}
        for name in ("transformer", "beam", "captioning"):
            try:
                extras[name] = bench_workloads.RUNNERS[name]()
            except Exception as exc:  # pylint: disable=broad-except
                extras[name] = {"error": "{}: {}".format(type(exc).__name__, exc)}

    h2d = 2 * batch * (TX + TY) * 8  # int64 ids: encoder ids, decoder targets + fed symbols
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (tf32 tensor-core products, fp32 accumulate)",
            "data": "synthetic", "config": workload_config(world, batch),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "last_loss": loss},
            "gpu_launches": launches,
            "gpu_launches_note": "kernels of libnmb200 per step x timed steps, counted by nm_launch_count() over "
                                 "eagerly issued steps" + (" (the timed steps replay the same kernels as one "
                                                           "captured CUDA graph per step)" if graphed else ""),
            "clocks": clocks.summary(), "roofline": roof,
            "cpu_baseline": cpu,
            "breakdown_ms_per_step": {n: round(t, 4) for n, t, _ in table[:16]},
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "extra_workloads": extras}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries (NCCL prints its version banner) write to file descriptor 1: keep the real stdout
    # for the JSON line only and send everything else to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: neuralmonkey_b200 has no CPU path "
                         "(use --impl reference for the CPU restatement)")
    run_b200(args)


if __name__ == "__main__":
    main()
