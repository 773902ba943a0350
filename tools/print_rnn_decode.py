"""Print the per-step numbers of a `bench_workloads.py rnn_decode` result file."""
import json
import sys

try:
    d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    for k in ("greedy", "beam8_batch", "beam8_latency"):
        e = d[k]
        print(k, round(e["us_per_step"], 1), "us/step", round(e["tokens_per_s"]), "tok/s", e.get("step_breakdown_us"))
except Exception as exc:  # pylint: disable=broad-except
    print("rnn_decode unreadable", exc)
