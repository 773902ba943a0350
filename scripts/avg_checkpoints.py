#!/usr/bin/env python3
"""Average the variables of several checkpoints written by `neuralmonkey-train` of this package
(same command line as the reference's scripts/avg_checkpoints.py: CHECKPOINT... OUTPUT_PATH).

A checkpoint here is a `torch.save`d dictionary {"variables": {TF-style name: tensor}, ...}; the output
holds the arithmetic mean of every variable and no optimizer moments, under OUTPUT_PATH (the reference's
saver appends "-0": `experiment.load_variables` looks for `variables.data.avg-0`, then
`variables.data.avg`, before falling back to the best checkpoint)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), ".."))
from neuralmonkey_b200.logging import log  # noqa: E402


def average(paths):
    missing = [p for p in paths if not os.path.exists(p)]
    if missing:
        raise ValueError("Provided checkpoints do not exist: {}".format(", ".join(missing)))
    total = None
    for path in paths:
        log("Reading from checkpoint {}".format(path), color="blue")
        variables = torch.load(path, map_location="cpu")["variables"]
        if total is None:
            total = {name: value.to(torch.float64).clone() for name, value in variables.items()}
            continue
        if set(variables) != set(total):
            raise ValueError("Checkpoint {} holds different variables than {}".format(path, paths[0]))
        for name, value in variables.items():
            total[name] += value.to(torch.float64)
    return {name: (value / len(paths)).to(torch.float32) for name, value in total.items()}


def main() -> None:
    parser = argparse.ArgumentParser(description=__doc__)
    parser.add_argument("checkpoints", type=str, nargs="+",
                        help="Space-separated list of checkpoints to average.")
    parser.add_argument("output_path", type=str, help="Path to output the averaged checkpoint to.")
    args = parser.parse_args()
    torch.save({"variables": average(args.checkpoints)}, args.output_path)
    log("Averaged checkpoints saved in {}".format(args.output_path), color="blue")


if __name__ == "__main__":
    main()
