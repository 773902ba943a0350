"""neuralmonkey-run entry point (behaviour of neuralmonkey/run.py:14-92): load a trained experiment,
read a second INI naming the test datasets (and optionally the variable files), decode / evaluate each
dataset, optionally dump the evaluation results as JSON."""
import argparse
import json
import os
from typing import Any, Dict, List, Optional

from neuralmonkey_b200.config.configuration import Configuration
from neuralmonkey_b200.dataset import Dataset
from neuralmonkey_b200.experiment import Experiment
from neuralmonkey_b200.logging import log

_SGE_VARIABLES = ("SGE_TASK_FIRST", "SGE_TASK_LAST", "SGE_TASK_STEPSIZE", "SGE_TASK_ID")


def load_runtime_config(config_path: str) -> Configuration:
    """The second, small INI: `test_datasets` and optionally `variables`."""
    cfg = Configuration()
    cfg.add_argument("test_datasets")
    cfg.add_argument("variables", required=False, default=None, cond=lambda x: x is None or isinstance(x, list))
    cfg.load_file(config_path)
    cfg.build_model()
    return cfg


def _cli() -> argparse.ArgumentParser:
    cli = argparse.ArgumentParser(description="Runs a model on the given datasets.")
    cli.add_argument("config", metavar="INI-FILE", help="the configuration file of the experiment")
    cli.add_argument("datasets", metavar="INI-TEST-DATASETS", help="the configuration of the test datasets")
    cli.add_argument("-s", "--set", type=str, metavar="SETTING", action="append", dest="config_changes", default=[],
                     help="override an option in the configuration; the syntax is [section.]option=value")
    cli.add_argument("-v", "--var", type=str, metavar="VAR", default=[], action="append", dest="config_vars",
                     help="set a variable in the configuration; the syntax is var=value (shorthand for "
                          "-s vars.var=value)")
    cli.add_argument("--json", type=str, help="write the evaluation results to this file in JSON format")
    cli.add_argument("-g", "--grid", dest="grid", action="store_true",
                     help="look at the SGE variables for slicing the data")
    return cli


def _grid_slice(dataset: Dataset) -> Dataset:
    """The part of the dataset this Sun Grid Engine array task is responsible for."""
    if any(name not in os.environ for name in _SGE_VARIABLES):
        raise EnvironmentError("Some SGE environment variables are missing")
    step = int(os.environ["SGE_TASK_STEPSIZE"])
    first = int(os.environ["SGE_TASK_ID"]) - 1
    last = int(os.environ["SGE_TASK_LAST"]) - 1
    length = min(step, last - first + 1) if first + step > last else step
    log("Running grid task {} starting at {} with step {}".format(first // length, first, length))
    return dataset.subset(first, length)


def main() -> None:
    args = _cli().parse_args()
    runtime_cfg = load_runtime_config(args.datasets)       # first, as the reference does: a bad dataset INI fails early
    changes = list(args.config_changes) + ["vars.{}".format(v) for v in args.config_vars]
    exp = Experiment(config_path=args.config, config_changes=changes)
    exp.build_model()
    exp.load_variables(runtime_cfg.model.variables)
    datasets = runtime_cfg.model.test_datasets
    if args.grid and len(datasets) > 1:
        raise ValueError("Only one test dataset supported when using --grid")
    batch_size = exp.config.args.batch_size
    results = []  # type: List[Dict[str, Any]]
    for dataset in datasets:
        if args.grid:
            dataset = _grid_slice(dataset)
        if exp.config.args.evaluation is None:
            exp.run_model(dataset, write_out=True, batch_size=batch_size)
        else:
            results.append(exp.evaluate(dataset, write_out=True, batch_size=batch_size))
    _dump(results, args.json)


def _dump(results: List[Dict[str, Any]], path: Optional[str]) -> None:
    if path:
        with open(path, "w") as handle:
            json.dump(results, handle)
            handle.write("\n")


if __name__ == "__main__":
    main()
