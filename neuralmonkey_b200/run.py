"""neuralmonkey-run entry point (behaviour of neuralmonkey/run.py:14-92)."""
import argparse
import json
import os

from neuralmonkey_b200.config.configuration import Configuration
from neuralmonkey_b200.experiment import Experiment
from neuralmonkey_b200.logging import log


def load_runtime_config(config_path: str) -> Configuration:
    """The second, small INI: `test_datasets` and optionally `variables`."""
    cfg = Configuration()
    cfg.add_argument("test_datasets")
    cfg.add_argument("variables", required=False, default=None)
    cfg.load_file(config_path)
    cfg.build_model()
    return cfg


def main() -> None:
    parser = argparse.ArgumentParser(description="Runs a model on the given datasets.")
    parser.add_argument("config", metavar="INI-FILE", help="the configuration file of the experiment")
    parser.add_argument("datasets", metavar="INI-TEST-DATASETS",
                        help="the configuration of the test datasets")
    parser.add_argument("--json", type=str, help="write the evaluation results to this file")
    parser.add_argument("-g", "--grid", dest="grid", action="store_true",
                        help="look at the SGE variables for slicing the data")
    args = parser.parse_args()
    exp = Experiment(config_path=args.config)
    exp.build_model()
    datasets_model = load_runtime_config(args.datasets)
    exp.load_variables(datasets_model.model.variables)
    test_datasets = datasets_model.model.test_datasets
    if args.grid and len(test_datasets) > 1:
        raise ValueError("Only one test dataset supported when using --grid")
    results = []
    for dataset in test_datasets:
        if args.grid:
            if "SGE_TASK_FIRST" not in os.environ or "SGE_TASK_LAST" not in os.environ \
                    or "SGE_TASK_STEPSIZE" not in os.environ or "SGE_TASK_ID" not in os.environ:
                raise EnvironmentError("Some SGE environment variables are missing")
            length = int(os.environ["SGE_TASK_STEPSIZE"])
            start = int(os.environ["SGE_TASK_ID"]) - 1
            end = int(os.environ["SGE_TASK_LAST"]) - 1
            if start + length > end:
                length = end - start + 1
            log("Running grid task {} starting at {} with step {}".format(
                start // length, start, length))
            dataset = dataset.subset(start, length)
        if exp.config.args.evaluation is None:
            exp.run_model(dataset, write_out=True, batch_size=exp.config.args.batch_size)
        else:
            results.append(exp.evaluate(dataset, write_out=True, batch_size=exp.config.args.batch_size))
    if args.json:
        with open(args.json, "w") as f_out:
            json.dump(results, f_out)
            f_out.write("\n")


if __name__ == "__main__":
    main()
