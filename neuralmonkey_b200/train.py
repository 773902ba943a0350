"""`neuralmonkey-train <experiment.ini>`: command line of the training entry point
(options and exit codes as neuralmonkey/train.py; under torchrun every rank runs this with
its own GPU and the trainer exchanges gradients over NCCL)."""
import argparse
import os
import shlex
import sys
import traceback
from shutil import copyfile

from neuralmonkey_b200.experiment import Experiment
from neuralmonkey_b200.logging import debug, log


def _parser() -> argparse.ArgumentParser:
    cli = argparse.ArgumentParser(description="Trains a model given by a configuration file.")
    cli.add_argument("config", metavar="INI-FILE", help="the configuration file for the experiment")
    cli.add_argument("-s", "--set", dest="config_changes", metavar="SETTING", type=str, action="append",
                     default=[], help="override an option in the configuration; the syntax is "
                                      "[section.]option=value")
    cli.add_argument("-v", "--var", dest="config_vars", metavar="VAR", type=str, action="append",
                     default=[], help="set a variable in the configuration; the syntax is var=value "
                                      "(shorthand for -s vars.var=value)")
    cli.add_argument("-i", "--init", dest="init_only", action="store_true",
                     help="initialize the experiment directory and exit without building the model")
    cli.add_argument("-f", "--overwrite", dest="overwrite", action="store_true",
                     help="force overwriting the output directory; can be used to start an experiment "
                          "with configuration files left behind by a previous ``--init``")
    return cli


def _quoted(argv) -> str:
    return " ".join(shlex.quote(arg) for arg in argv)


def _initialize_directory(exp: Experiment, config_path: str) -> int:
    """`--init`: write experiment.ini / original.ini and tell the user how to continue."""
    if exp.cont_index >= 0:
        log("The experiment directory already exists.", color="red")
        return 2
    target = exp.get_path("experiment.ini", 0)
    exp.config.save_file(target)
    copyfile(config_path, exp.get_path("original.ini", 0))
    log("Experiment directory initialized.")
    log("To start experiment, run: {}".format(_quoted([os.path.basename(sys.argv[0]), "-f", target])))
    return 0


def _run(options: argparse.Namespace) -> int:
    changes = list(options.config_changes) + ["vars.{}".format(v) for v in options.config_vars]
    exp = Experiment(config_path=options.config, config_changes=changes, train_mode=True,
                     overwrite_output_dir=options.overwrite)
    with open(exp.get_path("args", exp.cont_index + 1), "w", encoding="utf-8") as record:
        record.write(_quoted(sys.argv) + "\n")
    if options.init_only:
        return _initialize_directory(exp, options.config)
    try:
        exp.train()
    except KeyboardInterrupt:
        raise
    except Exception:  # pylint: disable=broad-except
        log(traceback.format_exc(), color="red")
        return 1
    return 0


def main() -> None:
    try:
        code = _run(_parser().parse_args())
    except KeyboardInterrupt:
        log("Training interrupted by user.")
        debug(traceback.format_exc())
        code = 1
    if code:
        sys.exit(code)


if __name__ == "__main__":
    main()
