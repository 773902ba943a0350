"""neuralmonkey-train entry point (behaviour of neuralmonkey/train.py:19-74)."""
import argparse
import os
import shlex
import sys
import traceback
from shutil import copyfile

from neuralmonkey_b200.experiment import Experiment
from neuralmonkey_b200.logging import debug, log


def _main() -> None:
    parser = argparse.ArgumentParser(description="Trains a model given by a configuration file.")
    parser.add_argument("config", metavar="INI-FILE", help="the configuration file for the experiment")
    parser.add_argument("-s", "--set", type=str, metavar="SETTING", action="append", dest="config_changes",
                        default=[], help="override an option in the configuration; the syntax is "
                        "[section.]option=value")
    parser.add_argument("-v", "--var", type=str, metavar="VAR", default=[], action="append",
                        dest="config_vars", help="set a variable in the configuration; the syntax is "
                        "var=value (shorthand for -s vars.var=value)")
    parser.add_argument("-i", "--init", dest="init_only", action="store_true",
                        help="initialize the experiment directory and exit without building the model")
    parser.add_argument("-f", "--overwrite", dest="overwrite", action="store_true",
                        help="force overwriting the output directory; can be used to start an "
                        "experiment with configuration files left behind by a previous ``--init``")
    args = parser.parse_args()
    args.config_changes.extend("vars.{}".format(s) for s in args.config_vars)
    exp = Experiment(config_path=args.config, config_changes=args.config_changes, train_mode=True,
                     overwrite_output_dir=args.overwrite)
    with open(exp.get_path("args", exp.cont_index + 1), "w", encoding="utf-8") as file:
        print(" ".join(shlex.quote(a) for a in sys.argv), file=file)
    if args.init_only:
        if exp.cont_index >= 0:
            log("The experiment directory already exists.", color="red")
            exit(2)
        exp.config.save_file(exp.get_path("experiment.ini", 0))
        copyfile(args.config, exp.get_path("original.ini", 0))
        log("Experiment directory initialized.")
        cmd = [os.path.basename(sys.argv[0]), "-f", exp.get_path("experiment.ini", 0)]
        log("To start experiment, run: {}".format(" ".join(shlex.quote(a) for a in cmd)))
        exit(0)
    try:
        exp.train()
    except KeyboardInterrupt:  # pylint: disable=try-except-raise
        raise
    except Exception:  # pylint: disable=broad-except
        log(traceback.format_exc(), color="red")
        exit(1)


def main() -> None:
    try:
        _main()
    except KeyboardInterrupt:
        log("Training interrupted by user.")
        debug(traceback.format_exc())
        exit(1)


if __name__ == "__main__":
    main()
