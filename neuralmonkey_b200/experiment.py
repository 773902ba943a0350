"""Experiment: configuration -> built model -> train / run (API of neuralmonkey/experiment.py)."""
import os
import random
import subprocess
from argparse import Namespace
from shutil import copyfile
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from neuralmonkey_b200 import distributed, runtime
from neuralmonkey_b200.config.configuration import Configuration
from neuralmonkey_b200.config.normalize import normalize_configuration
from neuralmonkey_b200.dataset import Dataset
from neuralmonkey_b200.learning_utils import (evaluation, print_final_evaluation, run_on_dataset,
                                              training_loop)
from neuralmonkey_b200.logging import Logging, log
from neuralmonkey_b200.runners.dataset_runner import DatasetRunner

# the reference's list (experiment.py:28-34) plus train_start_offset, which the reference forgets: there
# `neuralmonkey-run` rejects an INI that sets it ("Unexpected fields")
_TRAIN_ARGS = ["val_dataset", "trainer", "name", "train_dataset", "epochs", "test_datasets",
               "initial_variables", "validation_period", "val_preview_input_series",
               "val_preview_output_series", "val_preview_num_examples", "logging_period",
               "visualize_embeddings", "overwrite_output_dir", "train_start_offset"]
_EXPERIMENT_FILES = ["experiment.log", "experiment.ini", "original.ini", "git_commit", "git_diff",
                     "variables.data.best"]


class Experiment:
    _current_experiment = None  # type: Optional["Experiment"]

    def __init__(self, config_path: str, train_mode: bool = False,
                 overwrite_output_dir: bool = False, config_changes: List[str] = None) -> None:
        self.train_mode = train_mode
        self._config_path = config_path
        self.cont_index = -1
        self._model_built = False
        self._vars_loaded = False
        self._model = None  # type: Optional[Namespace]
        self.config = create_config(train_mode)
        self.config.load_file(config_path, config_changes)
        args = self.config.args
        if self.train_mode:
            if os.path.isdir(args.output) and os.path.exists(os.path.join(args.output, "experiment.ini")):
                if args.overwrite_output_dir or overwrite_output_dir:
                    log("Directory with experiment.ini '{}' exists, overwriting enabled, proceeding."
                        .format(args.output))
                else:
                    raise RuntimeError("Directory with experiment.ini '{}' exists, overwriting "
                                       "disabled.".format(args.output))
            os.makedirs(args.output, exist_ok=True)
        while any(os.path.exists(self.get_path(f, self.cont_index + 1)) for f in _EXPERIMENT_FILES):
            self.cont_index += 1

    @property
    def model(self) -> Namespace:
        if self._model is None:
            raise RuntimeError("Experiment argument model not initialized")
        return self._model

    def build_model(self) -> None:
        """Instantiate every object of the configuration, allocate the parameter arena."""
        if self._model_built:
            raise RuntimeError("build_model() called twice")
        distributed.init_from_env()
        runtime.reset()
        seed = self.config.args.random_seed
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        type(self)._current_experiment = self
        self.config.build_model(warn_unused=self.train_mode)
        self._model = self.config.model
        self._model_built = True
        normalize_configuration(self._model, self.train_mode)
        if not hasattr(self._model, "dataset_runner") or self._model.dataset_runner is None:
            self._model.dataset_runner = DatasetRunner()
        executors = list(self._model.runners)
        if self.train_mode:
            executors += list(self._model.trainers)
        for executor in executors:
            for part in executor.parameterizeds:
                part.ensure_declared()
        runtime.arena().finalize(runtime.device(), seed=seed)
        n_params = sum(v.numel for v in runtime.arena().variables.values())
        log("Model built: {} variables, {} parameters".format(len(runtime.arena().order), n_params))
        type(self)._current_experiment = None

    def train(self) -> None:
        if not self.train_mode:
            raise RuntimeError("train() was called, but the experiment is not in training mode")
        if not self._model_built:
            self.build_model()
        self.cont_index += 1
        main = distributed.rank() == 0
        if main:
            self.config.save_file(self.get_path("experiment.ini"))
            copyfile(self._config_path, self.get_path("original.ini"))
            save_git_info(self.get_path("git_commit"), self.get_path("git_diff"))
            Logging.set_log_file(self.get_path("experiment.log"))
        self.model.tf_manager.init_saving(self.get_path("variables.data"))
        training_loop(self.model)
        final_variables = self.get_path("variables.data.final")
        log("Saving final variables in {}".format(final_variables))
        if main:
            self.model.tf_manager.save(final_variables)
        distributed.barrier()        # the other ranks must not read checkpoints rank 0 is still writing
        if self.model.test_datasets:
            if os.path.exists(self.get_path("variables.data.best")):
                self.model.tf_manager.restore_best_vars()
            for test_id, dataset in enumerate(self.model.test_datasets):
                # every rank evaluates (the replicas are identical); rank 0 alone writes the output files
                self.evaluate(dataset, write_out=main, name="test_{}".format(test_id))
        log("Finished.")
        self._vars_loaded = True

    def load_variables(self, variable_files: List[str] = None) -> None:
        if not self._model_built:
            self.build_model()
        if variable_files is None:
            if os.path.exists(self.get_path("variables.data.avg-0")):
                variable_files = [self.get_path("variables.data.avg-0")]
            elif os.path.exists(self.get_path("variables.data.avg")):
                variable_files = [self.get_path("variables.data.avg")]
            elif os.path.exists(self.get_path("variables.data.best")):
                with open(self.get_path("variables.data.best")) as f_best:
                    variable_files = [os.path.join(self.config.args.output, f_best.read().rstrip())]
            else:
                variable_files = [self.get_path("variables.data.final")]
            log("Default variable file '{}' will be used for loading variables."
                .format(variable_files[0]))
        self.model.tf_manager.restore(variable_files)
        self._vars_loaded = True

    def run_model(self, dataset: Dataset, write_out: bool = False, batch_size: int = None,
                  log_progress: int = 0):
        if not self._model_built:
            self.build_model()
        if not self._vars_loaded:
            self.load_variables()
        return run_on_dataset(self.model.tf_manager, self.model.runners, self.model.dataset_runner,
                              dataset, self.model.postprocess, write_out=write_out,
                              log_progress=log_progress)

    def evaluate(self, dataset: Dataset, write_out: bool = False, batch_size: int = None,
                 log_progress: int = 0, name: str = None) -> Dict[str, Any]:
        execution_results, output_data, f_dataset = self.run_model(dataset, write_out, batch_size,
                                                                   log_progress)
        eval_result = evaluation(self.model.evaluation, f_dataset, execution_results, output_data)
        if eval_result:
            print_final_evaluation(eval_result, name or dataset.name)
        return eval_result

    def get_path(self, filename: str, cont_index: int = None) -> str:
        if cont_index is None:
            cont_index = self.cont_index
        cont_suffix = ".cont-{}".format(cont_index) if cont_index > 0 else ""
        if filename.startswith("variables.data"):
            new_filename = "variables.data" + cont_suffix + filename[len("variables.data"):]
        else:
            new_filename = filename + cont_suffix
        return os.path.join(self.config.args.output, new_filename)

    @classmethod
    def get_current(cls) -> "Experiment":
        if cls._current_experiment is None:
            raise RuntimeError("No experiment is being built")
        return cls._current_experiment


def create_config(train_mode: bool = True) -> Configuration:
    """The accepted [main] fields (experiment.py:453-491)."""
    config = Configuration()
    config.add_argument("tf_manager", required=False, default=None)
    config.add_argument("batch_size", required=False, default=None, cond=lambda x: x is None or x > 0)
    config.add_argument("output")
    config.add_argument("postprocess", required=False, default=None)
    config.add_argument("runners")
    config.add_argument("random_seed", required=False, default=2574600)
    if train_mode:
        config.add_argument("epochs", cond=lambda x: x >= 0)
        config.add_argument("trainer")
        config.add_argument("train_dataset")
        config.add_argument("val_dataset", required=False, default=[])
        config.add_argument("evaluation")
        config.add_argument("test_datasets", required=False, default=[])
        config.add_argument("logging_period", required=False, default=20)
        config.add_argument("validation_period", required=False, default=500)
        config.add_argument("visualize_embeddings", required=False, default=None)
        config.add_argument("val_preview_input_series", required=False, default=None)
        config.add_argument("val_preview_output_series", required=False, default=None)
        config.add_argument("val_preview_num_examples", required=False, default=15)
        config.add_argument("train_start_offset", required=False, default=0)
        config.add_argument("name", required=False, default="Neural Monkey Experiment")
        config.add_argument("initial_variables", required=False, default=None)
        config.add_argument("overwrite_output_dir", required=False, default=False)
    else:
        config.add_argument("evaluation", required=False, default=None)
        for argument in _TRAIN_ARGS:
            config.ignore_argument(argument)
    return config


def save_git_info(git_commit_file: str, git_diff_file: str, branch: str = "HEAD",
                  repo_dir: str = None) -> None:
    if repo_dir is None:
        repo_dir = os.path.abspath(os.path.join(os.path.dirname(os.path.realpath(__file__)), os.pardir))
    try:
        with open(git_commit_file, "wb") as file:
            subprocess.run(["git", "log", "-1", "--format=%H", branch], cwd=repo_dir, stdout=file,
                           stderr=subprocess.DEVNULL, check=False)
        with open(git_diff_file, "wb") as file:
            subprocess.run(["git", "--no-pager", "diff", "--color=always", branch], cwd=repo_dir,
                           stdout=file, stderr=subprocess.DEVNULL, check=False)
    except OSError:
        pass
