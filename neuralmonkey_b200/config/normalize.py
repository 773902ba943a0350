"""Normalisation of the built [main] namespace (behaviour of neuralmonkey/config/normalize.py)."""
import re
import time
from argparse import Namespace
from datetime import timedelta
from typing import Callable, List, Union

import numpy as np

from neuralmonkey_b200.logging import warn
from neuralmonkey_b200.tf_manager import get_default_tf_manager
from neuralmonkey_b200.trainers.delayed_update_trainer import DelayedUpdateTrainer


def normalize_configuration(cfg: Namespace, train_mode: bool) -> None:
    if train_mode:
        _normalize_train_cfg(cfg)
    if cfg.tf_manager is None:
        cfg.tf_manager = get_default_tf_manager()
    cfg.evaluation = [(e[0], e[0], e[1]) if len(e) == 2 else e for e in (cfg.evaluation or [])]
    if cfg.evaluation:
        cfg.main_metric = "{}/{}".format(cfg.evaluation[-1][0], cfg.evaluation[-1][-1].name)
    else:
        cfg.main_metric = "{}/{}".format(cfg.runners[-1].decoder_data_id,
                                         cfg.runners[-1].loss_names[0])
        if not cfg.tf_manager.minimize_metric:
            raise ValueError("minimize_metric must be set to True in TensorFlowManager when using "
                             "loss as the main metric")


def _normalize_train_cfg(cfg: Namespace) -> None:
    cfg.val_datasets = cfg.val_dataset if isinstance(cfg.val_dataset, list) else [cfg.val_dataset]
    cfg.trainers = cfg.trainer if isinstance(cfg.trainer, list) else [cfg.trainer]
    delayed = [t for t in cfg.trainers if isinstance(t, DelayedUpdateTrainer)]
    denominator = 1
    if len(cfg.trainers) > 1 and delayed:
        warn("Weird setup: using more trainers and one of them is delayed update trainer. "
             "No-one can vouch for your safety, user!")
        denominator = int(np.lcm.reduce([t.batches_per_update for t in delayed]))
    elif delayed:
        denominator = cfg.trainers[0].batches_per_update
    cfg.log_timer = _resolve_period(cfg.logging_period, denominator)
    cfg.val_timer = _resolve_period(cfg.validation_period, denominator)


def _resolve_period(period: Union[str, int], denominator: int) -> Callable[[int, float], bool]:
    """Batch count (int) or time span ("3h", "5m", "14s") -> predicate(step, last_time)."""
    if isinstance(period, int):
        if period % denominator != 0:
            raise ValueError("When using delayed update trainer, the logging/validation periods "
                             "must be divisible by batches_per_update.")
        return lambda step, _last: step != 0 and step % period == 0
    parts = re.match(r"((?P<days>\d+?)d)?((?P<hours>\d+?)h)?((?P<minutes>\d+?)m)?((?P<seconds>\d+?)s)?",
                     period)
    if not parts:
        raise ValueError("Validation or logging period have incorrect format. It should be in "
                         "format: 3h; 5m; 14s")
    params = {name: int(val) for name, val in parts.groupdict().items() if val}
    delta = timedelta(**params).total_seconds()
    if delta <= 0:
        raise ValueError("Validation or logging period must be bigger than 0")
    return lambda step, last: step % denominator == 0 and last + delta < time.process_time()
