"""Training / inference loops (behaviour of neuralmonkey/learning_utils.py:38-530).

The reference's TensorBoard summaries and CPU-time profiler are dropped; throughput is
reported in wall-clock target tokens per second."""
import time
from argparse import Namespace
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from neuralmonkey_b200 import distributed
from neuralmonkey_b200.dataset import Dataset
from neuralmonkey_b200.logging import log, log_print, warn
from neuralmonkey_b200.runners.base_runner import BaseRunner, ExecutionResult
from neuralmonkey_b200.runners.dataset_runner import DatasetRunner
from neuralmonkey_b200.tf_manager import TensorFlowManager


def shard_batch(batch: Dataset) -> Dataset:
    """This rank's sentences of a batch (data-parallel training)."""
    world = distributed.world_size()
    if world == 1:
        return batch
    rows = {s: list(batch.get_series(s)) for s in batch.series}
    n = len(batch)
    b = distributed.shard_bounds(n, world)
    lo, hi = b[distributed.rank()], b[distributed.rank() + 1]
    data = {s: (lambda v=vals[lo:hi]: iter(v)) for s, vals in rows.items()}
    return Dataset("{}.rank{}".format(batch.name, distributed.rank()), data, batch.batching)


def merge_batches(first: Dataset, second: Dataset) -> Dataset:
    """The sentences of two batches as one batch (same series)."""
    rows = {s: list(first.get_series(s)) + list(second.get_series(s)) for s in second.series}
    data = {s: (lambda v=vals: iter(v)) for s, vals in rows.items()}
    return Dataset(second.name, data, second.batching)


def training_loop(cfg: Namespace) -> None:
    _check_series_collisions(cfg.runners, cfg.postprocess)
    if cfg.initial_variables is not None:
        cfg.tf_manager.restore(cfg.initial_variables)
    cfg.tf_manager.initialize_model_parts(cfg.runners + cfg.trainers)
    feedables = set.union(*[ex.feedables for ex in cfg.runners + cfg.trainers])
    log("Starting training")
    step = 0
    seen_instances = 0
    last_log = last_val = time.process_time()
    t_start = time.perf_counter()
    interrupt = None
    try:
        for epoch_n in range(1, cfg.epochs + 1):
            log_print("")
            log("Epoch {} begins".format(epoch_n), color="red")
            batches = cfg.train_dataset.batches()
            if epoch_n == 1 and cfg.train_start_offset:
                _skip_lines(cfg.train_start_offset, batches)
            carry = None        # data parallel: a batch with fewer sentences than ranks is held over
            for batch_n, batch in enumerate(batches):
                if distributed.world_size() > 1:
                    # every rank needs a non-empty shard (the kernels require B > 0, and a rank that skips its
                    # step would leave the others waiting in the all-reduce).  All ranks see the same global
                    # batches, so they take this decision identically: the short batch (the remainder of an
                    # epoch, a flushed bucket) joins the next one; what is left at the end of the epoch
                    # (< world sentences) is dropped.
                    if carry is not None:
                        batch, carry = merge_batches(carry, batch), None
                    if len(batch) < distributed.world_size():
                        carry = batch
                        continue
                step += 1
                seen_instances += len(batch)
                local = shard_batch(batch)
                if cfg.log_timer(step, last_log):
                    trainer_result = cfg.tf_manager.execute(local, feedables, cfg.trainers, train=True)
                    results, outputs, f_batch = run_on_dataset(
                        cfg.tf_manager, cfg.runners, cfg.dataset_runner, batch, cfg.postprocess,
                        write_out=False)
                    evaluation_result = evaluation(cfg.evaluation, f_batch, results, outputs)
                    _log_evaluation(cfg.main_metric, evaluation_result, seen_instances, epoch_n,
                                    cfg.epochs, trainer_result, train=True,
                                    rate=seen_instances / (time.perf_counter() - t_start))
                    last_log = time.process_time()
                else:
                    cfg.tf_manager.execute(local, feedables, cfg.trainers, train=True, summaries=False)
                if cfg.val_timer(step, last_val) and cfg.val_datasets:
                    log_print("")
                    for val_id, valset in enumerate(cfg.val_datasets):
                        results, outputs, f_val = run_on_dataset(
                            cfg.tf_manager, cfg.runners, cfg.dataset_runner, valset, cfg.postprocess,
                            write_out=False)
                        val_eval = evaluation(cfg.evaluation, f_val, results, outputs)
                        header = "Validation (epoch {}, batch number {}):".format(epoch_n, batch_n)
                        log(header, color="blue")
                        _print_examples(f_val, outputs, cfg.val_preview_input_series,
                                        cfg.val_preview_output_series, cfg.val_preview_num_examples)
                        log_print("")
                        log(header, color="blue")       # the reference repeats the header below the preview
                        if val_id == len(cfg.val_datasets) - 1:
                            score = val_eval[cfg.main_metric]
                            if distributed.rank() == 0:
                                cfg.tf_manager.validation_hook(score, epoch_n, batch_n)
                                if score == cfg.tf_manager.best_score:
                                    # a new best: the parts with a `save_checkpoint` file store their
                                    # variables (learning_utils.py:146-159)
                                    cfg.tf_manager.initialize_model_parts(cfg.runners + cfg.trainers,
                                                                          save=True)
                            if hasattr(cfg.tf_manager, "sync_validation_state"):
                                cfg.tf_manager.sync_validation_state()
                            log("best {} on validation: {:.4g} (in epoch {}, after batch number {})"
                                .format(cfg.main_metric, cfg.tf_manager.best_score,
                                        cfg.tf_manager.best_score_epoch,
                                        cfg.tf_manager.best_score_batch), color="blue")
                        name = "val_{}".format(val_id) if len(cfg.val_datasets) > 1 else None
                        _log_evaluation(cfg.main_metric, val_eval, seen_instances, epoch_n, cfg.epochs,
                                        results, train=False, dataset_name=name)
                    last_val = time.process_time()
                    log_print("")
    except KeyboardInterrupt as ex:
        interrupt = ex
    log("Training finished. Maximum {} on validation data: {:.4g}, epoch {}".format(
        cfg.main_metric, cfg.tf_manager.best_score, cfg.tf_manager.best_score_epoch))
    if interrupt is not None:
        raise interrupt  # pylint: disable=raising-bad-type


def _skip_lines(start_offset: int, batches) -> None:
    skipped = 0
    while skipped < start_offset:
        try:
            skipped += len(next(batches))
        except StopIteration:
            raise ValueError("Trying to skip more instances than the size of the dataset")


def _check_series_collisions(runners: List[BaseRunner], postprocess) -> None:
    used = set()
    for runner in runners:
        if runner.output_series in used:
            raise Exception("Output series '{}' is multiple times among the runners' outputs."
                            .format(runner.output_series))
        used.add(runner.output_series)
    for series, _ in postprocess or []:
        if series in used:
            raise Exception("Postprocess output series '{}' already exists.".format(series))
        used.add(series)


def run_on_dataset(tf_manager: TensorFlowManager, runners: List[BaseRunner],
                   dataset_runner: DatasetRunner, dataset: Dataset, postprocess,
                   write_out: bool = False, log_progress: int = 0
                   ) -> Tuple[List[ExecutionResult], Dict[str, List], Dict[str, List]]:
    """Apply the model to a dataset in batches (learning_utils.py:272-393)."""
    contains_targets = all(r.decoder_data_id in dataset for r in runners
                           if r.decoder_data_id is not None)
    last_log_time = time.process_time()
    batch_results = [[] for _ in range(len(runners) + 1)]  # type: List[List[ExecutionResult]]
    feedables = set.union(*[runner.feedables for runner in runners]) | dataset_runner.feedables
    fetched_input = {s: [] for s in dataset.series}  # type: Dict[str, List]
    processed = 0
    for batch in dataset.batches():
        if 0 < log_progress < time.process_time() - last_log_time:
            log("Processed {} examples.".format(processed))
            last_log_time = time.process_time()
        results = tf_manager.execute(batch, feedables, list(runners) + [dataset_runner],
                                     compute_losses=contains_targets)
        processed += len(batch)
        for lst, res in zip(batch_results, results):
            lst.append(res)
        for s_id in batch.series:
            fetched_input[s_id].extend(batch.get_series(s_id))
    all_results = [join_execution_results(res) for res in batch_results[:-1]]
    lengths = {s: len(fetched_input[s]) for s in dataset.series}
    if len(set(lengths.values())) != 1:
        warn("Fetched input dataset series are not of the same length: {}".format(lengths))
    dataset_len = lengths[dataset.series[0]]
    result_data = {}  # type: Dict[str, Any]
    for res in all_results:
        for s_id, data in res.outputs.items():
            if s_id in result_data:
                raise ValueError("Overwriting output series forbidden.")
            result_data[s_id] = data
    if postprocess is not None:
        for series_name, postprocessor in postprocess:
            post = postprocessor(fetched_input, result_data)
            result_data[series_name] = post if hasattr(post, "__len__") else list(post)
    for series_id, data in result_data.items():
        if len(data) != dataset_len:
            warn("Output '{}' for dataset '{}' has length {}, but input dataset size is {}"
                 .format(series_id, dataset.name, len(data), dataset_len))
    if write_out and dataset.outputs is not None:
        for series_id, data in result_data.items():
            if series_id in dataset.outputs:
                path, writer = dataset.outputs[series_id]
                writer(path, data)
            else:
                log("There is no file for output series '{}' in dataset: '{}'"
                    .format(series_id, dataset.name), color="red")
    elif write_out:
        log("Dataset does not have any outputs, nothing to write out.", color="red")
    return all_results, result_data, fetched_input


def join_execution_results(results: List[ExecutionResult]) -> ExecutionResult:
    losses_sum = {loss: 0. for loss in results[0].losses}
    outputs = {}  # type: Dict[str, Any]
    for key in results[0].outputs:
        joined = []  # type: List[Any]
        for res in results:
            joined.extend(res.outputs[key])
        outputs[key] = np.array(joined) if joined and isinstance(joined[0], np.ndarray) else joined
    for res in results:
        for l_id, loss in res.losses.items():
            losses_sum[l_id] += loss * res.size
    total = sum(res.size for res in results)
    losses = {l_id: loss / total for l_id, loss in losses_sum.items()}
    return ExecutionResult(outputs, losses, total, [])


def evaluation(evaluators, batch, execution_results, result_data) -> Dict[str, float]:
    """Losses of the runs plus `series/metric` values (learning_utils.py:434-467)."""
    eval_result = {}  # type: Dict[str, float]
    for result in execution_results:
        if any(l in eval_result for l in result.losses):
            raise ValueError("Duplicate loss result keys found.")
        eval_result.update(result.losses)
    for hypothesis_id, reference_id, function in evaluators:
        if reference_id not in batch or hypothesis_id not in result_data:
            continue
        eval_result["{}/{}".format(hypothesis_id, function.name)] = function(
            result_data[hypothesis_id], batch[reference_id])
    return eval_result


def _format_evaluation_line(evaluation_res: Dict[str, float], main_metric: str) -> str:
    """`name: value` pairs in the order they were computed (runner losses, then the evaluators), four
    spaces apart, the main metric moved to the end (learning_utils.py:504-516)."""
    others = ["{}: {:.4g}".format(name, value) for name, value in evaluation_res.items() if name != main_metric]
    return "    ".join(others) + "    {}: {:.4g}".format(main_metric, evaluation_res[main_metric])


def _log_evaluation(main_metric: str, eval_result: Dict[str, float], seen_instances: int, epoch: int,
                    max_epochs: int, execution_results, train: bool = False,
                    dataset_name: str = None, rate: float = None) -> None:
    """One line per logged batch / validation, as _log_continuous_evaluation writes it (:470-502): yellow for
    training batches, blue for validation.  (TensorBoard summaries are out of scope.)"""
    if distributed.rank() != 0:
        return
    log("Epoch {}/{}  Instances {}  {}".format(epoch, max_epochs, seen_instances,
                                              _format_evaluation_line(eval_result, main_metric)),
        color="yellow" if train else "blue")


def print_final_evaluation(eval_result: Dict[str, float], name: str = None) -> None:
    """The aligned table printed after a test set was evaluated (learning_utils.py:519-530)."""
    if name is not None:
        log("Model evaluated on '{}'".format(name))
    for eval_name, value in eval_result.items():
        log("... {}:{} {:.4g}".format(eval_name, " " * (22 - len(eval_name)), value))
    log_print("")


def _data_item_to_str(item: Any) -> str:
    if isinstance(item, list):
        return " ".join(_data_item_to_str(i) for i in item)
    if isinstance(item, dict):
        rows = ["{}: {}".format(_data_item_to_str(k), _data_item_to_str(v)) for k, v in item.items()]
        return "{\n      " + "\n      ".join(rows) + "\n    }"
    if isinstance(item, np.ndarray) and item.ndim > 1:
        return "[numpy tensor, shape {}]".format(item.shape)
    return str(item)


def _print_examples(dataset: Dict[str, List[Any]], outputs: Dict[str, List[Any]],
                    val_preview_input_series: Optional[List[str]] = None,
                    val_preview_output_series: Optional[List[str]] = None,
                    num_examples: int = 15) -> None:
    """The validation preview (learning_utils.py:548-613): per example its number, the source series
    (in the dataset, not produced), the produced series, and `<series> (ref)` for produced series that
    the dataset also holds - each group sorted by name."""
    log_print("Examples:")
    assert outputs
    sources = [s for s in dataset if s not in outputs]
    targets = [s for s in dataset if s in outputs]
    produced = list(outputs.keys())
    if val_preview_input_series is not None:
        sources = [s for s in sources if s in val_preview_input_series]
        targets = [s for s in targets if s in val_preview_input_series]
    if val_preview_output_series is not None:
        produced = [s for s in produced if s in val_preview_output_series]
    columns = {s: list(dataset[s]) for s in sources + targets}
    results = {s: list(outputs[s]) for s in produced}
    for i in range(min(len(next(iter(dataset.values()))), num_examples)):
        log_print("  [{}]".format(i + 1))
        for series in sorted(sources):
            log_print("  {}: {}".format(series, _data_item_to_str(columns[series][i])))
        for series in sorted(produced):
            log_print("  {}: {}".format(series, _data_item_to_str(results[series][i])))
        for series in sorted(targets):
            log_print("  {} (ref): {}".format(series, _data_item_to_str(columns[series][i])))
        log_print("")
