"""Execution manager (API of neuralmonkey/tf_manager.py:29-306).

The class keeps its historical name because every Neural Monkey INI instantiates
`tf_manager.TensorFlowManager`; there is no TensorFlow behind it.  `execute()` feeds one batch
to the feedables and runs the executables of the given trainers / runners on the CUDA
kernels; the n-best checkpoint bookkeeping (`variables.data[.i]`, `variables.data.best`) is
kept; checkpoints are `torch.save`d {TF-style variable name: tensor} dictionaries.

"Sessions" (`num_sessions > 1`, tf_manager.py:78-79,179-180) are model ENSEMBLES in the reference: N
checkpoints of the same graph, run one after the other on every batch.  Here a session is one flat
copy of the parameter buffer; activating it copies it into the arena (every variable is a view of that
buffer) and swaps in that session's cache of per-batch tensors, so the model parts compute with those
parameters.  Only runners support several sessions, as in the reference.
"""
import os
from typing import List, Optional, Set, Union

import numpy as np
import torch

from neuralmonkey_b200 import runtime
from neuralmonkey_b200.logging import log
from neuralmonkey_b200.model.feedable import Feedable
from neuralmonkey_b200.runners.base_runner import ExecutionResult, GraphExecutor


class TensorFlowManager:
    # pylint: disable=too-many-arguments
    def __init__(self, num_sessions: int, num_threads: int, save_n_best: int = 1,
                 minimize_metric: bool = False, gpu_allow_growth: bool = True,
                 per_process_gpu_memory_fraction: float = 1.0,
                 enable_tf_debug: bool = False) -> None:
        if num_sessions < 1:
            raise ValueError("num_sessions must be positive")
        self.num_sessions = num_sessions
        self._session_params = None   # type: Optional[List[torch.Tensor]]
        self._session_caches = []     # type: List[dict]
        self.num_threads = num_threads  # host threads only matter for the CPU reference
        if save_n_best < 1:
            raise Exception("save_n_best parameter must be greater than zero")
        self.saver_max_to_keep = save_n_best
        self.minimize_metric = minimize_metric
        self.best_score_index = 0
        self.best_score_epoch = 0
        self.best_score_batch = 0
        init_score = np.inf if self.minimize_metric else -np.inf
        self.saved_scores = [init_score for _ in range(self.saver_max_to_keep)]
        self.best_vars_file = None
        self.variables_files = []  # type: List[str]
        self.sessions = [self] * num_sessions  # kept for code that iterates over sessions

    # -- sessions (ensembles) ------------------------------------------------------------------------
    def _session_buffers(self) -> List[torch.Tensor]:
        if self._session_params is None:
            params = runtime.arena().params
            self._session_params = [params.detach().clone() for _ in range(self.num_sessions)]
        return self._session_params

    def activate_session(self, index: int, parts=()) -> None:
        """Make session `index` the one the model parts compute with: its parameters into the arena,
        its per-batch tensor caches into the parts."""
        if self.num_sessions == 1:
            return
        with torch.no_grad():
            runtime.arena().params.copy_(self._session_buffers()[index])
        caches = self._session_caches[index]
        for part in parts:
            part.__dict__["_batch_cache"] = caches.setdefault(id(part), {})

    @property
    def best_score(self) -> float:
        return self.saved_scores[self.best_score_index]

    def _is_better(self, score1: float, score2: float) -> bool:
        return score1 < score2 if self.minimize_metric else score1 > score2

    def _argworst(self, scores: List[float]) -> int:
        return int(np.argmax(scores)) if self.minimize_metric else int(np.argmin(scores))

    def _update_best_vars(self, var_index: int) -> None:
        best_vars_prefix = os.path.basename(self.variables_files[var_index])
        with open(self.best_vars_file, "w", encoding="utf-8") as var_file:
            var_file.write(best_vars_prefix)

    def init_saving(self, vars_prefix: str) -> None:
        if self.saver_max_to_keep == 1:
            self.variables_files = [vars_prefix]
        else:
            self.variables_files = ["{}.{}".format(vars_prefix, i)
                                    for i in range(self.saver_max_to_keep)]
        self.best_vars_file = "{}.best".format(vars_prefix)
        self._update_best_vars(var_index=0)

    def validation_hook(self, score: float, epoch: int, batch: int) -> None:
        """Keep the n best checkpoints (tf_manager.py:133-155)."""
        if self._is_better(score, self.best_score):
            self.best_score_epoch = epoch
            self.best_score_batch = batch
        worst_index = self._argworst(self.saved_scores)
        worst_score = self.saved_scores[worst_index]
        if self._is_better(score, worst_score):
            worst_var_file = self.variables_files[worst_index]
            self.save(worst_var_file)
            self.saved_scores[worst_index] = score
            log("Variable file saved in {}".format(worst_var_file))
            if self._is_better(score, self.best_score):
                self.best_score_index = worst_index
                self._update_best_vars(self.best_score_index)
                log("Best scores saved so far: {}".format(self.saved_scores))
        log("Best scores saved so far: {}".format(self.saved_scores))

    # -- execution ---------------------------------------------------------------------------
    # pylint: disable=too-many-arguments
    def execute(self, batch, feedables: Set[Feedable], runners: List[GraphExecutor],
                train: bool = False, compute_losses: bool = True,
                summaries: bool = True) -> List[ExecutionResult]:
        """Feed `batch` and run every executor on it (tf_manager.py:188-225)."""
        for feedable in feedables:
            feedable.feed_dict(batch, train)
        executables = [runner.get_executable(compute_losses=compute_losses, summaries=summaries,
                                             num_sessions=self.num_sessions) for runner in runners]
        if self.num_sessions > 1:
            # every executable sees the batch once per session and combines what the sessions produced
            # (base_runner.py collect_results protocol); trainers refuse several sessions themselves
            parts = set(feedables)
            for runner in runners:
                parts |= set(getattr(runner, "parameterizeds", ())) | set(getattr(runner, "feedables", ()))
                parts.add(getattr(runner, "decoder", runner))
            self._session_caches = [dict() for _ in range(self.num_sessions)]
            for executable in executables:
                executable.execute_sessions(lambda i: self.activate_session(i, parts), self.num_sessions)
            return [executable.result for executable in executables]
        for i, executable in enumerate(executables):
            if train and i > 0:
                # several trainers on one batch (tests/bahdanau.ini:12): the reference runs their
                # train ops in one sess.run; here each trainer gets a fresh forward pass (its
                # backward consumed the previous graph), i.e. the updates are applied in order
                for feedable in feedables:
                    feedable.feed_dict(batch, train)
            executable.execute()
        return [executable.result for executable in executables]

    # -- checkpoints ---------------------------------------------------------------------------
    def save(self, variable_files: Union[str, List[str]]) -> None:
        if isinstance(variable_files, str):
            variable_files = [variable_files]
        if len(variable_files) != self.num_sessions:
            raise Exception("Provided {} files for restoring {} sessions.".format(
                len(variable_files), self.num_sessions))
        arena = runtime.arena()
        for index, path in enumerate(variable_files):
            self.activate_session(index)
            state = {"variables": arena.state_dict(), "adam_m": arena.moment_dict(arena.adam_m),
                     "adam_v": arena.moment_dict(arena.adam_v),
                     "global_step": runtime.global_step()}    # the Saver stores it too
            # per-optimizer state (the Saver stores every optimizer's slot variables and beta powers): the update
            # counts, and the moments of the optimizers beyond the first one
            slots = getattr(arena, "optimizer_slots", [])
            state["optimizer_steps"] = [int(opt.steps) for opt, _m, _v in slots]
            for i, (_opt, m, v) in enumerate(slots[1:], start=1):
                state["adam_m_{}".format(i)] = arena.moment_dict(m)
                state["adam_v_{}".format(i)] = arena.moment_dict(v)
            torch.save(state, path)

    def restore(self, variable_files: Union[str, List[str]]) -> None:
        if isinstance(variable_files, str):
            variable_files = [variable_files]
        if len(variable_files) != self.num_sessions:
            raise Exception("Provided {} files for restoring {} sessions.".format(
                len(variable_files), self.num_sessions))
        arena = runtime.arena()
        for index, path in enumerate(variable_files):
            log("Loading variables from {}".format(path))
            ckpt = torch.load(path, map_location="cpu")
            arena.load_dict(ckpt["variables"])
            if self.num_sessions > 1:                   # one flat copy of the parameters per session
                self._session_buffers()[index].copy_(arena.params.detach())
            elif isinstance(ckpt.get("adam_m"), dict):   # optimizer moments, keyed by variable name
                arena.load_moments(arena.adam_m, ckpt["adam_m"])
                arena.load_moments(arena.adam_v, ckpt["adam_v"])
                # warm moments need the steps they belong to: lr schedules (global step) and every optimizer's
                # bias correction (its own update count; checkpoints without it: one optimizer, = global step)
                gstep = int(ckpt.get("global_step", 0))
                runtime.set_global_step(gstep)
                steps = list(ckpt.get("optimizer_steps") or [gstep])
                slots = arena.optimizer_slots
                arena.restored_optimizer_state = {}
                for i, count in enumerate(steps):
                    m_i = ckpt.get("adam_m_{}".format(i)) if i else ckpt["adam_m"]
                    v_i = ckpt.get("adam_v_{}".format(i)) if i else ckpt["adam_v"]
                    if i < len(slots):                  # the optimizer is already at work: apply now
                        opt, m, v = slots[i]
                        if i and isinstance(m_i, dict):
                            arena.load_moments(m, m_i)
                            arena.load_moments(v, v_i)
                        opt.steps = int(count)
                    else:                               # applied when the optimizer claims its slot
                        arena.restored_optimizer_state[i] = {"m": m_i or {}, "v": v_i or {}, "steps": int(count)}

    def sync_validation_state(self) -> None:
        """Data parallel: rank 0 alone runs validation_hook (it owns the checkpoint files); the other ranks
        take over its bookkeeping, so that `best_score*` in their logs and - what matters -
        restore_best_vars() on them pick the same checkpoint."""
        from neuralmonkey_b200 import distributed
        if distributed.world_size() <= 1:
            return
        import torch.distributed as dist
        state = [None]
        if distributed.rank() == 0:
            state = [(self.best_score_index, self.best_score_epoch, getattr(self, "best_score_batch", 0),
                      list(self.saved_scores))]
        dist.broadcast_object_list(state, src=0)
        index, epoch, batch, scores = state[0]
        self.best_score_index, self.best_score_epoch, self.best_score_batch = index, epoch, batch
        self.saved_scores = list(scores)

    def restore_best_vars(self) -> None:
        self.restore(self.variables_files[self.best_score_index])

    def initialize_sessions(self) -> None:
        """Variables were initialised when the arena was finalised; nothing to run."""

    def initialize_model_parts(self, runners, save: bool = False) -> None:
        """Initialize model parts variables from their checkpoints (tf_manager.py:279-289); with `save`,
        write the parts' `save_checkpoint` files instead (what learning_utils.py:150-159 does on a new
        best validation score)."""
        if any(not hasattr(r, "parameterizeds") for r in runners):
            raise TypeError("Args to initialize_model_parts must be trainers or runners")
        parameterizeds = set.union(*[rnr.parameterizeds for rnr in runners]) if runners else set()
        for coder in sorted(parameterizeds, key=lambda c: getattr(c, "name", "")):
            if save:
                coder.save()
            else:
                coder.load()


def get_default_tf_manager() -> TensorFlowManager:
    return TensorFlowManager(num_sessions=1, num_threads=4)
