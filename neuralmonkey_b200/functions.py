"""Schedules (reference: neuralmonkey/functions.py:11-79).  The reference returns graph
tensors of the global step; here a schedule is a callable step -> value, which the
optimizer shim evaluates on the host each step."""
import math
from typing import Callable, List


def inverse_sigmoid_decay(param, rate, min_value: float = 0., max_value: float = 1.,
                          name=None, dtype=None) -> Callable[[int], float]:
    """rate / (rate + exp(step / rate)) style decay of scheduled sampling (functions.py:11-36)."""
    def value(step: int) -> float:
        x = rate / (rate + math.exp(param(step) / rate if callable(param) else param / rate))
        return x * (max_value - min_value) + min_value
    return value


def piecewise_function(param, values: List[float], changepoints: List[int], name=None,
                       dtype=None) -> Callable[[int], float]:
    """values[i] while changepoints[i-1] <= param < changepoints[i] (functions.py:39-55)."""
    if len(changepoints) != len(values) - 1:
        raise ValueError("changepoints must be one shorter than values")

    def value(step: int) -> float:
        p = param(step) if callable(param) else param
        idx = sum(1 for c in changepoints if p >= c)
        return values[idx]
    return value


def noam_decay(learning_rate: float, model_dimension: int,
               warmup_steps: int) -> Callable[[int], float]:
    """lr * d^-0.5 * min(step^-0.5, step * warmup^-1.5) with step = the global step BEFORE
    the update (functions.py:58-79): the very first update therefore runs with rate 0."""
    def value(step: int) -> float:
        step = float(step)
        inv_sq_dim = 1.0 / math.sqrt(model_dimension)
        inv_sq3_warmup = math.pow(warmup_steps, -1.5)
        inv_sq_step = 1.0 / math.sqrt(step) if step > 0 else float("inf")
        return learning_rate * inv_sq_dim * min(inv_sq_step, step * inv_sq3_warmup)
    return value
