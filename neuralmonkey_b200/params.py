"""Flat fp32 parameter arena.

All trainable variables of an experiment live in ONE contiguous device buffer
(`params`), with matching flat buffers for gradients and the Adam moments.  A
variable is a named segment; model parts get tensor *views* into the arena whose
``nm_grad`` attribute is the matching view of the gradient buffer, which is where
the kernels of ``neuralmonkey_b200.ops`` accumulate weight gradients.

Why flat: the optimizer (K13) is two HBM-bound passes over one buffer instead of
hundreds of tiny launches, and the data-parallel gradient exchange (K14) is one
NCCL all-reduce over one buffer.  Variable names follow the reference's TF variable
scopes so checkpoints can be keyed the same way.
"""
import math
import re
import zlib
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

Initializer = Callable[[Sequence[int], torch.Generator], torch.Tensor]

_BIAS_REGEX = re.compile(r"[Bb]ias")  # trainers/generic_trainer.py:17


def is_regularizable(name: str) -> bool:
    """Variables that receive L1/L2 terms (generic_trainer.py:87-91)."""
    return (not _BIAS_REGEX.findall(name) and not name.startswith("vgg")
            and not name.startswith("Inception") and not name.startswith("resnet"))


# ---- initialisers (the TF ones the reference's model parts use) -------------------
def normal_initializer(stddev: float = 0.001, mean: float = 0.0) -> Initializer:
    def init(shape, gen):
        return torch.randn(*shape, generator=gen, dtype=torch.float64) * stddev + mean
    return init


def uniform_initializer(minval: float, maxval: float) -> Initializer:
    def init(shape, gen):
        return torch.rand(*shape, generator=gen, dtype=torch.float64) * (maxval - minval) + minval
    return init


def constant_initializer(value: float) -> Initializer:
    def init(shape, gen):
        return torch.full(tuple(shape), float(value), dtype=torch.float64)
    return init


def zeros_initializer() -> Initializer:
    return constant_initializer(0.0)


def ones_initializer() -> Initializer:
    return constant_initializer(1.0)


def orthogonal_initializer() -> Initializer:
    """tf.orthogonal_initializer for 2-D kernels (nn/ortho_gru_cell.py:47-53)."""
    def init(shape, gen):
        rows, cols = shape
        a = torch.randn(max(rows, cols), min(rows, cols), generator=gen, dtype=torch.float64)
        q, r = torch.linalg.qr(a)
        q = q * torch.sign(torch.diagonal(r))
        if rows < cols:
            q = q.t()
        return q[:rows, :cols].contiguous()
    return init


def block_orthogonal_initializer() -> Initializer:
    """The reference's own `orthogonal_initializer()` (nn/ortho_gru_cell.py:4-37): a [dim, m*dim] kernel
    is m independent dim x dim orthogonal blocks side by side (the left singular vectors of a normal
    matrix there; the Q factor of one here - both are Haar-distributed orthogonal matrices)."""
    square = orthogonal_initializer()

    def init(shape, gen):
        if len(shape) != 2:
            raise ValueError("Orthogonal initializer only works with 2D matrices.")
        rows, cols = shape
        if cols % rows != 0:
            raise ValueError("Shape {} is not compatible with orthogonal initializer.".format(str(shape)))
        return torch.cat([square((rows, rows), gen) for _ in range(cols // rows)], 1).contiguous()
    return init


def variance_scaling_initializer(scale: float = 1.0, mode: str = "fan_avg",
                                 distribution: str = "uniform") -> Initializer:
    """tf.variance_scaling_initializer (encoders/transformer.py:152-153)."""
    def init(shape, gen):
        fan_in = shape[0] if len(shape) >= 1 else 1
        fan_out = shape[-1] if len(shape) >= 2 else fan_in
        if len(shape) > 2:  # conv kernels HWIO
            rf = 1
            for s in shape[:-2]:
                rf *= s
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[mode]
        if distribution == "uniform":
            limit = math.sqrt(3.0 * scale / n)
            return (torch.rand(*shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit
        return torch.randn(*shape, generator=gen, dtype=torch.float64) * math.sqrt(scale / n)
    return init


class Variable:
    """Declaration of one named segment."""

    def __init__(self, name: str, shape: Sequence[int], initializer: Initializer,
                 trainable: bool = True) -> None:
        self.name = name
        self.shape = tuple(int(s) for s in shape)
        self.initializer = initializer
        self.trainable = trainable
        self.offset = -1

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


class ParameterArena:
    """Declare variables, then `finalize()` once to allocate the flat buffers."""

    ALIGN = 64  # floats: every segment starts 256-byte aligned (TMA needs 16 B)
    STAT_SLOTS = 8

    def __init__(self) -> None:
        self.variables = {}  # type: Dict[str, Variable]
        self.order = []  # type: List[str]
        self.finalized = False
        self.params = None  # type: Optional[torch.Tensor]
        self.grads = None  # type: Optional[torch.Tensor]
        self.adam_m = None  # type: Optional[torch.Tensor]
        self.adam_v = None  # type: Optional[torch.Tensor]
        self._views = {}  # type: Dict[str, torch.Tensor]
        # segment tables for nm_clip_adam_step (device)
        self.seg_off = None  # type: Optional[torch.Tensor]
        self.seg_reg = None  # type: Optional[torch.Tensor]
        self.seg_norms = None  # type: Optional[torch.Tensor]
        self.size = 0
        self.train_names = []  # type: List[str]
        self.trainable_size = 0

    def declare(self, name: str, shape: Sequence[int], initializer: Initializer,
                trainable: bool = True) -> None:
        if self.finalized:
            raise RuntimeError("arena already finalized; cannot declare '{}'".format(name))
        if name in self.variables:
            old = self.variables[name]
            if old.shape != tuple(shape):
                raise ValueError("variable '{}' redeclared with shape {} != {}".format(
                    name, tuple(shape), old.shape))
            return  # AUTO_REUSE semantics (model/parameterized.py:90-99)
        self.variables[name] = Variable(name, shape, initializer, trainable)
        self.order.append(name)

    def finalize(self, device: torch.device, seed: int = 2574600) -> None:
        """Allocate and initialise.  Trainable variables first (one contiguous trainable
        prefix is what the optimizer and the all-reduce walk), frozen ones after."""
        if self.finalized:
            return
        # Layout and initial values must not depend on the order in which model parts happened to
        # declare their variables (it follows the iteration order of Python sets, which differs
        # between processes): data-parallel ranks all-reduce this buffer element by element, so
        # every rank needs the same layout and the same initial parameters.  Hence: sorted names,
        # and one generator per variable seeded from (seed, crc32(name)).
        names = (sorted(n for n in self.order if self.variables[n].trainable)
                 + sorted(n for n in self.order if not self.variables[n].trainable))
        off = 0
        for n in names:
            self.variables[n].offset = off
            off += -(-self.variables[n].numel // self.ALIGN) * self.ALIGN
            if self.variables[n].trainable:
                self.trainable_size = off
        if not any(self.variables[n].trainable for n in names):
            self.trainable_size = 0
        self.size = off
        self.order = names
        host = torch.zeros(max(off, 1), dtype=torch.float32)
        for n in names:
            var = self.variables[n]
            gen = torch.Generator().manual_seed((int(seed) * 1000003 + zlib.crc32(n.encode("utf-8"))) % (2 ** 63))
            val = var.initializer(var.shape, gen).to(torch.float32).reshape(-1)
            host[var.offset:var.offset + var.numel] = val
        self.params = host.to(device)
        tsz = max(self.trainable_size, 1)
        # the gradient buffer carries STAT_SLOTS extra floats behind the gradients: per-step
        # scalars (loss sum, token count) that ride along in the same all-reduce
        self.grad_buffer = torch.zeros(tsz + self.ALIGN, device=device, dtype=torch.float32)
        self.grads = self.grad_buffer[:tsz]
        self.stats = self.grad_buffer[tsz:tsz + self.STAT_SLOTS]
        self.adam_m = torch.zeros(tsz, device=device, dtype=torch.float32)
        self.adam_v = torch.zeros(tsz, device=device, dtype=torch.float32)
        # optimizer slots: slot 0 is (adam_m, adam_v); further optimizer OBJECTS of the experiment get their own
        # pair on first use (`optimizer_slot`), as every tf.train.Optimizer keeps its own moment variables
        self.optimizer_slots = []           # [(optimizer, m, v)], in order of first use
        self.restored_optimizer_state = {}  # slot index -> {"m": dict, "v": dict, "steps": int} from a checkpoint
        train_names = [n for n in names if self.variables[n].trainable]
        # segment i spans [seg_off[i], seg_off[i+1]) including its alignment padding (zeros)
        offs = [self.variables[n].offset for n in train_names] + [self.trainable_size]
        self.seg_off = torch.tensor(offs, dtype=torch.int64, device=device)
        self.seg_reg = torch.tensor([1 if is_regularizable(n) else 0 for n in train_names] or [0],
                                    dtype=torch.uint8, device=device)
        self.seg_norms = torch.zeros(max(len(train_names), 1), device=device, dtype=torch.float32)
        self.train_names = train_names
        self.finalized = True
        for n in names:
            self._make_view(n)

    def _make_view(self, name: str) -> torch.Tensor:
        var = self.variables[name]
        view = self.params[var.offset:var.offset + var.numel].view(var.shape)
        if var.trainable:
            view.requires_grad_(True)
            view.nm_grad = self.grads[var.offset:var.offset + var.numel].view(var.shape)
        self._views[name] = view
        return view

    def get(self, name: str) -> torch.Tensor:
        if not self.finalized:
            raise RuntimeError("arena not finalized")
        return self._views[name]

    def grad(self, name: str) -> torch.Tensor:
        return self._views[name].nm_grad

    def zero_grad(self) -> None:
        self.grad_buffer.zero_()

    def fold_autograd_grads(self) -> int:
        """Every op writes its weight gradients straight into the gradient buffer (`nm_grad` sinks).  If a
        model part uses a parameter in a plain torch expression instead, autograd leaves that gradient
        in the view's `.grad`, where the optimizer kernel would never see it: add it to the buffer.
        Returns how many views carried one (0 on every path built from the library's ops)."""
        folded = 0
        for name in self.train_names:
            view = self._views[name]
            if view.grad is not None:
                view.nm_grad.add_(view.grad)
                view.grad = None
                folded += 1
        return folded

    @property
    def allreduce_view(self) -> torch.Tensor:
        """Gradients + stat slots: the one buffer a data-parallel step exchanges."""
        return self.grad_buffer[:self.grads.numel() + self.STAT_SLOTS]

    def load_dict(self, values: Dict[str, torch.Tensor]) -> None:
        """Overwrite variables from a {TF-style name: tensor} dict (tests, checkpoints)."""
        with torch.no_grad():
            for n, v in values.items():
                if n not in self.variables:
                    raise KeyError("unknown variable '{}'".format(n))
                dst = self._views[n]
                src = v.detach().to(torch.float32).reshape(dst.shape)
                dst.copy_(src.to(dst.device))

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {n: self._views[n].detach().cpu().clone() for n in self.order}

    def optimizer_slot(self, optimizer) -> Tuple[torch.Tensor, torch.Tensor]:
        """(m, v) flat moment buffers of `optimizer`.  TensorFlow keeps Adam's moments (slot variables) and its
        beta-power accumulators per optimizer OBJECT: two trainers with an optimizer each (tests/bahdanau.ini)
        do not share them, one optimizer handed to two trainers does.  The first optimizer that updates the
        model gets `adam_m` / `adam_v`; every further one a fresh zero pair (order of first use, the same on
        every rank and in a continued run).  State restored from a checkpoint before the slot existed is
        applied when it is claimed."""
        for owner, m, v in self.optimizer_slots:
            if owner is optimizer:
                return m, v
        index = len(self.optimizer_slots)
        if index == 0:
            m, v = self.adam_m, self.adam_v
        else:
            m, v = torch.zeros_like(self.adam_m), torch.zeros_like(self.adam_v)
        self.optimizer_slots.append((optimizer, m, v))
        pending = self.restored_optimizer_state.pop(index, None)
        if pending is not None:
            if index > 0:
                self.load_moments(m, pending["m"])
                self.load_moments(v, pending["v"])
        optimizer.steps = int(pending["steps"]) if pending is not None else 0
        return m, v

    def moment_dict(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """A flat per-parameter buffer (Adam m / v) split by trainable variable name."""
        out = {}
        for n in self.train_names:
            var = self.variables[n]
            out[n] = flat[var.offset:var.offset + var.numel].detach().cpu().clone().view(var.shape)
        return out

    def load_moments(self, flat: torch.Tensor, values: Dict[str, torch.Tensor]) -> None:
        with torch.no_grad():
            for n, v in values.items():
                var = self.variables.get(n)
                if var is None or not var.trainable or tuple(v.shape) != var.shape:
                    continue
                flat[var.offset:var.offset + var.numel].copy_(v.reshape(-1).to(flat.device))

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return {n: self._views[n].nm_grad.detach().cpu().clone() for n in self.train_names}
