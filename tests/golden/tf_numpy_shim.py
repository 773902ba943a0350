"""A numpy stand-in for the handful of TensorFlow-1 ops that the *pure* functions of the reference's
hot path use, so that those functions - the reference's own code, imported from /root/reference - can
be executed in this container and their outputs committed as golden vectors.

What this pins and what it does not: the executed code is the reference's (op order, constants,
masking rules, reshapes); the semantics of each TF op are restated here in numpy from the TF-1.12
documentation.  Every op below is a one-liner over numpy, fp32 in / fp32 out.

Used only by tests/golden/make_tf_shim_golden.py.
"""
import contextlib
import sys
import types

import numpy as np


class Dim(int):
    """A static dimension: an int with the `.value` TF-1 dimensions carry."""

    @property
    def value(self):
        return int(self)


class VarScope:
    """What a model part keeps in `_variable_scope`: re-entering it makes its name the full scope."""

    def __init__(self, name):
        self.name, self.reuse, self.original_name_scope = name, False, name + "/"


class Shape(tuple):
    """What `tensor.shape` / `get_shape()` return: a tuple with `.as_list()` and `.ndims`."""

    def __new__(cls, dims=()):
        return super().__new__(cls, [None if d is None else Dim(d) for d in dims])

    def as_list(self):
        return list(self)

    @property
    def ndims(self):
        return len(self)

    @property
    def dims(self):
        return list(self)


class T(np.ndarray):
    """ndarray whose `.shape` behaves like a TensorShape."""

    @property
    def shape(self):  # pylint: disable=invalid-overridden-method
        return Shape(np.ndarray.shape.__get__(self))

    def get_shape(self):
        return self.shape

    def set_shape(self, _shape):
        return None

    # TF tensors are immutable: `x += y` rebinds the name, it must not write into x's buffer
    def __iadd__(self, other):
        return self + other

    def __imul__(self, other):
        return self * other

    def __isub__(self, other):
        return self - other


def t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(T)


class _Missing:
    """Anything of TF that is touched at import time only (annotations, default arguments)."""

    def __init__(self, name):
        self._name = name

    def __getattr__(self, item):
        return _Missing(self._name + "." + item)

    def __call__(self, *args, **kwargs):
        return _Missing(self._name + "()")

    def __getitem__(self, item):
        return self

    def __mro_entries__(self, bases):
        return (object,)


class _Namespace(types.ModuleType):
    def __getattr__(self, item):
        return _Missing(self.__name__ + "." + item)


VARIABLES = {}      # full variable name -> numpy array (layer_norm gamma / beta ...)
DENSE = {}          # dense layer name -> (kernel, bias or None)
USED = []           # full names of the dense kernels that were looked up (checks the naming scheme)
TRAINABLE = []      # what tf.trainable_variables() returns: T arrays with a `.name`
_scope = []
GLOBAL_STEP = [0]


@contextlib.contextmanager
def _scope_cm(name, *args, **kwargs):
    saved = list(_scope)
    if isinstance(name, VarScope):
        _scope[:] = [name.name]            # a scope OBJECT replaces the current scope
    else:
        _scope.append(name if isinstance(name, str) else "")
    try:
        yield
    finally:
        _scope[:] = saved


@contextlib.contextmanager
def _name_scope_cm(*args, **kwargs):
    yield                                  # op names do not matter here


def _conv2d_1x1(value, filters, strides, padding):
    """tf.nn.conv2d for the 1x1 / stride-1 case the attention key projection uses (NHWC x HWIO)."""
    assert filters.shape[0] == 1 and filters.shape[1] == 1 and list(strides) == [1, 1, 1, 1]
    return t(np.asarray(value) @ np.asarray(filters)[0, 0], np.float32)


def _full_name(name):
    return "/".join([s for s in _scope if s] + [name])


AUTO = [None]       # a numpy RandomState: variables nobody provided are then drawn on demand and recorded


def _auto(full, shape, ones=False):
    if AUTO[0] is None:
        raise KeyError("variable '{}' was not provided to the shim".format(full))
    value = np.asarray(AUTO[0].randn(*[int(d) for d in shape]) * 0.3, np.float32)
    VARIABLES[full] = value + 1.0 if ones else value
    return VARIABLES[full]


def _get_variable(name, shape=None, dtype=None, initializer=None, **kwargs):
    full = _full_name(name)
    if full not in VARIABLES:
        _auto(full, shape, ones=name == "gamma")
    return t(VARIABLES[full], np.float32)


def _dense(inputs, units, activation=None, use_bias=True, name=None, **kwargs):
    # a layer is looked up by its full TF variable name first (VARIABLES["<scope>/<name>/kernel"]),
    # then by its bare name in DENSE (the early, scope-less fixtures)
    full = _full_name(name or "dense")
    if full + "/kernel" in VARIABLES:
        kernel, bias = VARIABLES[full + "/kernel"], VARIABLES.get(full + "/bias")
        USED.append(full + "/kernel")
    elif name in DENSE and AUTO[0] is None:
        kernel, bias = DENSE[name]
    else:
        kernel = _auto(full + "/kernel", (np.shape(inputs)[-1], units))
        bias = _auto(full + "/bias", (units,)) if use_bias else None
        USED.append(full + "/kernel")
    assert kernel.shape[1] == units
    out = np.asarray(inputs) @ kernel
    if use_bias and bias is not None:
        out = out + bias
    out = t(out, np.float32)
    return activation(out) if activation is not None else out


def _matrix_band_part(x, lower, upper):
    rows, cols = x.shape[-2:]
    i, j = np.arange(rows)[:, None], np.arange(cols)[None, :]
    keep = ((lower < 0) | (i - j <= lower)) & ((upper < 0) | (j - i <= upper))
    return t(np.asarray(x) * keep)


def _softmax(x, axis=-1):
    x = np.asarray(x, np.float32)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return t(e / e.sum(axis=axis, keepdims=True), np.float32)


def _max_pool(value, ksize, strides, padding):
    """NHWC max pooling, only the window == stride, exact-division case the maxout layer uses."""
    assert list(ksize) == list(strides) and padding == "SAME"
    n, h, w, c = value.shape
    kh, kw = ksize[1], ksize[2]
    assert h % kh == 0 and w % kw == 0
    x = np.asarray(value).reshape(n, h // kh, kh, w // kw, kw, c)
    return t(x.max(axis=(2, 4)))


def _gather_nd(params, indices):
    indices = np.asarray(indices)
    return t(np.asarray(params)[tuple(indices[..., k] for k in range(indices.shape[-1]))])


def _pad(x, paddings):
    return t(np.pad(np.asarray(x), [(int(a), int(b)) for a, b in paddings]))


def _map_structure(fn, *structures):
    """tf.contrib.framework.nest.map_structure over (named) tuples and lists of tensors."""
    first = structures[0]
    if isinstance(first, tuple) and hasattr(first, "_fields"):
        return type(first)(*[_map_structure(fn, *parts) for parts in zip(*structures)])
    if isinstance(first, (list, tuple)):
        return type(first)(_map_structure(fn, *parts) for parts in zip(*structures))
    return fn(*structures)


def _top_k(values, k):
    """tf.nn.top_k over the last axis: descending, the lower index first among equal values."""
    values = np.asarray(values)
    order = np.argsort(-values, axis=-1, kind="stable")[..., :k]
    return t(np.take_along_axis(values, order, axis=-1)), t(order.astype(np.int32))


def _one_hot(index, depth, dtype=np.float32, on_value=1.0, off_value=0.0):
    idx = np.asarray(index).astype(np.int64)
    out = np.full(idx.shape + (int(depth),), off_value, dtype=dtype)
    np.put_along_axis(out, idx[..., None], on_value, axis=-1)
    return t(out)


def _log_softmax(x, axis=-1):
    x = np.asarray(x, np.float32)
    shifted = x - x.max(axis=axis, keepdims=True)
    return t(shifted - np.log(np.exp(shifted).sum(axis=axis, keepdims=True)), np.float32)


class GRUCell:
    """tf.contrib.rnn.GRUCell as PUBLISHED by TensorFlow 1.x (rnn_cell_impl.GRUCell.call) -- the cell
    arithmetic is library code the reference does not carry, so this is a restatement, not reference
    code being run:

        [r, u] = sigmoid([x, h] . gates/kernel + gates/bias)
        c      = tanh([x, r * h] . candidate/kernel + candidate/bias)
        h'     = u * h + (1 - u) * c                  (returned as both output and state)

    What running the reference AROUND it pins is everything else: which tensors are fed as x and h,
    the variable scope the cell is called in (ortho_gru_cell.py:51 passes scope="OrthoGRUCell"), and
    what is done with the result.  CALLS records (scope, input shape, state shape) per call."""
    CALLS = []

    def __init__(self, num_units, activation=None, reuse=None, kernel_initializer=None,
                 bias_initializer=None, **kwargs):
        self._num_units = num_units
        self._activation = activation or (lambda x: t(np.tanh(np.asarray(x, np.float32)), np.float32))

    @property
    def state_size(self):
        return self._num_units

    @property
    def output_size(self):
        return self._num_units

    def __call__(self, inputs, state, scope=None):
        # Layer.__call__: open the given scope, or the layer's default name (snake-cased class name -
        # TensorFlow-internal naming, not visible in the reference), then run `call`
        import re
        default = re.sub("([a-z])([A-Z])", r"\1_\2", re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", type(self).__name__)).lower()
        with _scope_cm(scope or default):
            GRUCell.CALLS.append((_full_name(""), tuple(np.shape(inputs)), tuple(np.shape(state))))
            return self.call(inputs, state)

    def call(self, inputs, state):
        x, h = np.asarray(inputs, np.float32), np.asarray(state, np.float32)
        wide, n = x.shape[1] + h.shape[1], h.shape[1]
        wg, bg = _get_variable("gates/kernel", [wide, 2 * n]), _get_variable("gates/bias", [2 * n])
        wc, bc = _get_variable("candidate/kernel", [wide, n]), _get_variable("candidate/bias", [n])
        assert h.shape[1] == self._num_units and wg.shape == (x.shape[1] + h.shape[1], 2 * h.shape[1])
        gates = 1.0 / (1.0 + np.exp(-(np.concatenate([x, h], 1) @ wg + bg)))
        r, u = gates[:, :h.shape[1]], gates[:, h.shape[1]:]
        c = np.asarray(self._activation(t(np.concatenate([x, r * h], 1) @ wc + bc, np.float32)))
        new_h = t(u * h + (1.0 - u) * c, np.float32)
        return new_h, new_h


LSTMStateTuple = __import__("collections").namedtuple("LSTMStateTuple", ["c", "h"])


class LSTMCell(GRUCell):
    """tf.nn.rnn_cell.LSTMCell with its defaults (no peepholes, no projection, forget_bias = 1), as
    PUBLISHED by TensorFlow 1.x - library code restated, like GRUCell above:
        [i, j, f, o] = [x, h] . kernel + bias;  c' = sigmoid(f + 1) * c + sigmoid(i) * tanh(j);
        h' = sigmoid(o) * tanh(c');  output h', state (c', h')"""

    @property
    def state_size(self):
        return LSTMStateTuple(self._num_units, self._num_units)

    def call(self, inputs, state):
        c, h = (np.asarray(s, np.float32) for s in state)
        wide = np.shape(inputs)[1] + h.shape[1]
        kernel, bias = _get_variable("kernel", [wide, 4 * h.shape[1]]), _get_variable("bias", [4 * h.shape[1]])
        z = np.concatenate([np.asarray(inputs, np.float32), h], 1) @ kernel + bias
        i, j, f, o = np.split(z, 4, axis=1)
        sig = lambda v: 1.0 / (1.0 + np.exp(-v))
        new_c = sig(f + 1.0) * c + sig(i) * np.tanh(j)
        new_h = sig(o) * np.tanh(new_c)
        return t(new_h, np.float32), LSTMStateTuple(t(new_c, np.float32), t(new_h, np.float32))


def _rnn_loop(cell, inputs, lengths):
    """The recurrence of tf.nn.dynamic_rnn as TensorFlow documents it (library code, restated): zero
    initial state; past a sentence's length the output is zero and the state is carried unchanged."""
    x = np.asarray(inputs, np.float32)
    lengths = np.full((x.shape[0],), x.shape[1]) if lengths is None else np.asarray(lengths)
    tupled = isinstance(cell.state_size, tuple)
    sizes = cell.state_size if tupled else (cell.state_size,)
    state = [np.zeros((x.shape[0], n), np.float32) for n in sizes]
    outputs = []
    for step in range(x.shape[1]):
        fed = type(cell.state_size)(*[t(s) for s in state]) if tupled else t(state[0])
        out, new = cell(t(x[:, step]), fed)
        alive = (step < lengths)[:, None]
        new = list(new) if tupled else [new]
        state = [np.where(alive, np.asarray(n), s) for n, s in zip(new, state)]
        outputs.append(np.where(alive, np.asarray(out), 0.0))
    final = type(cell.state_size)(*[t(s, np.float32) for s in state]) if tupled else t(state[0], np.float32)
    return t(np.stack(outputs, 1), np.float32), final


def _reverse_sequence(x, lengths, seq_axis=1, batch_axis=0):
    assert seq_axis == 1 and batch_axis == 0
    out = np.array(np.asarray(x))
    for b, n in enumerate(np.asarray(lengths)):
        out[b, :n] = out[b, :n][::-1]
    return t(out)


def _dynamic_rnn(cell, inputs, sequence_length=None, dtype=None, scope=None, **kwargs):
    with _scope_cm(scope or "rnn"):
        return _rnn_loop(cell, inputs, sequence_length)


def _bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, sequence_length=None, dtype=None, scope=None,
                               **kwargs):
    """tf.nn.bidirectional_dynamic_rnn (library code, restated): scopes bidirectional_rnn/{fw,bw}; the
    backward direction runs over the length-reversed input and its outputs are reversed back."""
    with _scope_cm(scope or "bidirectional_rnn"):
        with _scope_cm("fw"):
            out_fw, st_fw = _rnn_loop(cell_fw, inputs, sequence_length)
        with _scope_cm("bw"):
            lengths = (np.full((np.shape(inputs)[0],), np.shape(inputs)[1]) if sequence_length is None
                       else sequence_length)
            out_bw, st_bw = _rnn_loop(cell_bw, _reverse_sequence(inputs, lengths), lengths)
            out_bw = _reverse_sequence(out_bw, lengths)
    return (out_fw, out_bw), (st_fw, st_bw)


def _softmax_cross_entropy(onehot_labels, logits, weights=1.0, label_smoothing=0, **kwargs):
    """tf.losses.softmax_cross_entropy as TensorFlow documents it (library code, restated): the labels are
    smoothed to onehot * (1 - s) + s / num_classes, and the result is REDUCED to one scalar - the mean over
    the rows (reduction SUM_BY_NONZERO_WEIGHTS with weights 1)."""
    onehot = np.asarray(onehot_labels, np.float32)
    if label_smoothing:
        onehot = onehot * (1.0 - label_smoothing) + label_smoothing / onehot.shape[-1]
    losses = -(onehot * np.asarray(_log_softmax(logits))).sum(-1)
    return t(np.float32(losses.mean()), np.float32)


def _sequence_loss(logits, targets, weights, average_across_timesteps=True, average_across_batch=True,
                   softmax_loss_function=None, name=None):
    """tf.contrib.seq2seq.sequence_loss as TensorFlow documents it (library code, restated), without
    averaging: logits and targets are flattened, the loss function is applied (default: sparse softmax
    cross-entropy per row), the result is multiplied by the flattened weights and reshaped to [B, T]."""
    assert not average_across_timesteps and not average_across_batch
    lg = np.asarray(logits, np.float32)
    flat_logits, flat_targets = lg.reshape(-1, lg.shape[-1]), np.asarray(targets).reshape(-1)
    if softmax_loss_function is None:
        logprobs = np.asarray(_log_softmax(t(flat_logits)))
        crossent = -np.take_along_axis(logprobs, flat_targets[:, None].astype(np.int64), -1)[:, 0]
    else:
        crossent = np.asarray(softmax_loss_function(labels=t(flat_targets), logits=t(flat_logits)))
    crossent = crossent * np.asarray(weights, np.float32).reshape(-1)
    return t(crossent.reshape(lg.shape[0], lg.shape[1]), np.float32)


def feature_dropout_mask(shape, keep_prob):
    """A DETERMINISTIC stand-in for a dropout mask, the same for the reference run and for the product: along
    the last axis every second entry is dropped, the others scaled by 1/keep_prob.  Random streams cannot
    be matched, but with this both sides can be compared element by element, which pins WHERE dropout
    is applied (twice on the initial state, on the GRU output that becomes the next state, ...)."""
    last = int(shape[-1])
    pattern = np.where(np.arange(last) % 2 == 0, 1.0 / keep_prob, 0.0).astype(np.float32)
    return np.broadcast_to(pattern, tuple(int(d) for d in shape)).copy()


def _nn_dropout(x, keep_prob, **kwargs):
    x = np.asarray(x, np.float32)
    return t(x * feature_dropout_mask(x.shape, keep_prob), np.float32)


def _while_loop(cond, body, loop_vars, shape_invariants=None, **kwargs):
    """tf.while_loop run eagerly: the loop variables are one (named) tuple passed unpacked."""
    state = loop_vars
    while bool(np.asarray(cond(*state))):
        state = body(*state)
    return state


def install():
    """Put the shim into sys.modules as `tensorflow` and return it."""
    tf = _Namespace("tensorflow")
    tf.float32, tf.int32, tf.int64, tf.bool = np.float32, np.int32, np.int64, np.bool_
    tf.Tensor = T
    tf.TensorShape = Shape
    tf.to_float = lambda x: t(x, np.float32)
    tf.to_int32 = lambda x: t(x, np.int32)
    tf.to_int64 = lambda x: t(x, np.int64)
    tf.range = lambda *a, dtype=None: t(np.arange(*[int(v) for v in a]), dtype or np.int32)
    tf.exp = lambda x: t(np.exp(np.asarray(x, np.float32)), np.float32)
    tf.sin = lambda x: t(np.sin(np.asarray(x, np.float32)), np.float32)
    tf.cos = lambda x: t(np.cos(np.asarray(x, np.float32)), np.float32)
    tf.sqrt = lambda x: t(np.sqrt(np.asarray(x, np.float32)), np.float32)
    tf.rsqrt = lambda x: t(1.0 / np.sqrt(np.asarray(x, np.float32)), np.float32)
    tf.square = lambda x: t(np.square(np.asarray(x)))
    tf.tanh = lambda x: t(np.tanh(np.asarray(x, np.float32)), np.float32)
    tf.zeros = lambda shape, dtype=None, name=None: t(np.zeros(
        [int(d) for d in (shape if isinstance(shape, (list, tuple)) else [shape])], dtype or np.float32))
    tf.reduce_sum = lambda x, axis=None, keepdims=False: t(
        np.sum(np.asarray(x), axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims))
    tf.Variable = T
    tf.minimum = lambda a, b: t(np.minimum(a, b))
    tf.mod = lambda a, b: t(np.mod(np.asarray(a), b)) if isinstance(a, np.ndarray) else np.mod(a, b)
    tf.expand_dims = lambda x, axis: t(np.expand_dims(
        np.asarray(x, np.float32) if isinstance(x, list) else np.asarray(x), axis))
    tf.concat = lambda values, axis: t(np.concatenate([np.asarray(v) for v in values], axis=axis))
    tf.pad = _pad
    tf.reshape = lambda x, shape, name=None: t(np.reshape(np.asarray(x), [int(s) for s in shape]))
    tf.shape = lambda x: Shape(np.asarray(x).shape)
    tf.transpose = lambda x, perm=None: t(np.transpose(np.asarray(x), perm))
    tf.ones_like = lambda x: t(np.ones_like(np.asarray(x)))
    tf.matrix_band_part = _matrix_band_part
    tf.equal = lambda a, b: t(np.equal(a, b))
    tf.fill = lambda dims, value: t(np.full([int(d) for d in dims], value, np.float32))
    tf.where = lambda cond, a, b: t(np.where(np.asarray(cond), np.asarray(a), np.asarray(b)))
    tf.matmul = lambda a, b, transpose_b=False: t(
        np.matmul(np.asarray(a), np.swapaxes(np.asarray(b), -1, -2) if transpose_b else np.asarray(b)))
    tf.identity = lambda x, name=None: x
    tf.convert_to_tensor = lambda x, *a, **k: x if isinstance(x, T) else t(x)
    tf.gather_nd = _gather_nd
    tf.reduce_mean = lambda x, axis=None, keepdims=False: t(
        np.mean(np.asarray(x), axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims))
    tf.variable_scope = _scope_cm
    tf.name_scope = _name_scope_cm
    tf.get_variable = _get_variable
    tf.get_variable_scope = lambda: types.SimpleNamespace(name="/".join(s for s in _scope if s))
    tf.ones_initializer = lambda: None
    tf.zeros_initializer = lambda: None
    tf.AUTO_REUSE = object()
    tf.nn = _Namespace("tensorflow.nn")
    tf.nn.softmax = _softmax
    tf.nn.max_pool = _max_pool
    tf.nn.conv2d = _conv2d_1x1
    tf.nn.relu = lambda x: t(np.maximum(np.asarray(x), 0))
    tf.layers = _Namespace("tensorflow.layers")
    tf.layers.dense = _dense
    tf.train = _Namespace("tensorflow.train")
    tf.train.get_or_create_global_step = lambda: t(GLOBAL_STEP[0], np.int64)
    tf.contrib = _Namespace("tensorflow.contrib")
    tf.contrib.framework = _Namespace("tensorflow.contrib.framework")
    tf.contrib.framework.nest = _Namespace("tensorflow.contrib.framework.nest")
    tf.contrib.framework.nest.map_structure = _map_structure
    tf.nn.top_k = _top_k
    tf.nn.log_softmax = _log_softmax
    tf.one_hot = _one_hot
    tf.tile = lambda x, multiples, name=None: t(np.tile(np.asarray(x), [int(m) for m in multiples]))
    tf.placeholder_with_default = lambda x, shape=None, name=None: x
    tf.stack = lambda values, axis=0: t(np.stack([np.asarray(v) for v in values], axis=axis))
    tf.div = lambda a, b: t(np.floor_divide(np.asarray(a), b))
    tf.logical_or = lambda a, b: t(np.logical_or(a, b))
    tf.logical_not = lambda a: t(np.logical_not(a))
    tf.logical_and = lambda a, b: t(np.logical_and(a, b))
    tf.less = lambda a, b: t(np.less(a, b))
    tf.reduce_all = lambda x: t(np.all(np.asarray(x)))
    tf.constant = lambda x, dtype=None, **k: t(x, dtype)
    tf.nn.embedding_lookup = lambda table, ids: t(np.asarray(table)[np.asarray(ids)])
    tf.argmax = lambda x, axis=None: t(np.argmax(np.asarray(x), axis=axis).astype(np.int64))
    tf.contrib.rnn = _Namespace("tensorflow.contrib.rnn")
    tf.contrib.rnn.GRUCell = GRUCell
    tf.contrib.rnn.RNNCell = GRUCell
    tf.contrib.rnn.LSTMCell = LSTMCell
    tf.contrib.rnn.LSTMStateTuple = LSTMStateTuple
    tf.nn.rnn_cell = _Namespace("tensorflow.nn.rnn_cell")
    tf.nn.rnn_cell.RNNCell = GRUCell
    tf.nn.rnn_cell.LSTMCell = tf.contrib.rnn.LSTMCell
    tf.nn.dynamic_rnn = _dynamic_rnn
    tf.nn.bidirectional_dynamic_rnn = _bidirectional_dynamic_rnn
    tf.reverse_sequence = _reverse_sequence
    tf.not_equal = lambda a, b: t(np.not_equal(a, b))
    tf.while_loop = _while_loop
    tf.nn.dropout = _nn_dropout
    tf.ones = lambda shape, dtype=None, name=None: t(np.ones([int(d) for d in shape], dtype or np.float32))
    tf.sigmoid = lambda x: t(1.0 / (1.0 + np.exp(-np.asarray(x, np.float32))), np.float32)
    tf.split = lambda value, num_or_size_splits, axis=0: [t(p) for p in np.split(np.asarray(value), num_or_size_splits, axis=axis)]
    tf.trainable_variables = lambda: list(TRAINABLE)
    tf.get_collection = lambda key, scope=None: [v for v in TRAINABLE if scope is None or v.name.startswith(scope)]
    tf.GraphKeys = types.SimpleNamespace(TRAINABLE_VARIABLES="trainable_variables")
    # tf.clip_by_norm as documented: t * clip_norm / max(l2norm(t), clip_norm)
    tf.clip_by_norm = lambda g, clip, axes=None, name=None: t(
        np.asarray(g) * clip / max(float(np.sqrt(np.sum(np.square(np.asarray(g, np.float32))))), clip), np.float32)
    tf.contrib.seq2seq = _Namespace("tensorflow.contrib.seq2seq")
    tf.contrib.seq2seq.sequence_loss = _sequence_loss
    tf.losses = _Namespace("tensorflow.losses")
    tf.losses.softmax_cross_entropy = _softmax_cross_entropy
    tf.control_dependencies = _name_scope_cm
    tf.squeeze = lambda x, axis=None: t(np.squeeze(np.asarray(x), axis=tuple(axis) if isinstance(axis, list) else axis))
    tf.orthogonal_initializer = lambda *a, **k: None
    tf.random_normal_initializer = lambda *a, **k: None
    sys.modules["tensorflow"] = tf
    for sub in ("tensorflow.contrib", "tensorflow.contrib.slim", "tensorflow.contrib.slim.nets",
                "tensorflow.python", "tensorflow.python.framework", "tensorflow.contrib.tensorboard",
                "tensorflow.contrib.tensorboard.plugins"):
        sys.modules[sub] = _Namespace(sub)
    return tf
