"""Model parts that feed pre-computed numpy arrays (reference:
neuralmonkey/encoders/numpy_stateful_filler.py:16-245): a vector per sentence (`StatefulFiller`),
a sequence of vectors (`TemporalFiller`) or a feature map (`SpatialFiller`, e.g. the convolutional
maps of a captioning model extracted beforehand instead of running `ImageNet` in the step).

The optional projections are `tf.layers.dense` / 1x1 `tf.layers.conv2d` there: variables
`<name>/dense/{kernel,bias}` and `<name>/conv2d[_1]/{kernel,bias}` (kernel [1,1,in,out]), Glorot
uniform kernels and zero biases (the TF layer defaults).  A 1x1 convolution is the dense layer over
the channel axis, so both run on the projection GEMM.
"""
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from neuralmonkey_b200 import ops, runtime
from neuralmonkey_b200.decorators import tensor
from neuralmonkey_b200.model.model_part import ModelPart
from neuralmonkey_b200.model.parameterized import InitializerSpecs
from neuralmonkey_b200.model.stateful import Stateful, SpatialStatefulWithOutput, TemporalStateful
from neuralmonkey_b200.params import variance_scaling_initializer, zeros_initializer
from neuralmonkey_b200.typecheck import check_argument_types


def _glorot():
    return variance_scaling_initializer(1.0, "fan_avg", "uniform")


class _NumpyFed(ModelPart):
    """Shared feeding protocol: the batch lives in `self._fed` (name -> device tensor)."""

    def _feed(self, train: bool, **arrays: np.ndarray) -> None:
        self.reset_batch()
        self.train_mode = bool(train)
        self._fed = {k: runtime.to_device(torch.from_numpy(np.ascontiguousarray(v))) for k, v in arrays.items()}
        self.batch_size = int(next(iter(arrays.values())).shape[0])

    def static_inputs(self) -> Dict[str, Any]:
        return dict(getattr(self, "_fed", None) or {})

    def bind_static(self, tensors: Dict[str, Any]) -> None:
        self.reset_batch()
        if tensors:
            self._fed = dict(tensors)

    @property
    def input_types(self) -> Dict[str, Any]:
        return {self.data_id: float}


class StatefulFiller(_NumpyFed, Stateful):
    # pylint: disable=too-many-arguments
    def __init__(self, name: str, dimension: int, data_id: str, output_shape: int = None,
                 reuse: ModelPart = None, save_checkpoint: str = None, load_checkpoint: str = None,
                 initializers: InitializerSpecs = None) -> None:
        check_argument_types()
        ModelPart.__init__(self, name, reuse, save_checkpoint, load_checkpoint, initializers)
        self.data_id = data_id
        self.dimension = dimension
        self.output_shape = output_shape
        if self.dimension <= 0:
            raise ValueError("Input vector dimension must be positive.")
        if self.output_shape is not None and self.output_shape <= 0:
            raise ValueError("Output vector dimension must be positive.")

    @property
    def _projects(self) -> bool:
        return self.output_shape is not None and self.output_shape != self.dimension

    @property
    def output_dimension(self) -> int:
        return self.output_shape if self._projects else self.dimension

    def declare_variables(self) -> None:
        if self._projects:
            self.declare("dense/kernel", [self.dimension, self.output_shape], _glorot())
            self.declare("dense/bias", [self.output_shape], zeros_initializer())

    @property
    def input_shapes(self) -> Dict[str, Any]:
        return {self.data_id: [None, self.dimension]}

    def feed_dict(self, dataset, train: bool = False) -> Dict[str, Any]:
        fd = ModelPart.feed_dict(self, dataset, train)
        vectors = np.asarray(list(dataset.get_series(self.data_id)), dtype=np.float32)
        if vectors.ndim != 2 or vectors.shape[1] != self.dimension:
            raise ValueError("StatefulFiller '{}' expects vectors of size {}, got {}".format(
                self.name, self.dimension, vectors.shape[1:]))
        self._feed(train, vector=vectors)
        fd[self.data_id] = vectors
        return fd

    @tensor
    def vector(self) -> torch.Tensor:
        return self._fed["vector"]

    @tensor
    def output(self) -> torch.Tensor:
        if not self._projects:
            return self.vector
        return ops.linear(self.vector, self.var("dense/kernel"), self.var("dense/bias"))


class TemporalFiller(_NumpyFed, TemporalStateful):
    # pylint: disable=too-many-arguments
    def __init__(self, name: str, data_id: str, input_size: int, max_input_len: int = None,
                 dropout_keep_prob: float = 1.0, reuse: ModelPart = None, save_checkpoint: str = None,
                 load_checkpoint: str = None, initializers: InitializerSpecs = None) -> None:
        check_argument_types()
        ModelPart.__init__(self, name, reuse, save_checkpoint, load_checkpoint, initializers)
        self.data_id = data_id
        self.input_size = input_size
        self.max_input_len = max_input_len
        self.dropout_keep_prob = dropout_keep_prob      # accepted and unused, as in the reference

    @property
    def input_shapes(self) -> Dict[str, Any]:
        return {self.data_id: [None, None, self.input_size]}

    def feed_dict(self, dataset, train: bool = False) -> Dict[str, Any]:
        fd = ModelPart.feed_dict(self, dataset, train)
        series = [np.asarray(x, dtype=np.float32) for x in dataset.get_series(self.data_id)]
        max_len = max(x.shape[0] for x in series)
        if self.max_input_len is not None:
            max_len = min(self.max_input_len, max_len)
        padded = np.zeros((len(series), max_len) + series[0].shape[1:], dtype=np.float32)
        lengths = np.zeros(len(series), dtype=np.int64)
        for i, x in enumerate(series):
            lengths[i] = min(max_len, x.shape[0])
            padded[i, :lengths[i]] = x[:lengths[i]]
        mask = (np.arange(max_len)[None, :] < lengths[:, None]).astype(np.float32)
        self._feed(train, states=padded, mask=mask)
        fd[self.data_id] = padded
        return fd

    @tensor
    def temporal_states(self) -> torch.Tensor:
        return self._fed["states"]

    @tensor
    def temporal_mask(self) -> torch.Tensor:
        return self._fed["mask"]

    @property
    def dimension(self) -> int:
        return self.input_size


class SpatialFiller(_NumpyFed, SpatialStatefulWithOutput):
    # pylint: disable=too-many-arguments
    def __init__(self, name: str, input_shape: List[int], data_id: str, projection_dim: int = None,
                 ff_hidden_dim: int = None, reuse: ModelPart = None, save_checkpoint: str = None,
                 load_checkpoint: str = None, initializers: InitializerSpecs = None) -> None:
        check_argument_types()
        ModelPart.__init__(self, name, reuse, save_checkpoint, load_checkpoint, initializers)
        self.data_id = data_id
        self.input_shape = input_shape
        self.projection_dim = projection_dim
        self.ff_hidden_dim = ff_hidden_dim
        if self.ff_hidden_dim is not None and self.projection_dim is None:
            raise ValueError("projection_dim must be provided when using ff_hidden_dim")
        if len(self.input_shape) != 3:
            raise ValueError("The input shape should have 3 dimensions.")

    def _conv_scopes(self) -> List[Optional[str]]:
        """TF numbers the layers of one scope in creation order: conv2d, conv2d_1."""
        hidden = "conv2d" if self.ff_hidden_dim else None
        proj = None
        if self.projection_dim:
            proj = "conv2d_1" if hidden else "conv2d"
        return [hidden, proj]

    def declare_variables(self) -> None:
        hidden, proj = self._conv_scopes()
        channels = self.input_shape[2]
        if hidden:
            self.declare(hidden + "/kernel", [1, 1, channels, self.ff_hidden_dim], _glorot())
            self.declare(hidden + "/bias", [self.ff_hidden_dim], zeros_initializer())
            channels = self.ff_hidden_dim
        if proj:
            self.declare(proj + "/kernel", [1, 1, channels, self.projection_dim], _glorot())
            self.declare(proj + "/bias", [self.projection_dim], zeros_initializer())

    @property
    def input_shapes(self) -> Dict[str, Any]:
        return {self.data_id: [None] + list(self.input_shape)}

    def feed_dict(self, dataset, train: bool = False) -> Dict[str, Any]:
        fd = ModelPart.feed_dict(self, dataset, train)
        maps = np.asarray(list(dataset.get_series(self.data_id)), dtype=np.float32)
        if list(maps.shape[1:]) != list(self.input_shape):
            raise ValueError("SpatialFiller '{}' expects maps of shape {}, got {}".format(
                self.name, list(self.input_shape), list(maps.shape[1:])))
        self._feed(train, maps=maps)
        fd[self.data_id] = maps
        return fd

    @tensor
    def spatial_input(self) -> torch.Tensor:
        return self._fed["maps"]

    def _conv1x1(self, x: torch.Tensor, scope: str, act: Optional[str]) -> torch.Tensor:
        kernel = self.var(scope + "/kernel")
        return ops.linear(x, kernel.reshape(kernel.shape[2], kernel.shape[3]), self.var(scope + "/bias"), act)

    @tensor
    def spatial_states(self) -> torch.Tensor:
        hidden, proj = self._conv_scopes()
        x = self.spatial_input
        if hidden:
            x = self._conv1x1(x, hidden, "relu")
        if proj:
            x = self._conv1x1(x, proj, None)
        return x

    @tensor
    def spatial_mask(self) -> torch.Tensor:
        s = self.spatial_states
        return torch.ones(s.shape[:3], device=s.device, dtype=torch.float32)

    @tensor
    def output(self) -> torch.Tensor:
        return self.spatial_states.mean(dim=(1, 2))

    @property
    def dimension(self) -> int:
        return self.projection_dim or self.input_shape[2]
