"""Pre-trained ImageNet encoder (reference: neuralmonkey/encoders/imagenet_encoder.py:16-258).

The reference imports the network definition from a checkout of tensorflow/models
(`research/slim/nets`, a third-party dependency that is not vendored) and runs it frozen
(`tf.stop_gradient`, :212,234).  What the hot path uses of it is the VGG convolution stack up
to a convolutional endpoint (tests/captioning.ini: `vgg_16/conv5/conv5_3`, [B,14,14,512]),
restated here from the published slim `nets/vgg.py`: blocks of 3x3/SAME conv + bias + ReLU
with 64-128-256-512-512 channels (2,2,3,3,3 convs for VGG-16; 2,2,4,4,4 for VGG-19), each
followed by a 2x2/2 max pool; variables `<net>/convB/convB_I/{weights,biases}` in HWIO
layout, which is the checkpoint layout.  The fully connected endpoints (fc6-fc8), AlexNet
and ResNet are outside the path.
"""
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from neuralmonkey_b200 import ops, runtime
from neuralmonkey_b200.decorators import tensor
from neuralmonkey_b200.model.model_part import ModelPart
from neuralmonkey_b200.model.parameterized import InitializerSpecs
from neuralmonkey_b200.model.stateful import SpatialStatefulWithOutput
from neuralmonkey_b200.params import variance_scaling_initializer, zeros_initializer

VGG_BLOCKS = {"vgg_16": (2, 2, 3, 3, 3), "vgg_19": (2, 2, 4, 4, 4)}
VGG_CHANNELS = (64, 128, 256, 512, 512)
SUPPORTED_NETWORKS = ["alexnet_v2", "vgg_16", "vgg_19", "resnet_v2_50", "resnet_v2_101",
                      "resnet_v2_152"]


def vgg_layers(network_type: str) -> List[Tuple[str, str, int, int]]:
    """[(endpoint, kind, cin, cout)] in execution order."""
    layers = []
    cin = 3
    for block, (convs, cout) in enumerate(zip(VGG_BLOCKS[network_type], VGG_CHANNELS), 1):
        for i in range(1, convs + 1):
            layers.append(("{}/conv{}/conv{}_{}".format(network_type, block, block, i), "conv", cin, cout))
            cin = cout
        layers.append(("{}/pool{}".format(network_type, block), "pool", cout, cout))
    return layers


class ImageNet(ModelPart, SpatialStatefulWithOutput):
    # pylint: disable=too-many-arguments
    def __init__(self, name: str, data_id: str, network_type: str, slim_models_path: str = None,
                 load_checkpoint: str = None, spatial_layer: str = None, encoded_layer: str = None,
                 initializers: InitializerSpecs = None) -> None:
        ModelPart.__init__(self, name, load_checkpoint=load_checkpoint, initializers=initializers,
                           save_checkpoint=None)
        self.data_id = data_id
        self.network_type = network_type
        self.spatial_layer = spatial_layer
        self.encoded_layer = encoded_layer
        if self.network_type not in SUPPORTED_NETWORKS:
            raise ValueError("Network '{}' is not among the supported ones ({})".format(
                self.network_type, ", ".join(SUPPORTED_NETWORKS)))
        if self.network_type not in VGG_BLOCKS:
            raise NotImplementedError("Only the VGG convolution stacks are built (got '{}')".format(
                self.network_type))
        self.height, self.width = 224, 224
        self._layers = vgg_layers(network_type)
        endpoints = [l[0] for l in self._layers]
        if self.spatial_layer is not None and self.spatial_layer not in endpoints:
            raise ValueError("Network '{}' does not contain endpoint '{}'".format(
                self.network_type, self.spatial_layer))
        if self.encoded_layer is not None:
            raise NotImplementedError("encoded_layer endpoints (fc6-fc8) are outside the hot path; "
                                      "leave it unset to average the convolutional maps")
        self._images = None  # type: Optional[torch.Tensor]

    def declare_variables(self) -> None:
        # slim networks live in their own top-level scope, independently of `name` (:113-114);
        # frozen: not trainable, so they sit behind the trainable prefix of the arena
        for endpoint, kind, cin, cout in self._layers:
            if kind != "conv":
                continue
            self.declare(endpoint + "/weights", [3, 3, cin, cout],
                         variance_scaling_initializer(2.0, "fan_in", "normal"), trainable=False,
                         absolute=True)
            self.declare(endpoint + "/biases", [cout], zeros_initializer(), trainable=False,
                         absolute=True)
            if endpoint == self.spatial_layer:
                break

    @property
    def input_types(self) -> Dict[str, Any]:
        return {self.data_id: float}

    @property
    def input_shapes(self) -> Dict[str, Any]:
        return {self.data_id: [None, self.height, self.width, 3]}

    def feed_dict(self, dataset, train: bool = False) -> Dict[str, Any]:
        fd = ModelPart.feed_dict(self, dataset, train)
        images = np.array(dataset.get_series(self.data_id), dtype=np.float32)
        if images.shape[1:] != (self.height, self.width, 3):
            raise ValueError("ImageNet '{}' expects images of shape {}, got {}".format(
                self.name, (self.height, self.width, 3), images.shape[1:]))
        self.feed_images(torch.from_numpy(images), train)
        fd[self.data_id] = images
        return fd

    def feed_images(self, images: torch.Tensor, train: bool = False) -> None:
        """[batch, H, W, 3] float32 (any H, W divisible by the pooling the endpoint needs)."""
        self.reset_batch()
        self.train_mode = bool(train)
        self.batch_size = int(images.shape[0])
        self._images = runtime.to_device(images)

    def static_inputs(self) -> Dict[str, Any]:
        return {"images": self._images} if self._images is not None else {}

    def bind_static(self, tensors: Dict[str, Any]) -> None:
        self.reset_batch()
        if tensors:
            self._images = tensors["images"]

    @tensor
    def input_image(self) -> torch.Tensor:
        return self._images

    @tensor
    def end_points(self) -> Dict[str, torch.Tensor]:
        points = {}
        x = self.input_image
        for endpoint, kind, _cin, _cout in self._layers:
            if kind == "conv":
                x = ops.conv3x3_bias_relu(x, self.var(endpoint + "/weights", absolute=True),
                                          self.var(endpoint + "/biases", absolute=True))
            else:
                x = ops.maxpool2x2(x)
            points[endpoint] = x
            if endpoint == self.spatial_layer:
                break
        return points

    @tensor
    def spatial_states(self) -> Optional[torch.Tensor]:
        if self.spatial_layer is None:
            return None
        return self.end_points[self.spatial_layer].detach()

    @tensor
    def spatial_mask(self) -> Optional[torch.Tensor]:
        if self.spatial_layer is None:
            return None
        s = self.spatial_states
        return torch.ones(s.shape[:3], device=s.device, dtype=torch.float32)

    @tensor
    def output(self) -> torch.Tensor:
        return self.spatial_states.mean(dim=(1, 2))

    @property
    def dimension(self) -> int:
        for endpoint, _kind, _cin, cout in self._layers:
            if endpoint == self.spatial_layer:
                return cout
        raise ValueError("spatial_layer is not set")
