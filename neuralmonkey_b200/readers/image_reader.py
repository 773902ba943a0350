"""Image readers (reference: neuralmonkey/readers/image_reader.py:102-178).  Host-side
data format code next to the path (SURVEY.md section 8, row f): a list file names one image
per line; `.npy` arrays are taken as they are, anything else is opened with PIL."""
import os
from typing import Callable, Iterable, List

import numpy as np

VGG_RGB_MEANS = [[[123.68, 116.779, 103.939]]]  # image_reader.py:99


def _center_crop(image: np.ndarray, width: int, height: int) -> np.ndarray:
    """_crop (:206-216): symmetric crop, the odd pixel is dropped on the right / bottom."""
    h, w = image.shape[:2]
    w_shift, h_shift = max(w - width, 0) // 2, max(h - height, 0) // 2
    even_w, even_h = max(w - width, 0) % 2, max(h - height, 0) % 2
    return image[h_shift:h - h_shift - even_h, w_shift:w - w_shift - even_w]


def single_image_for_imagenet(path: str, target_height: int, target_width: int,
                              vgg_normalization: bool, zero_one_normalization: bool) -> np.ndarray:
    """The reference computes a rescaled image and discards it (:153-165: the result of
    `_rescale_or_crop` is never assigned), so what reaches the network is the centre crop of
    the ORIGINAL image, zero-padded to the target size.  Same here."""
    if path.endswith(".npy"):
        image = np.load(path)
    else:
        from PIL import Image
        image = np.array(Image.open(path).convert("RGB"))
    cropped = _center_crop(image, target_width, target_height)
    res = np.zeros((target_height, target_width, 3))
    res[:cropped.shape[0], :cropped.shape[1], :] = cropped
    if vgg_normalization:
        res -= VGG_RGB_MEANS
    if zero_one_normalization:
        res /= 255.
    return res


def imagenet_reader(prefix: str, target_width: int = 227, target_height: int = 227,
                    vgg_normalization: bool = False,
                    zero_one_normalization: bool = False) -> Callable:
    def load(list_files: List[str]) -> Iterable[np.ndarray]:
        for list_file in list_files:
            with open(list_file) as f_list:
                for i, image_file in enumerate(f_list):
                    path = os.path.join(prefix, image_file.rstrip())
                    if not os.path.exists(path):
                        raise Exception("Image file '{}' no. {} does not exist.".format(path, i + 1))
                    yield single_image_for_imagenet(path, target_height, target_width,
                                                    vgg_normalization, zero_one_normalization)
    return load
