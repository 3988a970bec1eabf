"""`vipformer.preproc` mirror.

The north star names "vipformer/preproc point-patch sampling and grouping
functions"; in the reference those live in vipformer/model/pointcloud/utils.py
(vipformer/preproc holds only the unused ImagePreprocessor,
vipformer/preproc/image.py:4-12).  Both are exported here.
"""
import torch

from ..model.pointcloud.utils import (divide_patches, farthest_point_sample, fps, index_points,  # noqa: F401
                                      knn_point, square_distance)


class ImagePreprocessor:
    """vipformer/preproc/image.py:4-12."""

    def __init__(self, transform):
        self.transform = transform

    def preprocess(self, img):
        return self.transform(img)

    def preprocess_batch(self, img_batch):
        return torch.stack([self.preprocess(img) for img in img_batch])
