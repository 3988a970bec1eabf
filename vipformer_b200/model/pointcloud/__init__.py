"""Mirror of vipformer/model/pointcloud/__init__.py."""
from .classifier import PointCloudInputAdapter  # noqa: F401
from .partseg import (CrossFormer_img_mp, CrossFormer_partseg, CrossFormer_pc_mp, CrossFormer_pc_mp_ft,  # noqa: F401
                      load_pretrained)
