"""The sparse-pooling switches of the reference's protobuf configs, as dataclasses
(`protoc` is not available here, and only these fields concern the SHPL path).

Field names, numbers and defaults follow (pinned by tests/golden/proto_fields.json, which oracle/gen_goldens.py
parses out of the reference's .proto text)
  /root/reference/avod/avod/protos/model.proto:86-91   (RpnConfig fields 6-10)
  /root/reference/avod/avod/protos/model.proto:120-121 (RetinaNetConfig)
  /root/reference/avod/avod/protos/kitti_dataset.proto:38-40
Meanings: /root/reference/avod/README.md:120-125.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class RpnSparsePoolingConfig:
    rpn_use_sparse_pooling: bool = False                # SHPL right before the RPN (rpn_model.py:328)
    rpn_sparse_pooling_use_batch_norm: bool = False     # concat_bn_op instead of tf.concat
    rpn_sparse_pooling_conv_after_fusion: bool = True   # 3x3 conv back to C (rpn_model.py:338-354): conv_fusion.py fuses it
    rpn_sparse_pooling_after_vgg: bool = False          # SHPL at conv4 inside FusionVggPyr (rpn_model.py:291)
    rpn_dual_sparse_pooling_after_vgg: bool = False     # also pool BEV -> image there (rpn_model.py:294-298)

    def bv_index_indicator(self):
        """rpn_model.py:294-298: the dual switch is a non-None dummy array."""
        return np.zeros((1, 3)) if self.rpn_dual_sparse_pooling_after_vgg else None


@dataclass
class RetinaNetSparsePoolingConfig:
    use_sparse_pooling: bool = False
    use_pyramid_level_at_SHPL: str = "P2"


@dataclass
class KittiDatasetSparsePoolingConfig:
    output_indices: bool = False                        # SHPL is silently off when false (rpn_model.py:111-118)
    use_pyramid_level_at_SHPL: str = "P2"               # kitti_dataset.proto:40 default: stride 4 (the SHPL configs set P0)

    def feat_stride(self):
        """kitti_dataset.py:375: 2 ** int(level[-1])."""
        return 2 ** int(self.use_pyramid_level_at_SHPL[-1])


# MV3D: /root/reference/MV3D_TF_release/lib/fast_rcnn/config.py:229 -- images are padded to [W, H] before the image
# network; augment_fv clips to it (minibatch_mv3d_img.py:200-203).
PAD_IMAGE_TO = [1280, 384]


# field name -> (message, field number) as the .proto files declare them; checked against the parsed fixture
PROTO_FIELDS = {
    "rpn_use_sparse_pooling": ("RpnConfig", 6),
    "rpn_sparse_pooling_use_batch_norm": ("RpnConfig", 7),
    "rpn_sparse_pooling_conv_after_fusion": ("RpnConfig", 8),
    "rpn_sparse_pooling_after_vgg": ("RpnConfig", 9),
    "rpn_dual_sparse_pooling_after_vgg": ("RpnConfig", 10),
    "use_sparse_pooling": ("RetinaNetConfig", 6),
    "use_pyramid_level_at_SHPL": ("RetinaNetConfig", 9),
    "output_indices": ("KittiDatasetConfig", 11),
    "dataset.use_pyramid_level_at_SHPL": ("KittiDatasetConfig", 12),
}
