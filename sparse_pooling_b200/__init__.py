"""sparse_pooling_b200 -- the Sparse Non-homogeneous Pooling Layer (SHPL) of
YeungLy/Sparse_Pooling, rebuilt for B200 (sm_100a).

Layout (only what the hot path needs):
  csrc/                CUDA kernels + the C ABI of include/shpl.h  -> libshpl.so
  _cabi.py             ctypes binding (no fallback: import fails without the library)
  ops.py               CSR plan, workspace, autograd op around the pooling kernels
  sparse_pool_utils.py drop-in mirror of the reference module (same names / signatures)
  builder.py           batched device-resident correspondence builder
  bev_slices.py        drop-in mirror of the BEV slicing feeder (BevSlices.generate_bev)
  construct_voxel.py   drop-in mirror of the MV3D voxel feeder (point_cloud_2_top_sparse)
  group_pointcloud.py  the VFE scatter_nd into the dense voxel grid (FeatureNet's sparse half) + build_input
  lidar_ingest.py      drop-in mirror of the point-cloud ingest (get_lidar_point_cloud)
  augment.py           augmentation hooks on the correspondence arrays (avod flip, MV3D augment_fv / augment_voxel points)
  config.py            the model.proto / kitti_dataset.proto sparse-pooling switches
  torch_op.py          the pooling kernels as registered PyTorch custom ops (torch.ops.shpl.pool) with autograd
"""
from . import _cabi  # noqa: F401  (raises ImportError when libshpl.so is missing)
from .config import (KittiDatasetSparsePoolingConfig, RetinaNetSparsePoolingConfig,  # noqa: F401
                     RpnSparsePoolingConfig)
from .ops import SparsePoolFunction, SparsePoolPlan, sparse_pool  # noqa: F401
from .sparse_pool_utils import (SparsePoolLayer, SparseTensor, _sparse_pool_op, _sparse_pool_trans_op,  # noqa: F401
                                concat_bn_op, gen_sparse_pooling_input_avod, produce_sparse_pooling_input,
                                sparse_pool_layer)
from .builder import build_avod_plan, build_pairs_plan  # noqa: F401
from .bev_slices import BevSlices  # noqa: F401
from . import construct_voxel  # noqa: F401
from . import group_pointcloud  # noqa: F401
from . import lidar_ingest  # noqa: F401
from . import augment  # noqa: F401
from . import torch_op  # noqa: F401  (registers torch.ops.shpl.pool / pool_backward)
