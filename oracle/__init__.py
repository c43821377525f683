"""CPU oracle for the SHPL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, on the CPU, the algorithm of the reference's Sparse
Non-homogeneous Pooling Layer (avod/avod/utils/sparse_pool_utils.py and its
MV3D twin) so that the CUDA path can be checked against it.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from here.  Nothing under
``sparse_pooling_b200/`` imports it; the product path raises if ``libshpl.so``
is missing rather than falling back to this code.

Pinning status
--------------
* index path (correspondence builder): PINNED -- checked against outputs of the
  reference's own numpy code (imported from /root/reference by
  ``oracle/gen_goldens.py``; fixtures under ``tests/golden/``) and against the
  known-answer vectors KAT-1 / KAT-2 of SURVEY.md Appendix B.
* feeders and ingest (BevSlices.generate_bev, MV3D point_cloud_2_top_sparse, get_lidar_point_cloud): PINNED --
  ``feeder_oracle.py`` is checked against outputs of the reference's own classes / functions run by
  ``gen_goldens.py`` on seeded synthetic scans (for the ingest: through KITTI-format files in a temporary directory).
* augmentation hooks (kitti_aug flips, MV3D augment_voxel's point transforms + img_index2, augment_fv's index
  update): PINNED -- ``gen_goldens.py augment_goldens`` imports kitti_aug.py and executes the source of augment_voxel /
  augment_fv where it lies (the file as a whole is Python 2) with np.random seeded.
* value path (gather -> SpMM -> concat, scatter, the VFE scatter_nd, gradients): the arithmetic lives
  in TensorFlow 1.8 (third-party, not vendored, not installable here) and the
  reference holds no golden vector, test or fixture for it:
  PARITY UNPINNED at the TF boundary.  The restatement follows TF's documented
  op semantics and is cross-checked against torch CPU autograd.
"""
