"""Seeded synthetic inputs (moved to tools/synth.py so that the GPU arm of bench.py and the tools never import
the oracle package); re-exported here for the tests and the golden generator, which refer to ``oracle.synth``."""
from tools.synth import *  # noqa: F401,F403
from tools.synth import P2_KITTI, AVOD_EXTENTS, AVOD_VOXEL, AVOD_BEV_HW, AVOD_IMG_WH  # noqa: F401
