"""Hand-derived exact-arithmetic known-answer test of the VALUE path (sparse_pool_layer forward and gradients).

The arithmetic of the reference's device half lives in TensorFlow, which cannot run here, and the reference holds no
vector for it: the value path stays "parity unpinned" against TF itself.  What CAN be pinned without TF
is TF's documented semantics on inputs where every summation order gives the same bits: integer-valued features and
power-of-two weights, so that every product and every partial sum is exactly representable.  The expected arrays below
were worked out BY HAND from sparse_pool_utils.py:61-117 (gather_nd -> SpMM -> concat; sparse_transpose + SpMM ->
scatter_nd with duplicates summed -> concat) and the gradients TF registers for those ops; they cover duplicate rows,
duplicate pixels, an empty row, both directions and both gradients.  The numpy oracle, the C oracle and the CUDA kernels
are all checked against these literals.

BEV 2x2 (row r = z*2 + x), image 2x3 (pixel p = v*3 + u), C_b = C_i = 2.
pairs k = 0..4: (row, pixel, weight) = (1,4,0.5) (1,4,2) (3,0,1) (1,2,0.25) (0,0,4)
"""
import numpy as np

BEV = np.array([[1, -1], [2, -2], [3, -3], [4, -4]], dtype=np.float32).reshape(1, 2, 2, 2)
IMG = np.array([[10, 0], [20, 1], [30, 2], [40, 3], [50, 4], [60, 5]], dtype=np.float32).reshape(1, 2, 3, 2)
ROWS = np.array([1, 1, 3, 1, 0])
PIX = np.array([4, 4, 0, 2, 0])
VAL = np.array([0.5, 2.0, 1.0, 0.25, 4.0], dtype=np.float32)
MIJ = np.stack([ROWS, np.arange(5)], axis=1).astype(np.int64)
FLIP = np.stack([np.zeros(5, np.int64), PIX // 3, PIX % 3], axis=1)          # rows [0, v, u]
M_SIZE = np.array([4, 5])

# forward, img -> bev:  Y[r] = sum_k val_k * img[pix_k]
#   Y[0] = 4*[10,0]                                   = [40, 0]
#   Y[1] = .5*[50,4] + 2*[50,4] + .25*[30,2]          = [132.5, 10.5]
#   Y[2] = 0 ; Y[3] = 1*[10,0]                        = [10, 0]
FUSED_BEV = np.array([[1, -1, 40, 0], [2, -2, 132.5, 10.5], [3, -3, 0, 0], [4, -4, 10, 0]], dtype=np.float32).reshape(1, 2, 2, 4)
# forward, bev -> img:  S[p] = sum_{k at p} val_k * bev[row_k]
#   p0: 1*[4,-4] + 4*[1,-1] = [8,-8] ; p2: .25*[2,-2] = [.5,-.5] ; p4: .5*[2,-2] + 2*[2,-2] = [5,-5]
FUSED_IMG = np.array([[10, 0, 8, -8], [20, 1, 0, 0], [30, 2, 0.5, -0.5], [40, 3, 0, 0], [50, 4, 5, -5], [60, 5, 0, 0]],
                     dtype=np.float32).reshape(1, 2, 3, 4)

# upstream gradients: g_fused_bev[r] = [r, r+10 | gY[r]] with gY = [1,-1],[2,-2],[4,-4],[8,-8];
#                     g_fused_img[p] = [p, -p | 1, 2]
G_FUSED_BEV = np.array([[0, 10, 1, -1], [1, 11, 2, -2], [2, 12, 4, -4], [3, 13, 8, -8]], dtype=np.float32).reshape(1, 2, 2, 4)
G_FUSED_IMG = np.array([[p, -p, 1, 2] for p in range(6)], dtype=np.float32).reshape(1, 2, 3, 4)

# single direction (bv_index = None):  g_bev = slice ; g_img[p] = sum_{k at p} val_k * gY[row_k]
#   p0: 1*[8,-8] + 4*[1,-1] = [12,-12] ; p2: .25*[2,-2] = [.5,-.5] ; p4: .5*[2,-2] + 2*[2,-2] = [5,-5]
G_BEV_SINGLE = np.array([[0, 10], [1, 11], [2, 12], [3, 13]], dtype=np.float32).reshape(1, 2, 2, 2)
G_IMG_SINGLE = np.array([[12, -12], [0, 0], [0.5, -0.5], [0, 0], [5, -5], [0, 0]], dtype=np.float32).reshape(1, 2, 3, 2)
# dual direction: each input feeds two consumers (AddN):
#   g_bev[r] = slice + (sum of the weights in row r) * [1, 2]:  row 0: 4, row 1: 2.75, row 3: 1
#   g_img[p] = [p, -p] + G_IMG_SINGLE[p]
G_BEV_DUAL = np.array([[4, 18], [3.75, 16.5], [2, 12], [4, 15]], dtype=np.float32).reshape(1, 2, 2, 2)
G_IMG_DUAL = np.array([[12, -12], [1, -1], [2.5, -2.5], [3, -3], [9, -9], [5, -5]], dtype=np.float32).reshape(1, 2, 3, 2)
