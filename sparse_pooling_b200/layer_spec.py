"""Shapes and algorithmic-byte formulas of one SHPL instance.  Pure Python: no CUDA, no
libshpl.so -- bench.py's reference arm loads this file on its own so that the CPU arm never maps the product
library."""
from dataclasses import dataclass
from typing import Tuple


@dataclass
class LayerSpec:
    """One SHPL instance of a model."""
    name: str
    bev_hw: Tuple[int, int]      # BEV feature map H, W   (= floor(bv_size / stride_bv))
    img_hw: Tuple[int, int]      # image feature map H, W (= floor(im_size / stride_img))
    c_bev: int
    c_img: int
    stride: Tuple[int, int]      # (stride_img, stride_bv): produce_sparse_pooling_input's stride argument
    dual: bool                   # bev -> img as well (bv_index is not None)
    im_size: Tuple[int, int]     # (W, H) handed to gen_sparse_pooling_input_avod
    bv_size: Tuple[int, int]     # (H_b, W_b) handed to gen_sparse_pooling_input_avod

    @property
    def R(self):
        return self.bev_hw[0] * self.bev_hw[1]

    @property
    def Q(self):
        return self.img_hw[0] * self.img_hw[1]

    def bytes_forward(self, nnz):
        """Algorithmic bytes of the forward launches: the drop-in (concat) form."""
        b = 4 * (self.R * self.c_bev + self.R * (self.c_bev + self.c_img) + nnz * (self.c_img + 2) + self.R + 1)
        if self.dual:
            b += 4 * (self.Q * self.c_img + self.Q * (self.c_img + self.c_bev) + nnz * (self.c_bev + 2) + self.Q + 1)
        return b

    def bytes_backward(self, nnz):
        b = 4 * (2 * self.R * self.c_bev + nnz * (self.c_img + 2) + self.Q * self.c_img + self.Q + 1)
        if self.dual:
            b += 4 * (2 * self.Q * self.c_img + nnz * (self.c_bev + 2) + self.R * self.c_bev + self.R + 1)
        return b

    def bytes_forward_sparse_only(self, nnz):
        """The no-concat form of SURVEY.md 8(d): the producer of the destination map writes straight into the fused
        buffer, the op writes only the pooled channels (zeros for empty cells)."""
        b = 4 * (self.R * self.c_img + nnz * (self.c_img + 2) + self.R + 1)
        if self.dual:
            b += 4 * (self.Q * self.c_bev + nnz * (self.c_bev + 2) + self.Q + 1)
        return b

    def bytes_backward_sparse_only(self, nnz):
        """No-concat backward: g_dst is a view of g_fused (no slice copy)."""
        b = 4 * (nnz * (self.c_img + 2) + self.Q * self.c_img + self.Q + 1)
        if self.dual:
            b += 4 * (nnz * (self.c_bev + 2) + self.R * self.c_bev + self.R + 1)
        return b
