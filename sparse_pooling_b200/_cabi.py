"""ctypes binding of libshpl.so -- the C ABI declared in include/shpl.h.

The library is built in-tree by ``sparse_pooling_b200/csrc/Makefile`` (nvcc,
sm_100a only).  There is NO fallback: if the shared object is missing or a symbol
cannot be resolved, importing this module raises, and every op of the package
fails loudly rather than computing on the CPU.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHPL_LIB") or os.path.join(_HERE, "libshpl.so")   # SHPL_LIB: try an experimental build

c_void_p = ctypes.c_void_p
c_int32 = ctypes.c_int32
c_int64 = ctypes.c_int64
c_size_t = ctypes.c_size_t

SHPL_OK = 0
SHPL_ERR_INVALID_ARGUMENT = -1
SHPL_ERR_CUDA = -2
SHPL_ERR_WORKSPACE_TOO_SMALL = -3
SHPL_ERR_UNSUPPORTED = -4

ABI_VERSION = 10
HEAVY_LEN = 512           # SHPL_HEAVY_LEN of include/shpl.h
EXACT_LEN = 2048          # SHPL_EXACT_LEN: listed cells up to this many entries keep the sequential order


class ShplPlan(ctypes.Structure):
    """struct shpl_plan of include/shpl.h (device pointers are plain integers)."""
    _fields_ = [
        ("n_rows", c_int32),
        ("n_src", c_int32),
        ("capacity", c_int32),
        ("row_ptr", c_void_p),
        ("csr_row", c_void_p),
        ("csr_src", c_void_p),
        ("csr_val", c_void_p),
        ("pix_ptr", c_void_p),
        ("csrT_pix", c_void_p),
        ("csrT_dst", c_void_p),
        ("csrT_val", c_void_p),
        ("heavy_cap", c_int32),
        ("heavy_row", c_void_p),
        ("heavy_pix", c_void_p),
        ("heavy_count", c_void_p),
        ("counts", c_void_p),
    ]


# name -> (restype, argtypes); exactly the functions include/shpl.h declares
SIGNATURES = {
    "shpl_abi_version": (ctypes.c_int, []),
    "shpl_last_error": (ctypes.c_char_p, []),
    "shpl_kernel_launches": (ctypes.c_uint64, []),
    "shpl_debug_checks_enabled": (ctypes.c_int, []),
    "shpl_debug_check_failures": (c_int64, []),
    "shpl_build_workspace_bytes": (c_size_t, [c_int64]),
    "shpl_gen_input_avod": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "shpl_produce_input": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64,
                                          c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                          c_int32, c_int32,
                                          c_void_p, c_void_p, c_void_p, c_void_p,
                                          ctypes.POINTER(ShplPlan), c_int32, c_int32, c_void_p,
                                          c_void_p, c_size_t, c_void_p]),
    "shpl_build_avod": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                       c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                       c_int32, c_int32,
                                       c_void_p, c_void_p, c_void_p, c_void_p,
                                       ctypes.POINTER(ShplPlan), c_int32, c_int32, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "shpl_plan_from_coo": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int64,
                                          c_int32, c_int32,
                                          ctypes.POINTER(ShplPlan), c_int32, c_int32, c_void_p,
                                          c_void_p, c_size_t, c_void_p]),
    "shpl_plan_from_voxel_coords": (ctypes.c_int, [c_void_p, c_int32, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                                   ctypes.POINTER(ShplPlan), c_void_p, c_size_t, c_void_p]),
    "shpl_pool_forward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "shpl_pool_backward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "shpl_pool_forward_dual": (ctypes.c_int, [c_void_p] * 10 + [c_int32] * 6 + [c_void_p] * 3),
    "shpl_pool_backward_dual": (ctypes.c_int, [c_void_p] * 10 + [c_int32] * 6 + [c_void_p] * 3),
    "shpl_bev_grid_dims": (ctypes.c_int, [c_void_p, ctypes.c_double, c_void_p, c_void_p]),
    "shpl_bev_workspace_bytes": (c_size_t, [c_void_p, ctypes.c_double, c_int32]),
    "shpl_bev_slices": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, c_int32, ctypes.c_double, c_void_p, c_int32,
                                       c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "shpl_mv3d_workspace_bytes": (c_size_t, [c_int64]),
    "shpl_mv3d_voxelize": (ctypes.c_int, [c_void_p, c_void_p, c_int64, ctypes.c_double, ctypes.c_double, c_void_p, c_int32,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                          c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "shpl_lidar_workspace_bytes": (c_size_t, [c_int64]),
    "shpl_lidar_to_cam": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, ctypes.c_float,
                                         c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "shpl_flip_point_cloud": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "shpl_mv3d_project_augment": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int32, ctypes.c_double, ctypes.c_double,
                                                 ctypes.c_double, c_void_p, c_void_p, c_void_p]),
    "shpl_augment_fv_index": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_void_p, ctypes.c_double, ctypes.c_double,
                                             ctypes.c_double, c_void_p]),
    "shpl_pool_heavy_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int32]),
    "shpl_pool_heavy_split": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_int32, c_void_p, c_int32, c_void_p, c_int32, c_int64, c_void_p, c_size_t, c_void_p]),
    "shpl_pool_forward_into": (ctypes.c_int, [c_void_p] * 5 + [c_int32] * 5 + [c_void_p, c_int32, c_int32, c_void_p]),
    "shpl_pool_forward_into_dual": (ctypes.c_int, [c_void_p] * 10 + [c_int32] * 6 + [c_void_p] * 3),
    "shpl_pool_backward_from": (ctypes.c_int, [c_void_p, c_int32, c_int32] + [c_void_p] * 4 + [c_int32] * 5 + [c_void_p, c_void_p]),
    "shpl_conv3x3_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "shpl_pool_conv3x3_forward": (ctypes.c_int, [c_void_p] * 6 + [c_int32] * 7 + [c_void_p, c_int32, c_void_p, c_void_p, c_int32,
                                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "shpl_conv3x3_backward_workspace_bytes": (c_size_t, [c_int32]),
    "shpl_pool_conv3x3_backward": (ctypes.c_int, [c_void_p] * 11 + [c_int32] * 7 + [c_void_p, c_int32] + [c_void_p] * 4 + [c_size_t, c_void_p]),
    "shpl_pool_heavy": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p]),
}


class ShplError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libshpl.so not found at %s -- build it with `make -C sparse_pooling_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.shpl_abi_version()
    if got != ABI_VERSION:
        raise ImportError("libshpl.so ABI version %d, binding expects %d" % (got, ABI_VERSION))
    return lib


lib = _load()


def check(rc, what):
    """Map the C status to the Python exceptions the reference's callers would see:
    invalid arguments -> ValueError (reference: assert / TF InvalidArgumentError)."""
    if rc == SHPL_OK:
        return
    msg = lib.shpl_last_error().decode("utf-8", "replace")
    if rc in (SHPL_ERR_INVALID_ARGUMENT, SHPL_ERR_UNSUPPORTED):
        raise ValueError("%s: %s" % (what, msg))
    raise ShplError("%s failed (%d): %s" % (what, rc, msg))
