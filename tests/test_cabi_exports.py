"""CPU-side checks of the drop-in boundary: libshpl.so loads, exports exactly the
symbols include/shpl.h declares, and rejects bad arguments before touching a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "shpl.h")
LIB = os.path.join(ROOT, "sparse_pooling_b200", "libshpl.so")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shpl_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_expected_entry_points():
    assert declared_functions() == sorted([
        "shpl_abi_version", "shpl_last_error", "shpl_kernel_launches", "shpl_build_workspace_bytes", "shpl_gen_input_avod",
        "shpl_debug_checks_enabled", "shpl_debug_check_failures",
        "shpl_produce_input", "shpl_build_avod", "shpl_plan_from_coo", "shpl_plan_from_voxel_coords", "shpl_pool_forward", "shpl_pool_backward",
        "shpl_pool_forward_dual", "shpl_pool_backward_dual", "shpl_pool_heavy",
        "shpl_bev_grid_dims", "shpl_bev_workspace_bytes", "shpl_bev_slices",
        "shpl_mv3d_workspace_bytes", "shpl_mv3d_voxelize", "shpl_lidar_workspace_bytes", "shpl_lidar_to_cam",
        "shpl_flip_point_cloud", "shpl_mv3d_project_augment", "shpl_augment_fv_index",
        "shpl_conv3x3_workspace_bytes", "shpl_pool_conv3x3_forward",
        "shpl_pool_forward_into", "shpl_pool_forward_into_dual", "shpl_pool_backward_from",
        "shpl_pool_heavy_workspace_bytes", "shpl_pool_heavy_split",
        "shpl_conv3x3_backward_workspace_bytes", "shpl_pool_conv3x3_backward"])


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), "libshpl.so does not export %s" % name


def test_binding_covers_the_header(lib):
    from sparse_pooling_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == declared_functions()
    assert _cabi.lib.shpl_abi_version() == _cabi.ABI_VERSION == 10


def test_workspace_query_grows_with_n(lib):
    lib.shpl_build_workspace_bytes.restype = ctypes.c_size_t
    lib.shpl_build_workspace_bytes.argtypes = [ctypes.c_int64]
    a, b = lib.shpl_build_workspace_bytes(1000), lib.shpl_build_workspace_bytes(1000000)
    assert 0 < a < b and b >= 1000000 * (4 * 8 + 12)


def test_invalid_arguments_return_error_codes_without_a_gpu(lib):
    from sparse_pooling_b200 import _cabi
    L = _cabi.lib
    assert L.shpl_pool_forward(None, None, None, None, None, None, 0, 0, 10, 4, 10, 0, None, None) == _cabi.SHPL_ERR_INVALID_ARGUMENT
    assert b"bad sizes" in L.shpl_last_error()
    assert L.shpl_pool_backward(None, None, None, None, None, 0, 0, 10, 4, 10, 4, None, None, None) == _cabi.SHPL_ERR_INVALID_ARGUMENT
    with pytest.raises(ValueError):
        _cabi.check(L.shpl_pool_forward(None, None, None, None, None, None, 0, 0, -1, 4, 10, 4, None, None), "shpl_pool_forward")
    st = _cabi.ShplPlan()
    rc = L.shpl_produce_input(None, None, None, 0, 8, 8, 8, 8, 0, 1, None, 0, 0, None, None, None, None,
                              ctypes.byref(st), 0, 0, None, None, 0, None)
    assert rc == _cabi.SHPL_ERR_INVALID_ARGUMENT and b"stride" in L.shpl_last_error()


def test_feeder_host_side_queries_and_argument_checks(lib):
    """The feeders' host-side entry points run without a GPU: grid geometry (voxel_grid_2d.py:126-149), workspace
    sizes, and the argument checks in front of any launch."""
    import numpy as np
    from sparse_pooling_b200 import _cabi
    L = _cabi.lib
    ext = np.array([-40.0, 40.0, -5.0, 3.0, 0.0, 70.0])
    nx, nz = ctypes.c_int32(0), ctypes.c_int32(0)
    assert L.shpl_bev_grid_dims(ext.ctypes.data_as(ctypes.c_void_p), 0.1, ctypes.byref(nx), ctypes.byref(nz)) == 0
    assert (nx.value, nz.value) == (800, 700)          # voxel_grid_2d_test.py:38-59: num_divisions == [800, 1, 700]
    w5 = L.shpl_bev_workspace_bytes(ext.ctypes.data_as(ctypes.c_void_p), 0.1, 5)
    w2 = L.shpl_bev_workspace_bytes(ext.ctypes.data_as(ctypes.c_void_p), 0.1, 2)
    assert w5 > w2 > 560000 * 8
    assert L.shpl_bev_workspace_bytes(ext.ctypes.data_as(ctypes.c_void_p), 0.1, 9) == 0          # > SHPL_BEV_MAX_SLICES
    assert L.shpl_bev_workspace_bytes(ext.ctypes.data_as(ctypes.c_void_p), -1.0, 5) == 0
    assert L.shpl_mv3d_workspace_bytes(1000) < L.shpl_mv3d_workspace_bytes(100000)
    assert L.shpl_lidar_workspace_bytes(1000) < L.shpl_lidar_workspace_bytes(1000000)
    gp = np.array([0.0, -1.0, 0.0, 1.65])
    rc = L.shpl_bev_slices(None, 1, 1, 10, None, gp.ctypes.data_as(ctypes.c_void_p), ext.ctypes.data_as(ctypes.c_void_p), 0.1,
                           -0.2, 2.3, 5, float(np.log(16)), None, 0, None, None, 0, None, None, None, 0, None)
    assert rc == _cabi.SHPL_ERR_INVALID_ARGUMENT and b"null pointer" in L.shpl_last_error()
    rc = L.shpl_mv3d_voxelize(None, None, 5, 0.2, 0.4, None, 45, None, None, None, None, 0, None, None, None, 0, None, None, 0, None)
    assert rc == _cabi.SHPL_ERR_INVALID_ARGUMENT
    rc = L.shpl_lidar_to_cam(None, 5, None, None, 0, 0, 0, 0.0, None, 0, None, None, 0, None)
    assert rc == _cabi.SHPL_ERR_INVALID_ARGUMENT
    # the VFE scatter plan and the augmentation hooks: argument checks in front of any launch
    st = _cabi.ShplPlan()
    st.n_rows, st.n_src = 10 * 20 * 24, 100
    rc = L.shpl_plan_from_voxel_coords(None, 1, 5, None, 1, 10, 20, 24, ctypes.byref(st), None, 0, None)
    assert rc == _cabi.SHPL_ERR_INVALID_ARGUMENT and b"null pointer" in L.shpl_last_error()
    rc = L.shpl_plan_from_voxel_coords(None, 1, 0, None, 1, 10, 20, 25, ctypes.byref(st), None, 0, None)
    assert rc == _cabi.SHPL_ERR_INVALID_ARGUMENT and b"grid" in L.shpl_last_error()
    assert L.shpl_flip_point_cloud(None, 1, 0, None, None) == 0                       # nothing to do, nothing launched
    assert L.shpl_flip_point_cloud(None, 1, 5, None, None) == _cabi.SHPL_ERR_INVALID_ARGUMENT
    assert L.shpl_flip_point_cloud(None, 0, 5, None, None) == _cabi.SHPL_ERR_INVALID_ARGUMENT
    assert L.shpl_mv3d_project_augment(None, 5, None, None, 0, 0.0, 0.0, 1.0, None, None, None) == _cabi.SHPL_ERR_INVALID_ARGUMENT
    assert L.shpl_mv3d_project_augment(None, 0, None, None, 0, 0.0, 0.0, 1.0, None, None, None) == 0
    assert L.shpl_augment_fv_index(None, 3, 5, None, 1.0, 0.0, 0.0, None) == _cabi.SHPL_ERR_INVALID_ARGUMENT    # ld < n
    assert L.shpl_augment_fv_index(None, 8, 0, None, 1.0, 0.0, 0.0, None) == 0


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under sparse_pooling_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sparse_pooling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libshpl_oracle" not in src, f


def test_ops_refuse_cpu_tensors():
    import torch
    from sparse_pooling_b200 import ops
    with pytest.raises(RuntimeError):
        ops.require_cuda(torch.zeros(2), "x")


def test_config_dataclasses_mirror_the_proto_fields(golden_dir):
    """Names, field numbers and DEFAULTS of the sparse-pooling switches come from tests/golden/proto_fields.json, which
    oracle/gen_goldens.py parsed out of the reference's .proto text (model.proto:87-91, :120-121; kitti_dataset.proto:38-40)."""
    import json
    import numpy as np
    from sparse_pooling_b200 import (KittiDatasetSparsePoolingConfig, RetinaNetSparsePoolingConfig, RpnSparsePoolingConfig,
                                     config)
    fields = json.load(open(os.path.join(golden_dir, "proto_fields.json")))
    assert len(fields) == 9
    by_message = {"RpnConfig": RpnSparsePoolingConfig(), "RetinaNetConfig": RetinaNetSparsePoolingConfig(),
                  "KittiDatasetConfig": KittiDatasetSparsePoolingConfig()}
    for f in fields:
        obj = by_message[f["message"]]
        assert hasattr(obj, f["name"]), f
        assert getattr(obj, f["name"]) == f["default"], (f, getattr(obj, f["name"]))
        assert f["label"] == "optional"
        key = ("dataset." + f["name"]) if (f["message"] == "KittiDatasetConfig" and f["name"] == "use_pyramid_level_at_SHPL") else f["name"]
        assert config.PROTO_FIELDS[key] == (f["message"], f["number"]), f
    assert len(config.PROTO_FIELDS) == len(fields)
    # the two defaults round 1 had wrong
    assert RpnSparsePoolingConfig().rpn_sparse_pooling_conv_after_fusion is True            # model.proto:89
    assert KittiDatasetSparsePoolingConfig().use_pyramid_level_at_SHPL == "P2"              # kitti_dataset.proto:40
    assert KittiDatasetSparsePoolingConfig().feat_stride() == 4                             # kitti_dataset.py:375
    c = RpnSparsePoolingConfig()
    assert c.bv_index_indicator() is None
    c.rpn_dual_sparse_pooling_after_vgg = True
    assert np.array_equal(c.bv_index_indicator(), np.zeros((1, 3)))                            # rpn_model.py:295
    assert KittiDatasetSparsePoolingConfig(use_pyramid_level_at_SHPL="P3").feat_stride() == 8    # kitti_dataset.py:375
    assert KittiDatasetSparsePoolingConfig(use_pyramid_level_at_SHPL="P0").feat_stride() == 1


def test_build_input_pads_like_the_reference_and_repairs_only_on_request():
    """group_pointcloud.py:88-105 ALWAYS pads a new first column (np.pad): with this fork's [K,4] coordinate buffers
    that gives five columns.  The default reproduces it; fix_coordinate_columns=True is the explicit repair."""
    import numpy as np
    from sparse_pooling_b200 import group_pointcloud as gp
    rng = np.random.default_rng(0)
    dicts = []
    for k in (5, 3):
        c4 = np.concatenate([np.zeros((k, 1), np.int64), rng.integers(0, 10, (k, 3))], axis=1)
        dicts.append(dict(feature_buffer=rng.standard_normal((k, 45, 7)), number_buffer=rng.integers(1, 45, k), coordinate_buffer=c4))
    B, feat, num, coord = gp.build_input(dicts)
    ref = np.concatenate([np.pad(d["coordinate_buffer"], ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, d in enumerate(dicts)])
    assert B == 2 and coord.shape == (8, 5) and np.array_equal(coord, ref)
    assert np.array_equal(feat, np.concatenate([d["feature_buffer"] for d in dicts])) and num.shape == (8,)
    _, _, _, fixed = gp.build_input(dicts, fix_coordinate_columns=True)
    assert fixed.shape == (8, 4) and np.array_equal(fixed[:, 0], [0] * 5 + [1] * 3)
    assert np.array_equal(fixed[:, 1:], np.concatenate([d["coordinate_buffer"][:, 1:] for d in dicts]))
    # [K,3] buffers (what the pad was written for): both modes agree
    d3 = [dict(d, coordinate_buffer=d["coordinate_buffer"][:, 1:]) for d in dicts]
    assert np.array_equal(gp.build_input(d3)[3], gp.build_input(d3, fix_coordinate_columns=True)[3])
    assert np.array_equal(gp.build_input(d3)[3], fixed)


def test_header_is_plain_c_and_links_from_c(tmp_path, lib):
    """include/shpl.h is the boundary a non-Python host binds: it must compile as C99 and as C++ with warnings as
    errors, and a C program linked against libshpl.so must see the declared ABI version (no GPU needed for that call)."""
    import shutil
    import subprocess
    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    if not gcc or not gxx:
        pytest.skip("no host compiler")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "shpl.h"\n'
                   'int main(void) {\n'
                   '    shpl_plan p;\n'
                   '    (void)p;\n'
                   '    printf("%d %d %d %d\\n", shpl_abi_version(), SHPL_ABI_VERSION, SHPL_HEAVY_LEN, (int)sizeof(shpl_plan));\n'
                   '    return shpl_pool_forward(0, 0, 0, 0, 0, 0, 0, 0, -1, 4, 10, 4, 0, 0) == SHPL_ERR_INVALID_ARGUMENT ? 0 : 1;\n'
                   '}\n')
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(LIB)
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, str(src), "-o", str(exe),
                    "-L", libdir, "-lshpl", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == out[1] == "10" and out[2] == "512"
    from sparse_pooling_b200 import _cabi
    assert int(out[3]) == ctypes.sizeof(_cabi.ShplPlan)                   # the ctypes mirror of struct shpl_plan has its layout
    cxx = tmp_path / "abi.cpp"
    cxx.write_text('#include "shpl.h"\nint main() { return shpl_abi_version() == SHPL_ABI_VERSION ? 0 : 1; }\n')
    subprocess.run([gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(cxx)], check=True)


def test_cpp_host_example_builds_against_the_library(lib):
    """examples/cabi_step.cu -- a C++ host on the bare C ABI, no Python -- compiles and links (running it needs a GPU)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-B", "NVCC=" + nvcc], check=True, capture_output=True)
    assert os.path.exists(os.path.join(ROOT, "examples", "cabi_step"))
