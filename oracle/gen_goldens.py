"""Generate tests/golden/*.npz by running the REFERENCE's own numpy code.

Run in the build container only (needs /root/reference; the GPU box has none):

    python -m oracle.gen_goldens

TensorFlow / matplotlib / easydict are absent here, so tiny stand-in modules are
put in sys.modules before the reference files are imported;
no reference source is copied -- the files are imported where they lie.

Every fixture stores the outputs the reference produced plus a sha256 of the
inputs, which tests regenerate from the same seed through oracle/synth.py.
"""
import hashlib
import os
import sys
import types

import numpy as np

from oracle import synth

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _install_stubs():
    tf = types.ModuleType("tensorflow")
    tf.contrib = types.SimpleNamespace(slim=types.SimpleNamespace())
    sys.modules.setdefault("tensorflow", tf)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    ed = types.ModuleType("easydict")

    class EasyDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    ed.EasyDict = EasyDict
    sys.modules.setdefault("easydict", ed)


def import_avod():
    _install_stubs()
    for p in (REF + "/avod", REF + "/avod/wavedata"):
        if p not in sys.path:
            sys.path.insert(0, p)
    import avod.utils.sparse_pool_utils as spu
    return spu


def import_mv3d_construct_voxel():
    _install_stubs()
    p = REF + "/MV3D_TF_release/lib"
    if p not in sys.path:
        sys.path.append(p)
    import importlib.util
    spec = importlib.util.spec_from_file_location("mv3d_config_voxels", p + "/utils/config_voxels.py")
    cfgmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfgmod)
    # construct_voxel does `from utils.config_voxels import cfg` and `from utils.transform import ...`
    pkg = types.ModuleType("utils")
    pkg.__path__ = [p + "/utils"]
    sys.modules["utils"] = pkg
    sys.modules["utils.config_voxels"] = cfgmod
    spec = importlib.util.spec_from_file_location("utils.construct_voxel", p + "/utils/construct_voxel.py")
    cv = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(cv)
    finally:
        pass
    return cv


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


class _Calib:
    def __init__(self, p2):
        self.p2 = p2


def ref_gen_and_produce(spu, frame, stride):
    d = spu.gen_sparse_pooling_input_avod(frame["points"], frame["voxel_indices"], _Calib(frame["P"]),
                                          frame["im_size"], frame["bv_size"])
    gen = {k: np.array(v, copy=True) for k, v in d.items()}
    out = spu.produce_sparse_pooling_input(d, stride=list(stride))
    return gen, out, d


def main():
    os.makedirs(OUT, exist_ok=True)
    spu = import_avod()

    # ---- KAT-1 / KAT-2 --------------------------------
    P = np.array([[1.0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]])
    uv = [(0.5, 0.5), (1.5, 1.5), (2.5, 2.5), (-0.0, 0.0), (-1e-9, 0.0), (8.999, 3.0), (9.0, 3.0), (8.4999, 4.999), (3.0, 5.0)]
    pts = np.array([[u, v, 1.0] for u, v in uv])
    vox = np.array([[k, z] for k, z in enumerate([1, 2, 3, 4, 5, 6, 7, 8, 0])])
    frame = dict(points=pts, voxel_indices=vox, P=P, im_size=[10, 6], bv_size=(8, 16))
    gen, out, mutated = ref_gen_and_produce(spu, frame, (2, 2))
    np.savez_compressed(os.path.join(OUT, "kat1.npz"), points=pts, voxel_indices=vox, P=P, im_size=[10, 6],
                        bv_size=[8, 16], stride=[2, 2], gen_bv_index=gen["bv_index"], gen_img_index=gen["img_index"],
                        Mij_pool=out["Mij_pool"], M_val=out["M_val"], M_size=out["M_size"],
                        img_index_flip_pool=out["img_index_flip_pool"], img_index_after=mutated["img_index"])
    d2 = dict(bv_index=np.array([[3, 8], [3, 7], [15, 7], [16, 6]]), img_index=np.array([[1.0, 2, 3, 4], [1, 1, 2, 2], [0, 0, 0, 0]]),
              bv_size=np.array([8, 16]), img_size=np.array([10, 6]))
    o2 = spu.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d2.items()}, stride=[1, 1])
    np.savez_compressed(os.path.join(OUT, "kat2.npz"), **{"in_" + k: v for k, v in d2.items()},
                        Mij_pool=o2["Mij_pool"], M_val=o2["M_val"], M_size=o2["M_size"],
                        img_index_flip_pool=o2["img_index_flip_pool"])

    # ---- avod frame through the reference builder at the three strides -------
    for seed, az in ((1, 0.09), (2, 0.05)):
        frame = synth.avod_frame(seed, az_step_deg=az)
        rec = dict(input_sha=digest(frame["points"], frame["voxel_indices"]), seed=seed, az=az,
                   n_in=frame["points"].shape[0])
        for s in (1, 4, 8):
            gen, out, _ = ref_gen_and_produce(spu, frame, (s, s))
            if s == 1:
                rec["gen_bv_index"] = gen["bv_index"]
                rec["gen_img_index"] = gen["img_index"].astype(np.int32)   # integer-valued f64, stored compactly
            rec["Mij_pool_s%d" % s] = out["Mij_pool"][:, 0].astype(np.int32)  # col is arange (checked below)
            assert (out["Mij_pool"][:, 1] == np.arange(out["Mij_pool"].shape[0])).all()
            assert out["Mij_pool"].dtype == np.int64 and out["img_index_flip_pool"].dtype == np.int64
            assert out["M_val"].dtype == np.float64 and (out["M_val"] == 1).all()
            rec["M_size_s%d" % s] = out["M_size"]
            rec["flip_s%d" % s] = out["img_index_flip_pool"].astype(np.int32)
        np.savez_compressed(os.path.join(OUT, "avod_frame_seed%d.npz" % seed), **rec)
        print("avod_frame seed", seed, "N", rec["n_in"], "nnz", {s: int(rec["M_size_s%d" % s][1]) for s in (1, 4, 8)})

    # ---- direct pairs (config 1 'direct', and skewed stress shapes) ----------
    for name, kw in (("direct_uniform", dict(seed=0, n=20000)),
                     ("direct_ground", dict(seed=3, n=50000, skew="ground")),
                     ("direct_zipf", dict(seed=4, n=30000, skew="zipf"))):
        d = synth.direct_pairs(**kw)
        rec = dict(input_sha=digest(d["bv_index"], d["img_index"]))
        for s in ((1, 1), (8, 8), (8, 2)):
            o = spu.produce_sparse_pooling_input({k: np.array(v, copy=True) for k, v in d.items()}, stride=list(s))
            tag = "s%d_%d" % s
            rec["row_" + tag] = o["Mij_pool"][:, 0].astype(np.int32)
            rec["M_size_" + tag] = o["M_size"]
            rec["flip_" + tag] = o["img_index_flip_pool"].astype(np.int32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, {k: v.tolist() for k, v in rec.items() if k.startswith("M_size")})

    # ---- MV3D feeder: reference point_cloud_2_top_sparse ---------------------
    mv3d_goldens(spu)



# ---- the sparse-pooling switches, parsed out of the reference's .proto TEXT (protoc is not available here) ----
PROTO_SOURCES = (("avod/avod/protos/model.proto", ("RpnConfig", "RetinaNetConfig")),
                 ("avod/avod/protos/kitti_dataset.proto", ("KittiDatasetConfig",)))
PROTO_WANTED = ("rpn_use_sparse_pooling", "rpn_sparse_pooling_use_batch_norm", "rpn_sparse_pooling_conv_after_fusion",
                "rpn_sparse_pooling_after_vgg", "rpn_dual_sparse_pooling_after_vgg", "use_sparse_pooling",
                "use_pyramid_level_at_SHPL", "output_indices")


def proto_goldens(ref_root="/root/reference"):
    """tests/golden/proto_fields.json: for every sparse-pooling field of model.proto / kitti_dataset.proto its message,
    label, type, field number, default and source line, read from the .proto text itself."""
    import json
    import re
    field_re = re.compile(r"^\s*(optional|required|repeated)\s+(\w+)\s+(\w+)\s*=\s*(\d+)\s*(?:\[\s*default\s*=\s*([^\]]+?)\s*\])?\s*;")
    out = []
    for rel, messages in PROTO_SOURCES:
        message = None
        for ln, line in enumerate(open(os.path.join(ref_root, rel)), 1):
            m = re.match(r"^\s*message\s+(\w+)", line)
            if m:
                message = m.group(1)
                continue
            f = field_re.match(line)
            if f and message in messages and f.group(3) in PROTO_WANTED:
                default = f.group(5)
                if default is not None:
                    default = default.strip("'\"")
                    if f.group(2) == "bool":
                        default = {"true": True, "false": False}[default]
                out.append(dict(file=rel, line=ln, message=message, label=f.group(1), type=f.group(2), name=f.group(3),
                                number=int(f.group(4)), default=default))
    with open(os.path.join(OUT, "proto_fields.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("proto_fields:", [(r["message"], r["name"], r["number"], r["default"]) for r in out])
    return out


MV3D_CASES = {
    # name: (seed, n_points, kwargs of synth.mv3d_frame)
    "mv3d_seed5": (5, 6000, {}),                                  # ped/cyc ranges (config_voxels.py:50-64), cap 45
    "mv3d_car_seed6": (6, 9000, dict(car=True)),                  # car ranges (config_voxels.py:33-48), cap 35
}


def mv3d_goldens(spu):
    cv = import_mv3d_construct_voxel()
    for name, (seed, n, kw) in MV3D_CASES.items():
        f = synth.mv3d_frame(seed=seed, n_points=n, **kw)
        cam4 = synth.mv3d_cam4(f)                                 # x, y, z, reflectance
        calib = np.zeros((4, 12))
        calib[0] = synth.P2_KITTI.reshape(-1)
        cv.MAX_NUM_POINTS = f["max_points"]                       # module global = cfg.VOXEL_POINT_COUNT (construct_voxel.py:17)
        vd, vfs, img_index, bv_index, M_val = cv.point_cloud_2_top_sparse(
            cam4.copy(), res=f["res"], zres=f["zres"], side_range=f["side_range"], fwd_range=f["fwd_range"],
            height_range=f["height_range"], points_in_cam=True, calib=calib, img_index2=f["img_index2"].copy())
        o = spu.produce_sparse_pooling_input(dict(img_index=np.array(img_index, dtype=np.float64), img_size=f["img_size"],
                                                  bv_index=bv_index, bv_size=[vfs[1], vfs[2]]), M_val=M_val, stride=[8, 2])
        fb = vd["feature_buffer"]
        assert fb.dtype == np.float64 and fb.shape[1:] == (f["max_points"], 7)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), input_sha=digest(f["points_fsh"], f["img_index2"]),
                            voxel_full_size=vfs, img_index=np.asarray(img_index).astype(np.int32), bv_index=bv_index.astype(np.int32),
                            M_val=M_val, row=o["Mij_pool"][:, 0].astype(np.int32), M_size=o["M_size"],
                            flip=o["img_index_flip_pool"].astype(np.int32),
                            coordinate_buffer=vd["coordinate_buffer"].astype(np.int32),
                            number_buffer=vd["number_buffer"].astype(np.int32),
                            feature_buffer_sha=digest(fb), feature_buffer_head=fb[:64])
        print(name, vfs, "voxels", fb.shape[0], "pairs", len(M_val), "M_size", o["M_size"], "min w", M_val.min())


def feeder_goldens():
    """BevSlices.generate_bev(output_indices=True) of the reference on the synthetic scan."""
    import_avod()
    from avod.core.bev_generators.bev_slices import BevSlices
    from wavedata.tools.obj_detection import obj_utils

    class KU:   # the two calls of KittiUtils.create_slice_filter (kitti_utils.py:97-107)
        def create_slice_filter(self, pc, ext, gp, lo, hi):
            return np.logical_xor(obj_utils.get_point_filter(pc, ext, gp, hi), obj_utils.get_point_filter(pc, ext, gp, lo))

    cfg = types.SimpleNamespace(height_lo=-0.2, height_hi=2.3, num_slices=5)
    for seed, az in ((1, 0.09), (2, 0.05), (3, None)):
        pts = synth.lidar_scan_gappy(seed) if az is None else synth.lidar_scan(seed, az_step_deg=az)
        gp = np.array([0.0, -1.0, 0.0, 1.65])
        maps, idx, upts = BevSlices(cfg, KU()).generate_bev("lidar", pts.T, gp, synth.AVOD_EXTENTS, synth.AVOD_VOXEL,
                                                            output_indices=True)
        rec = dict(input_sha=digest(pts), n_points=len(pts), voxel_indices=idx.astype(np.int32), unique_pts=upts)
        for i, hm in enumerate(maps["height_maps"]):
            nz = np.nonzero(hm)
            rec["hm%d_idx" % i] = np.stack(nz, axis=1).astype(np.int32)
            rec["hm%d_val" % i] = hm[nz]
            assert hm.shape == (700, 800) and hm.dtype == np.float64
        dm = maps["density_map"]
        nz = np.nonzero(dm)
        rec["dm_idx"] = np.stack(nz, axis=1).astype(np.int32)
        rec["dm_val"] = dm[nz]
        np.savez_compressed(os.path.join(OUT, "bev_slices_seed%d.npz" % seed), **rec)
        print("bev_slices seed", seed, "points", len(pts), "pairs", len(idx))
        if az is None:      # the fixture must exercise the re-use quirk
            n_in = [int(KU().create_slice_filter(pts.T, synth.AVOD_EXTENTS, gp, -0.2 + i * 0.5, -0.2 + (i + 1) * 0.5).sum())
                    for i in range(5)]
            assert n_in[2] == 1 and n_in[4] == 0 and min(n_in[0], n_in[1], n_in[3]) > 1, n_in


def ingest_goldens():
    """obj_utils.get_lidar_point_cloud of the reference on synthetic KITTI files (a calib .txt and a velodyne .bin
    written to a temporary directory in the format calib_utils.read_calibration / read_lidar parse)."""
    import tempfile
    import_avod()
    from wavedata.tools.obj_detection import obj_utils
    for seed, az, im_size in ((1, 0.4, [1242, 375]), (2, 0.15, [1242, 375])):
        scan = synth.velodyne_scan(seed, az_step_deg=az)
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(tmp + "/calib")
            os.makedirs(tmp + "/velodyne")
            with open(tmp + "/calib/%06d.txt" % seed, "w") as f:
                f.write(synth.kitti_calib_text())
            scan.tofile(tmp + "/velodyne/%06d.bin" % seed)
            pc = obj_utils.get_lidar_point_cloud(seed, tmp + "/calib", tmp + "/velodyne", im_size=im_size)
            pc_all = obj_utils.get_lidar_point_cloud(seed, tmp + "/calib", tmp + "/velodyne")
            from wavedata.tools.core import calib_utils
            cal = calib_utils.read_calibration(tmp + "/calib", seed)
        assert pc.dtype == np.float64 and pc.shape[0] == 3
        rec = dict(input_sha=digest(scan), n_in=len(scan), im_size=np.array(im_size), fov_sha=digest(pc), all_sha=digest(pc_all),
                   n_fov=pc.shape[1], p2=cal.p2, r0_rect=cal.r0_rect, tr_velodyne_to_cam=cal.tr_velodyne_to_cam)
        if seed == 1:
            rec["fov_points"] = pc
        np.savez_compressed(os.path.join(OUT, "lidar_ingest_seed%d.npz" % seed), **rec)
        print("lidar ingest seed", seed, "points", len(scan), "in FOV", pc.shape[1])


def _mv3d_minibatch_functions(names):
    """The reference's minibatch_mv3d_img.py is Python 2 as a whole (print statements), but augment_voxel and
    augment_fv themselves are version-neutral: their source is read where it lies and executed in a namespace that
    holds what they use (numpy, cv2, cfg, and calib_to_P / projectToImage from the reference's lib/utils/transform.py)."""
    import importlib.util
    import cv2
    lib = os.path.join(REF, "MV3D_TF_release", "lib")
    spec = importlib.util.spec_from_file_location("mv3d_transform", lib + "/utils/transform.py")
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    lines = open(lib + "/roi_data_layer/minibatch_mv3d_img.py").read().split("\n")
    cfg = types.SimpleNamespace(PAD_IMAGE_TO=[1280, 384], TRAIN=types.SimpleNamespace(AUGMENT_PC=False))
    # augment_fv hands cv2.resize a 1-element array as fx / fy, which the OpenCV of the reference's day accepted and
    # 4.13 rejects: the shim converts the two factors to float and changes nothing else
    cv2_shim = types.SimpleNamespace(resize=lambda img, dsize, fx, fy: cv2.resize(img, dsize, fx=float(np.ravel(fx)[0]),
                                                                              fy=float(np.ravel(fy)[0])),
                                     warpAffine=cv2.warpAffine)
    ns = dict(np=np, cv2=cv2_shim, cfg=cfg, calib_to_P=tr.calib_to_P, projectToImage=tr.projectToImage)
    for name in names:
        i = [k for k, l in enumerate(lines) if l.startswith("def " + name + "(")][0]
        j = i + 1
        while j < len(lines) and not lines[j].startswith("def "):
            j += 1
        exec(compile("\n".join(lines[i:j]), name, "exec"), ns)
    return [ns[n] for n in names]


def augment_goldens():
    """The augmentation hooks of the reference run here: kitti_aug's flips (imported) and MV3D's augment_voxel /
    augment_fv (executed from their source, see above) with np.random seeded."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("kitti_aug", os.path.join(REF, "avod/avod/datasets/kitti/kitti_aug.py"))
    ka = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ka)
    pts = synth.lidar_scan(1, az_step_deg=0.09).T                 # [3,N] camera frame
    gp = np.array([0.01, -1.0, 0.02, 1.65])
    rec = dict(flip_input_sha=digest(pts), flip_sha=digest(ka.flip_point_cloud(pts)), flip_head=ka.flip_point_cloud(pts)[:, :16],
               flip_plane=ka.flip_ground_plane(gp), flip_p2=ka.flip_stereo_calib_p2(synth.P2_KITTI, (375, 1242)))
    augment_voxel, augment_fv = _mv3d_minibatch_functions(["augment_voxel", "augment_fv"])
    f = synth.mv3d_frame(seed=7, n_points=5000)
    pc = synth.mv3d_cam4(f)
    calib = np.zeros((4, 12))
    calib[0] = synth.P2_KITTI.reshape(-1)
    np.random.seed(1234)
    blobs = dict(gt_boxes_3d=np.zeros((1, 7)), gt_rys=np.zeros((1, 1)))
    _, pc_aug, img_index2 = augment_voxel(blobs, scale=0.8, lidar_pc=pc.copy(), calib=calib)
    np.random.seed(1234)                                         # the draws augment_voxel made, in its order (:133-136)
    sx, sz = np.random.uniform(-0.8, 0.8, 2)
    ratio = np.random.uniform(0.95, 1.05, 1)
    angle = np.random.uniform(-np.pi / 10, np.pi / 10, 1)
    rec.update(voxel_input_sha=digest(pc), voxel_params=np.array([sx, sz, ratio[0], angle[0]]), voxel_pc_sha=digest(pc_aug),
               voxel_pc_head=pc_aug[:16], voxel_img_index2=img_index2.astype(np.int32))
    img_index = np.vstack((f["img_index2"], np.zeros((1, f["img_index2"].shape[1]), dtype=int)))
    np.random.seed(99)
    blobs = dict(image_data=np.zeros((375, 1242, 3), np.float32), gt_boxes=np.zeros((1, 5)), img_index=img_index.copy())
    blobs, shift, fv_ratio = augment_fv(blobs, scale=10)
    rec.update(fv_input_sha=digest(img_index), fv_params=np.array([shift[0], shift[1], fv_ratio[0]]),
               fv_img_index=blobs["img_index"].astype(np.int32), fv_image_shape=np.array(blobs["image_data"].shape))
    np.savez_compressed(os.path.join(OUT, "augment_hooks.npz"), **rec)
    print("augment hooks: flip", pts.shape, "voxel params", rec["voxel_params"], "fv params", rec["fv_params"],
          "fv image", rec["fv_image_shape"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "proto":      # only the .proto fixture (cheap)
        proto_goldens()
        sys.exit(0)
    main()
    proto_goldens()
    feeder_goldens()
    ingest_goldens()
    augment_goldens()
