"""Generates tests/golden/indoor_eval.json by running the REFERENCE's own indoor evaluation on the CPU (TEST INFRASTRUCTURE).

    python tests/golden/make_eval_golden.py          # needs /root/reference; run in the build container only

pcdet/datasets/scannet/scannet_object_eval_python/eval.py is loaded unmodified from /root/reference.  Substituted, because
they do not exist offline:
  * terminaltables.AsciiTable (printing only) -> a stub;
  * pcdet.datasets.kitti.kitti_object_eval_python.rotate_iou.rotate_iou_gpu_eval (numba-CUDA BEV intersection area,
    criterion 2) -> case "scannet": the exact axis-aligned intersection (the boxes have no heading);
                    case "rotated": oracle/rotate_iou_oracle.py, a float32 CPU restatement of rotate_iou.py itself
                    (clockwise corner rotation, :216-242).
The fixture stores the seeded inputs and the reference's result dict.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CG3D_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def make_case(seed, n_scenes, n_cls, rotated):
    rng = np.random.default_rng(seed)
    gts, dts = [], []
    for s in range(n_scenes):
        m = int(rng.integers(0, 9))
        gb = np.concatenate([rng.uniform(-3, 3, (m, 2)), rng.uniform(0, 1.5, (m, 1)), rng.uniform(0.3, 1.5, (m, 3))], 1).astype(np.float32)
        if rotated:
            gb = np.concatenate([gb, rng.uniform(-1.5, 1.5, (m, 1)).astype(np.float32)], 1)
        gc = rng.integers(0, n_cls, m)
        gts.append({"gt_num": m, "gt_boxes_upright_depth": gb, "class": gc})
        # detections: jittered copies of the ground truth (some duplicated, some mislabelled) + clutter
        det_b, det_l, det_s = [], [], []
        for i in range(m):
            for _ in range(int(rng.integers(0, 3))):
                b = np.zeros(7, np.float32)
                b[:gb.shape[1]] = gb[i]
                b[:6] += rng.normal(0, 0.12, 6).astype(np.float32) * np.array([1, 1, 1, .5, .5, .5], np.float32)
                b[3:6] = np.abs(b[3:6]) + 0.05
                det_b.append(b)
                det_l.append(int(gc[i]) if rng.random() > 0.15 else int(rng.integers(0, n_cls)))
                det_s.append(np.float32(np.round(rng.uniform(0.05, 1.0), 2)))          # rounded: score ties occur
        for _ in range(int(rng.integers(0, 6))):
            b = np.concatenate([rng.uniform(-3, 3, 2), rng.uniform(0, 1.5, 1), rng.uniform(0.3, 1.5, 3),
                                rng.uniform(-1.5, 1.5, 1) * float(rotated)]).astype(np.float32)
            det_b.append(b)
            det_l.append(int(rng.integers(0, n_cls + 1)))                              # class n_cls has no ground truth
            det_s.append(np.float32(np.round(rng.uniform(0.05, 0.6), 2)))
        dts.append({"labels_3d": np.array(det_l, np.int64), "boxes_3d": np.array(det_b, np.float32).reshape(-1, 7),
                    "scores_3d": np.array(det_s, np.float32)})
    return gts, dts


def load_reference_eval(rinc_fn):
    for name in ("pcdet", "pcdet.datasets", "pcdet.datasets.kitti", "pcdet.datasets.kitti.kitti_object_eval_python"):
        sys.modules[name] = types.ModuleType(name)
    ri = types.ModuleType("pcdet.datasets.kitti.kitti_object_eval_python.rotate_iou")
    ri.rotate_iou_gpu_eval = rinc_fn
    sys.modules[ri.__name__] = ri
    tt = types.ModuleType("terminaltables")

    class AsciiTable:
        def __init__(self, data):
            self.table = "\n".join(" | ".join(map(str, r)) for r in data)
            self.inner_footing_row_border = False
    tt.AsciiTable = AsciiTable
    sys.modules["terminaltables"] = tt
    path = os.path.join(REF, "pcdet/datasets/scannet/scannet_object_eval_python/eval.py")
    spec = importlib.util.spec_from_file_location("reference_indoor_eval", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from pcdet_shim_path import indoor_eval_module          # our implementation, loaded by path (see below)

    def rinc_axis(b, q, criterion):
        assert criterion == 2
        return indoor_eval_module.axis_aligned_bev_overlap(
            np.concatenate([b[:, :2], np.zeros((len(b), 1)), b[:, 2:4], np.ones((len(b), 1))], 1),
            np.concatenate([q[:, :2], np.zeros((len(q), 1)), q[:, 2:4], np.ones((len(q), 1))], 1)).astype(np.float32)

    def rinc_rot(b, q, criterion):
        # the numba-CUDA kernel itself restated on the CPU, function by function (oracle/rotate_iou_oracle.py): corners
        # rotated CLOCKWISE as rbbox_to_corners does -- NOT the iou3d_nms convention, which a first version of this
        # fixture used and which made the rotated test circular
        from oracle import rotate_iou_oracle
        return rotate_iou_oracle.rotate_iou_eval(b, q, criterion)

    out = {}
    for name, rotated, fn in (("scannet", False, rinc_axis), ("rotated", True, rinc_rot)):
        ev = load_reference_eval(fn)
        gts, dts = make_case(7 if not rotated else 8, 10, 5, rotated)
        label2cat = {i: f"c{i}" for i in range(6)}
        res = ev.indoor_eval(gts, dts, [0.25, 0.5], label2cat)
        out[name] = {"seed": 7 if not rotated else 8, "n_scenes": 10, "n_cls": 5, "result": res}
        print(name, {k: round(v, 4) for k, v in res.items() if k.startswith("m")})
    json.dump(out, open(os.path.join(HERE, "indoor_eval.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    # our module is loaded by path so that the stub 'pcdet' packages installed for the reference import do not shadow it
    spec = importlib.util.spec_from_file_location("indoor_eval_module", os.path.join(ROOT, "pcdet/datasets/indoor_eval.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    shim = types.ModuleType("pcdet_shim_path")
    shim.indoor_eval_module = m
    sys.modules["pcdet_shim_path"] = shim
    main()
