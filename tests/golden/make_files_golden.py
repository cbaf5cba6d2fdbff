"""Generates tests/golden/indoor_files.npz (TEST INFRASTRUCTURE; run in the build container only, needs /root/reference).

`global_alignment` and `points_random_sampling` are executed VERBATIM from the reference file
pcdet/datasets/augmentor/augmentor_utils.py (the two function bodies are cut out of the source text and exec'd: the
module itself imports pcdet packages that need compiled extensions).  The fixture stores the inputs and their outputs.
"""
import os

import numpy as np

REF = os.environ.get("CG3D_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def grab(src, fn):
    i = src.index("def " + fn + "(")
    j = src.find("\ndef ", i + 1)
    return src[i:] if j < 0 else src[i:j]


def main():
    src = open(os.path.join(REF, "pcdet/datasets/augmentor/augmentor_utils.py")).read()
    ns = {"np": np}
    exec(grab(src, "global_alignment"), ns)
    exec(grab(src, "points_random_sampling"), ns)
    rng = np.random.default_rng(3)
    pts = rng.normal(0, 2, (50, 6)).astype(np.float32)
    th = 0.7
    M = np.eye(4, dtype=np.float32)
    M[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
    M[:3, 3] = [0.5, -1.25, 0.1]
    out = ns["global_alignment"](pts.copy(), M, 2)
    np.random.seed(5)
    _, c1 = ns["points_random_sampling"](pts, 20, return_choices=True)
    np.random.seed(5)
    _, c2 = ns["points_random_sampling"](pts, 80, return_choices=True)
    np.savez(os.path.join(HERE, "indoor_files.npz"), pts=pts, M=M, aligned=out, c1=c1, c2=c2)
    print("written", out[:2, :3])


if __name__ == "__main__":
    main()
