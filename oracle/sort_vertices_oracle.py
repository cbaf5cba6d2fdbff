"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's polygon vertex ordering,
pcdet/ops/rotated_iou/cuda_op/sort_vert_kernel.cu:15-134 (sort_vertices_forward), in numpy.

Only tests/ (and the golden generators under tests/golden/) may import this.  Pinned on the GPU box against the
reference's own CUDA op built by oracle/build_ref.py (tests/test_gpu_ref_ops.py) and, through it, against cg3d_sort_vertices.

compare_vertices (:15-40) has no return statement when a y coordinate is exactly 0; like the CUDA product kernel this
restatement returns False there.  fp32 arithmetic with the double-precision EPSILON comparisons of the C source.
"""
import numpy as np

EPS = 1e-8                     # sort_vert_kernel.cu:8 (a double literal)
MAX_IDX, INTER_OFF = 9, 8      # :6-7
f32 = np.float32


def compare_vertices(x1, y1, x2, y2) -> bool:          # :15-40
    x1, y1, x2, y2 = f32(x1), f32(y1), f32(x2), f32(y2)
    if abs(float(f32(x1 - x2))) < EPS and abs(float(f32(y2 - y1))) < EPS:
        return False
    if y1 > 0 and y2 < 0:
        return True
    if y1 < 0 and y2 > 0:
        return False
    n1 = f32(float(f32(f32(x1 * x1) + f32(y1 * y1))) + EPS)
    n2 = f32(float(f32(f32(x2 * x2) + f32(y2 * y2))) + EPS)
    lhs = f32(f32(f32(abs(x1) * x1) / n1) - f32(f32(abs(x2) * x2) / n2))
    if y1 > 0 and y2 > 0:
        return float(lhs) > EPS
    if y1 < 0 and y2 < 0:
        return float(lhs) < EPS
    return False


def sort_vertices(vertices: np.ndarray, mask: np.ndarray, num_valid: np.ndarray) -> np.ndarray:
    """vertices (B, N, M, 2) fp32 (normalised around the polygon's mean), mask (B, N, M) bool, num_valid (B, N) int
    -> idx (B, N, 9) int32: the valid vertices in anti-clockwise order, the first one repeated, padded with the index of
    an invalid intersection point (:42-128)."""
    B, N, M, _ = vertices.shape
    out = np.zeros((B, N, MAX_IDX), np.int32)
    for b in range(B):
        for i in range(N):
            v, mk, nv = vertices[b, i], mask[b, i], int(num_valid[b, i])
            pad = 0
            for j in range(INTER_OFF, M):                                  # :55-60
                if not mk[j]:
                    pad = j
                    break
            idx = out[b, i]
            if nv < 3:                                                     # :61-66
                idx[:] = pad
                continue
            for j in range(nv):                                            # :70-96
                x_min, y_min, take = f32(1.0), f32(-EPS), 0
                for k in range(M):
                    x, y = v[k]
                    if j == 0:
                        if mk[k] and compare_vertices(x, y, x_min, y_min):
                            x_min, y_min, take = x, y, k
                    else:
                        x2, y2 = v[idx[j - 1]]
                        if mk[k] and compare_vertices(x, y, x_min, y_min) and compare_vertices(x2, y2, x, y):
                            x_min, y_min, take = x, y, k
                idx[j] = take
            idx[nv] = idx[0]                                               # :98
            idx[nv + 1:] = pad                                             # :101-103
            if nv == 8:                                                    # :109-123: two identical boxes
                counter = sum(int(idx[k] == idx[j]) for j in range(4) for k in range(4, INTER_OFF))
                if counter == 4:
                    idx[4] = idx[0]
                    idx[5:] = pad
    return out
