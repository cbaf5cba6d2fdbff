"""CPU restatement of the reference's numba-CUDA BEV intersection (TEST INFRASTRUCTURE: only tests/ and the golden
generators may import this).

Follows pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py of the reference, function by function, in float32:
  rbbox_to_corners (:216-242: corners generated clockwise and rotated CLOCKWISE, x' = cos x + sin y, y' = -sin x + cos y),
  point_in_quadrilateral (:162-180), line_segment_intersection (:66-116), quadrilateral_intersection (:183-213),
  sort_vertex_in_convex_polygon (:30-63), area (:20-27), inter (:245-259), devRotateIoUEval criterion 2 (:262-274).
It is what scannet_object_eval_python/eval.py:38-42 gets from rotate_iou_gpu_eval(boxes[:, [0, 1, 3, 4, 6]], ..., 2).
Pinned by hand-derivable answers in tests/test_indoor_eval.py (identical boxes, axis-aligned shifts, a 90-degree turn,
the advisor's A = (0,0,2,1,.5) / B = (.3,.4,1.8,1.1,.3) pair whose clockwise and counter-clockwise answers differ).
"""
import math

import numpy as np

f32 = np.float32


def rbbox_to_corners(rbbox):
    angle = rbbox[4]
    a_cos, a_sin = f32(math.cos(angle)), f32(math.sin(angle))
    cx, cy, xd, yd = rbbox[0], rbbox[1], rbbox[2], rbbox[3]
    xs = [-xd / f32(2), -xd / f32(2), xd / f32(2), xd / f32(2)]
    ys = [-yd / f32(2), yd / f32(2), yd / f32(2), -yd / f32(2)]
    c = np.zeros(8, f32)
    for i in range(4):
        c[2 * i] = a_cos * xs[i] + a_sin * ys[i] + cx
        c[2 * i + 1] = -a_sin * xs[i] + a_cos * ys[i] + cy
    return c


def point_in_quadrilateral(px, py, c):
    ab0, ab1 = c[2] - c[0], c[3] - c[1]
    ad0, ad1 = c[6] - c[0], c[7] - c[1]
    ap0, ap1 = px - c[0], py - c[1]
    abab, abap = ab0 * ab0 + ab1 * ab1, ab0 * ap0 + ab1 * ap1
    adad, adap = ad0 * ad0 + ad1 * ad1, ad0 * ap0 + ad1 * ap1
    return abab >= abap and abap >= 0 and adad >= adap and adap >= 0


def line_segment_intersection(p1, p2, i, j):
    A = (p1[2 * i], p1[2 * i + 1])
    B = (p1[2 * ((i + 1) % 4)], p1[2 * ((i + 1) % 4) + 1])
    C = (p2[2 * j], p2[2 * j + 1])
    D = (p2[2 * ((j + 1) % 4)], p2[2 * ((j + 1) % 4) + 1])
    BA0, BA1 = B[0] - A[0], B[1] - A[1]
    DA0, CA0, DA1, CA1 = D[0] - A[0], C[0] - A[0], D[1] - A[1], C[1] - A[1]
    acd = DA1 * CA0 > CA1 * DA0
    bcd = (D[1] - B[1]) * (C[0] - B[0]) > (C[1] - B[1]) * (D[0] - B[0])
    if acd != bcd:
        abc = CA1 * BA0 > BA1 * CA0
        abd = DA1 * BA0 > BA1 * DA0
        if abc != abd:
            DC0, DC1 = D[0] - C[0], D[1] - C[1]
            ABBA = A[0] * B[1] - B[0] * A[1]
            CDDC = C[0] * D[1] - D[0] * C[1]
            DH = BA1 * DC0 - BA0 * DC1
            Dx = ABBA * DC0 - BA0 * CDDC
            Dy = ABBA * DC1 - BA1 * CDDC
            return f32(Dx / DH), f32(Dy / DH)
    return None


def quadrilateral_intersection(p1, p2):
    pts = []
    for i in range(4):
        if point_in_quadrilateral(p1[2 * i], p1[2 * i + 1], p2):
            pts.append((p1[2 * i], p1[2 * i + 1]))
        if point_in_quadrilateral(p2[2 * i], p2[2 * i + 1], p1):
            pts.append((p2[2 * i], p2[2 * i + 1]))
    for i in range(4):
        for j in range(4):
            r = line_segment_intersection(p1, p2, i, j)
            if r is not None:
                pts.append(r)
    return pts


def sort_vertex_in_convex_polygon(pts):
    n = len(pts)
    if n == 0:
        return pts
    cx = f32(sum(p[0] for p in pts) / f32(n))
    cy = f32(sum(p[1] for p in pts) / f32(n))
    vs = []
    for p in pts:
        v0, v1 = p[0] - cx, p[1] - cy
        d = f32(math.sqrt(v0 * v0 + v1 * v1))
        v0, v1 = v0 / d, v1 / d
        if v1 < 0:
            v0 = -2 - v0
        vs.append(v0)
    pts = list(pts)
    for i in range(1, n):                       # the reference's insertion sort (stable)
        if vs[i - 1] > vs[i]:
            temp, tp = vs[i], pts[i]
            j = i
            while j > 0 and vs[j - 1] > temp:
                vs[j], pts[j] = vs[j - 1], pts[j - 1]
                j -= 1
            vs[j], pts[j] = temp, tp
    return pts


def area(pts):
    a = f32(0)
    for i in range(len(pts) - 2):
        p0, p1, p2 = pts[0], pts[i + 1], pts[i + 2]
        a += abs(((p0[0] - p2[0]) * (p1[1] - p2[1]) - (p0[1] - p2[1]) * (p1[0] - p2[0])) / f32(2))
    return a


def inter(r1, r2):
    c1, c2 = rbbox_to_corners(r1), rbbox_to_corners(r2)
    return area(sort_vertex_in_convex_polygon(quadrilateral_intersection(c1, c2)))


def rotate_iou_eval(boxes, qboxes, criterion=-1):
    """rotate_iou_gpu_eval: boxes (n, 5) / qboxes (k, 5) = (x, y, dx, dy, angle) -> (n, k) float32."""
    boxes, qboxes = np.asarray(boxes, f32), np.asarray(qboxes, f32)
    out = np.zeros((len(boxes), len(qboxes)), f32)
    with np.errstate(all="ignore"):
        for i, b in enumerate(boxes):
            for j, q in enumerate(qboxes):
                ai = inter(b, q)
                a1, a2 = b[2] * b[3], q[2] * q[3]
                out[i, j] = (ai / (a1 + a2 - ai) if criterion == -1 else ai / a1 if criterion == 0 else
                             ai / a2 if criterion == 1 else ai)
    return out
