/* CPU restatement (plain C, fp32) of the reference's BEV IoU + NMS op.
 *
 * ORACLE / TEST INFRASTRUCTURE -- never linked into the product library.
 *
 * Follows pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:
 *   segment intersection        :63-94      corner-in-box test (margin 1e-2) :51-61
 *   rotated overlap area        :104-225    rotated BEV IoU                  :227-234
 *   axis-aligned BEV IoU        :314-325    64x64 bitmask tiles              :267-311, :328-372
 * and the greedy suppression loop of pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:103-132.
 * Boxes are (N,7) [x, y, z, dx, dy, dz, heading]; only x, y, dx, dy, heading are read.
 * Build with -ffp-contract=off so every +,-,*,/ is a separately rounded fp32 op.
 * Pinned against the reference's own CPU IoU (oracle/_ref, tests/test_oracle_iou.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EPS 1e-8f

typedef struct { float x, y; } pt;

static float cross2(pt a, pt b) { return a.x * b.y - a.y * b.x; }
static float cross3(pt p1, pt p2, pt p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static int rect_cross(pt p1, pt p2, pt q1, pt q2) {
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
static int in_box2d(const float *box, pt p) {
    const float MARGIN = 1e-2f;
    float c = cosf(-box[6]), s = sinf(-box[6]);
    float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
    float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
    return fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN;
}
static int seg_intersection(pt p1, pt p0, pt q1, pt q0, pt *ans) {
    if (!rect_cross(p0, p1, q0, q1)) return 0;
    float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0);
    float s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > EPS) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}
static pt rot_center(pt c, float co, float si, pt p) {
    pt r;
    r.x = (p.x - c.x) * co + (p.y - c.y) * (-si) + c.x;
    r.y = (p.x - c.x) * si + (p.y - c.y) * co + c.y;
    return r;
}

float cg_oracle_overlap_bev(const float *a, const float *b) {
    float ahx = a[3] / 2, bhx = b[3] / 2, ahy = a[4] / 2, bhy = b[4] / 2;
    pt ca = {a[0], a[1]}, cb = {b[0], b[1]};
    pt A[5] = {{a[0] - ahx, a[1] - ahy}, {a[0] + ahx, a[1] - ahy}, {a[0] + ahx, a[1] + ahy}, {a[0] - ahx, a[1] + ahy}};
    pt B[5] = {{b[0] - bhx, b[1] - bhy}, {b[0] + bhx, b[1] - bhy}, {b[0] + bhx, b[1] + bhy}, {b[0] - bhx, b[1] + bhy}};
    float aco = cosf(a[6]), asi = sinf(a[6]), bco = cosf(b[6]), bsi = sinf(b[6]);
    for (int k = 0; k < 4; k++) { A[k] = rot_center(ca, aco, asi, A[k]); B[k] = rot_center(cb, bco, bsi, B[k]); }
    A[4] = A[0]; B[4] = B[0];
    pt cp[16], ctr = {0, 0};
    int cnt = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (seg_intersection(A[i + 1], A[i], B[j + 1], B[j], &cp[cnt])) {
                ctr.x = ctr.x + cp[cnt].x; ctr.y = ctr.y + cp[cnt].y; cnt++;
            }
    for (int k = 0; k < 4; k++) {
        if (in_box2d(a, B[k])) { ctr.x = ctr.x + B[k].x; ctr.y = ctr.y + B[k].y; cp[cnt++] = B[k]; }
        if (in_box2d(b, A[k])) { ctr.x = ctr.x + A[k].x; ctr.y = ctr.y + A[k].y; cp[cnt++] = A[k]; }
    }
    ctr.x /= cnt; ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; j++)
        for (int i = 0; i < cnt - j - 1; i++)
            if (atan2f(cp[i].y - ctr.y, cp[i].x - ctr.x) > atan2f(cp[i + 1].y - ctr.y, cp[i + 1].x - ctr.x)) {
                pt t = cp[i]; cp[i] = cp[i + 1]; cp[i + 1] = t;
            }
    float area = 0;
    for (int k = 0; k < cnt - 1; k++) {
        pt u = {cp[k].x - cp[0].x, cp[k].y - cp[0].y}, v = {cp[k + 1].x - cp[0].x, cp[k + 1].y - cp[0].y};
        area += cross2(u, v);
    }
    return fabsf(area) / 2.0f;
}

float cg_oracle_iou_bev(const float *a, const float *b) {
    float sa = a[3] * a[4], sb = b[3] * b[4];
    float so = cg_oracle_overlap_bev(a, b);
    return so / fmaxf(sa + sb - so, EPS);
}

float cg_oracle_iou_normal(const float *a, const float *b) {
    float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    float inter = w * h;
    return inter / fmaxf(a[3] * a[4] + b[3] * b[4] - inter, EPS);
}

/* mode: 0 overlap area, 1 rotated IoU, 2 axis-aligned IoU */
void cg_oracle_pairwise(const float *a, int na, const float *b, int nb, int mode, float *out) {
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++) {
            const float *p = a + 7 * i, *q = b + 7 * j;
            out[(long)i * nb + j] = mode == 0 ? cg_oracle_overlap_bev(p, q)
                                  : mode == 1 ? cg_oracle_iou_bev(p, q) : cg_oracle_iou_normal(p, q);
        }
}

/* boxes already sorted by descending score; returns number kept, indices into the sorted list */
int cg_oracle_nms_sorted(const float *boxes, int n, float thr, int rotated, long *keep) {
    unsigned char *dead = (unsigned char *)calloc(n > 0 ? n : 1, 1);
    int nk = 0;
    for (int i = 0; i < n; i++) {
        if (dead[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < n; j++) {
            if (dead[j]) continue;
            float v = rotated ? cg_oracle_iou_bev(boxes + 7 * i, boxes + 7 * j)
                              : cg_oracle_iou_normal(boxes + 7 * i, boxes + 7 * j);
            if (v > thr) dead[j] = 1;
        }
    }
    free(dead);
    return nk;
}

/* knn: pcdet/ops/knn/src/knn_cuda.cu:26-94 -- k smallest squared distances per query, ascending,
 * strict '<' so the lowest point index wins ties. xyz (n,3), q (m,3) -> idx (m,k), d2 (m,k) */
void cg_oracle_knn(const float *xyz, int n, const float *q, int m, int k, int *idx, float *d2) {
    for (int j = 0; j < m; j++) {
        int *bi = idx + (long)j * k; float *bd = d2 + (long)j * k;
        for (int t = 0; t < k; t++) { bi[t] = 0; bd[t] = 1e10f; }
        for (int i = 0; i < n; i++) {
            float dx = q[3 * j] - xyz[3 * i], dy = q[3 * j + 1] - xyz[3 * i + 1], dz = q[3 * j + 2] - xyz[3 * i + 2];
            /* the reference binary contracts this as FMUL dy*dy, FFMA dx, FFMA dz (cuobjdump of knn_cuda.cu) */
            float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            if (d < bd[k - 1]) {
                int t = k - 1;
                while (t > 0 && bd[t - 1] > d) { bd[t] = bd[t - 1]; bi[t] = bi[t - 1]; t--; }
                bd[t] = d; bi[t] = i;
            }
        }
    }
}
