// Device functions of the rotated / axis-aligned BEV IoU, shared by iou3d.cu and proposal.cu.
// Arithmetic follows the reference binaries operation for operation (see oracle/jmodt_oracle.c).
#pragma once

#include "common.cuh"

namespace jmb {

struct P2 {
    float x, y;
};

__device__ __forceinline__ P2 rot_center(P2 c, float cs, float sn, P2 p) {
    const float dx = __fsub_rn(p.x, c.x), dy = __fsub_rn(p.y, c.y);
    P2 r;
    r.x = __fadd_rn(__fmaf_rn(dx, cs, __fmul_rn(dy, sn)), c.x);
    r.y = __fadd_rn(__fmaf_rn(cs, dy, -__fmul_rn(sn, dx)), c.y);
    return r;
}

// iou3d_kernel.cu:48-63
__device__ __forceinline__ bool in_box2d(const float *box, float cs, float sn, P2 p) {
    const float MARGIN = 1e-5f;
    const float cx = __fmul_rn(__fadd_rn(box[0], box[2]), 0.5f);
    const float cy = __fmul_rn(__fadd_rn(box[1], box[3]), 0.5f);
    const float dx = __fsub_rn(p.x, cx), dy = __fsub_rn(p.y, cy);
    const float rx = __fadd_rn(__fmaf_rn(dx, cs, __fmul_rn(dy, sn)), cx);
    const float ry = __fadd_rn(__fmaf_rn(cs, dy, -__fmul_rn(sn, dx)), cy);
    return rx > __fsub_rn(box[0], MARGIN) && rx < __fadd_rn(box[2], MARGIN) &&
           ry > __fsub_rn(box[1], MARGIN) && ry < __fadd_rn(box[3], MARGIN);
}

// iou3d_kernel.cu:65-96
__device__ __forceinline__ bool seg_intersection(P2 p1, P2 p0, P2 q1, P2 q0, P2 &ans) {
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
          fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return false;
    const float s1 = fmsub2(__fsub_rn(q0.x, p0.x), __fsub_rn(p1.y, p0.y), __fsub_rn(p1.x, p0.x),
                            __fsub_rn(q0.y, p0.y));
    const float pa = __fmul_rn(__fsub_rn(p1.x, p0.x), __fsub_rn(q1.y, p0.y));
    const float pb = __fmul_rn(__fsub_rn(q1.x, p0.x), __fsub_rn(p1.y, p0.y));
    const float s2 = __fsub_rn(pa, pb);
    const float s3 = fmsub2(__fsub_rn(p0.x, q0.x), __fsub_rn(q1.y, q0.y), __fsub_rn(q1.x, q0.x),
                            __fsub_rn(p0.y, q0.y));
    const float s4 = fmsub2(__fsub_rn(q1.x, q0.x), __fsub_rn(p1.y, q0.y), __fsub_rn(p1.x, q0.x),
                            __fsub_rn(q1.y, q0.y));
    if (!(__fmul_rn(s1, s2) > 0.f && __fmul_rn(s3, s4) > 0.f)) return false;
    const float s5 = __fsub_rn(pb, pa);
    const float den = __fsub_rn(s5, s1);
    if ((double)fabsf(den) > 1e-8) {
        ans.x = __fdiv_rn(fmsub2(s5, q0.x, s1, q1.x), den);
        ans.y = __fdiv_rn(fmsub2(s5, q0.y, s1, q1.y), den);
    } else {
        const float a0 = __fsub_rn(p0.y, p1.y), b0 = __fsub_rn(p1.x, p0.x),
                    c0 = fmsub2(p0.x, p1.y, p1.x, p0.y);
        const float a1 = __fsub_rn(q0.y, q1.y), b1 = __fsub_rn(q1.x, q0.x),
                    c1 = fmsub2(q0.x, q1.y, q1.x, q0.y);
        const float D = fmsub2(a0, b1, a1, b0);
        ans.x = __fdiv_rn(fmsub2(b0, c1, b1, c0), D);
        ans.y = __fdiv_rn(fmsub2(a1, c0, a0, c1), D);
    }
    return true;
}

// iou3d_kernel.cu:108-212
static __device__ float box_overlap(const float *box_a, const float *box_b) {
    const float a_x1 = box_a[0], a_y1 = box_a[1], a_x2 = box_a[2], a_y2 = box_a[3];
    const float b_x1 = box_b[0], b_y1 = box_b[1], b_x2 = box_b[2], b_y2 = box_b[3];
    const P2 ca = {__fmul_rn(__fadd_rn(a_x1, a_x2), 0.5f), __fmul_rn(__fadd_rn(a_y1, a_y2), 0.5f)};
    const P2 cb = {__fmul_rn(__fadd_rn(b_x1, b_x2), 0.5f), __fmul_rn(__fadd_rn(b_y1, b_y2), 0.5f)};
    P2 A[5] = {{a_x1, a_y1}, {a_x2, a_y1}, {a_x2, a_y2}, {a_x1, a_y2}, {0.f, 0.f}};
    P2 B[5] = {{b_x1, b_y1}, {b_x2, b_y1}, {b_x2, b_y2}, {b_x1, b_y2}, {0.f, 0.f}};
    const float acs = cosf(box_a[4]), asn = sinf(box_a[4]);
    const float bcs = cosf(box_b[4]), bsn = sinf(box_b[4]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        A[k] = rot_center(ca, acs, asn, A[k]);
        B[k] = rot_center(cb, bcs, bsn, B[k]);
    }
    A[4] = A[0];
    B[4] = B[0];

    P2 cp[16];
    float ang[16];
    float pcx = 0.f, pcy = 0.f;
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            P2 r;
            if (seg_intersection(A[i + 1], A[i], B[j + 1], B[j], r)) {
                pcx = __fadd_rn(pcx, r.x);
                pcy = __fadd_rn(pcy, r.y);
                cp[cnt++] = r;
            }
        }
    }
    // check_in_box2d evaluates cos/sin of the NEGATED angle (iou3d_kernel.cu:53)
    const float nacs = cosf(-box_a[4]), nasn = sinf(-box_a[4]);
    const float nbcs = cosf(-box_b[4]), nbsn = sinf(-box_b[4]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (in_box2d(box_a, nacs, nasn, B[k])) {
            pcx = __fadd_rn(pcx, B[k].x);
            pcy = __fadd_rn(pcy, B[k].y);
            cp[cnt++] = B[k];
        }
        if (in_box2d(box_b, nbcs, nbsn, A[k])) {
            pcx = __fadd_rn(pcx, A[k].x);
            pcy = __fadd_rn(pcy, A[k].y);
            cp[cnt++] = A[k];
        }
    }
    if (cnt < 3) return 0.f;  // fewer than 3 vertices: the reference area loop adds nothing / one zero-area term
    pcx = __fdiv_rn(pcx, (float)cnt);
    pcy = __fdiv_rn(pcy, (float)cnt);
    for (int k = 0; k < cnt; ++k) ang[k] = atan2f(__fsub_rn(cp[k].y, pcy), __fsub_rn(cp[k].x, pcx));
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (ang[i] > ang[i + 1]) {
                const P2 t = cp[i]; cp[i] = cp[i + 1]; cp[i + 1] = t;
                const float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ax = __fsub_rn(cp[k].x, cp[0].x), ay = __fsub_rn(cp[k].y, cp[0].y);
        const float bx = __fsub_rn(cp[k + 1].x, cp[0].x), by = __fsub_rn(cp[k + 1].y, cp[0].y);
        area = __fadd_rn(area, fmsub2(ax, by, ay, bx));
    }
    return __fmul_rn(fabsf(area), 0.5f);
}

// iou3d_kernel.cu:214-221 — SASS: ov / fmaxf(fma(wa,ha, fl(wb*hb)) - ov, 1e-8f)
__device__ __forceinline__ float iou_bev(const float *a, const float *b) {
    const float sb = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
    const float u = __fmaf_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]), sb);
    const float ov = box_overlap(a, b);
    return __fdiv_rn(ov, fmaxf(__fsub_rn(u, ov), 1e-8f));
}

// iou3d_kernel.cu:295-303 — PTX: inter / fmaxf(fma(wb,hb, fl(wa*ha)) - inter, 1e-8f)
__device__ __forceinline__ float iou_normal(const float *a, const float *b) {
    const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
    const float inter = __fmul_rn(width, height);
    const float sa = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
    const float u = __fmaf_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]), sa);
    return __fdiv_rn(inter, fmaxf(__fsub_rn(u, inter), 1e-8f));
}


}  // namespace jmb
