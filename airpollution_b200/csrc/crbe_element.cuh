// Per-element Crouzeix-Raviart matrices, evaluated in the reference's operation
// order so that assembled values agree to the last bit wherever the reference's
// own arithmetic is exact or unfused.  Translation units including this header
// are compiled with -fmad=false (no contraction of a*b+c into FMA).
//
//   stiffness  crbe.py:249-277   K = (D*area) * G (B^T B) G^T,  B = adj(J)/|det J|
//   mass       crbe.py:280-282   diag ((1/6)*2)*area, explicit zero off-diagonals
//   advection  crbe.py:284-313   A[a][b] = 2*((area/6) * (grad_phi_b . v)), same for every a
//
// Note the reference's B^T B is J^-T J^-1 (crbe.py:272-273), kept as is.
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define CRBE_HD __host__ __device__ __forceinline__
#else
#define CRBE_HD inline
#endif

struct CrbeElement {
    double K[3][3];
    double Arow[3];
    double Md;
};

// Reference-element gradients G = dphi/dxi  (crbe.py:198-203)
#define CRBE_G(a, i) ((a) == 0 ? 2.0 : ((a) == 1 ? ((i) == 0 ? -2.0 : 0.0) : ((i) == 0 ? 0.0 : -2.0)))

CRBE_HD void crbe_element_eval(double x0, double y0, double x1, double y1, double x2, double y2, double area,
                               double D, double vx, double vy, CrbeElement& e) {
    // J = [v1-v0 | v2-v0]                                           crbe.py:256-258
    const double j00 = x1 - x0, j10 = y1 - y0, j01 = x2 - x0, j11 = y2 - y0;
    const double det = fabs(j00 * j11 - j01 * j10);                  // :261 (sign dropped)
    const double b00 = j11 / det, b01 = (-j01) / det;                // :264-267
    const double b10 = (-j10) / det, b11 = j00 / det;
    // BTB = B^T B                                                   :273
    const double m00 = b00 * b00 + b10 * b10;
    const double m01 = b00 * b01 + b10 * b11;
    const double m11 = b01 * b01 + b11 * b11;
    const double m10 = m01;
    const double da = D * area;                                      // :277 (left to right)
    double gb[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {                                    // G @ BTB
        gb[a][0] = CRBE_G(a, 0) * m00 + CRBE_G(a, 1) * m10;
        gb[a][1] = CRBE_G(a, 0) * m01 + CRBE_G(a, 1) * m11;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)                                  // (G BTB) @ G^T, then * (D*area)
            e.K[a][b] = da * (gb[a][0] * CRBE_G(b, 0) + gb[a][1] * CRBE_G(b, 1));
    e.Md = ((1.0 / 6.0) * 2) * area;                                 // :282
    const double phi_int = area / 6.0;                               // :310
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        // grad_phi[b] = B^T G[b]                                    :305
        const double gx = b00 * CRBE_G(b, 0) + b10 * CRBE_G(b, 1);
        const double gy = b01 * CRBE_G(b, 0) + b11 * CRBE_G(b, 1);
        e.Arow[b] = 2 * (phi_int * (gx * vx + gy * vy));             // :311-313
    }
}

// 0.5*|(x2-x1)(y3-y1) - (x3-x1)(y2-y1)|                              crbe.py:152
CRBE_HD double crbe_triangle_area(double x1, double y1, double x2, double y2, double x3, double y3) {
    return 0.5 * fabs((x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1));
}

// Advection row only (crbe.py:284-313), same operation order as crbe_element_eval: used when the velocity
// changes every step and A is rebuilt while K and M stay.
CRBE_HD void crbe_element_advection(double x0, double y0, double x1, double y1, double x2, double y2, double area, double vx,
                                    double vy, double (&arow)[3]) {
    const double j00 = x1 - x0, j10 = y1 - y0, j01 = x2 - x0, j11 = y2 - y0;
    const double det = fabs(j00 * j11 - j01 * j10);
    const double b00 = j11 / det, b01 = (-j01) / det;
    const double b10 = (-j10) / det, b11 = j00 / det;
    const double phi_int = area / 6.0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const double gx = b00 * CRBE_G(b, 0) + b10 * CRBE_G(b, 1);
        const double gy = b01 * CRBE_G(b, 0) + b11 * CRBE_G(b, 1);
        arow[b] = 2 * (phi_int * (gx * vx + gy * vy));
    }
}
