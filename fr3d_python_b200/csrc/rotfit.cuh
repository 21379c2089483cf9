// Small fp64 device helpers shared by the frame-fit, Kabsch and discrepancy kernels.
//
// The reference finds the optimal proper rotation with LAPACK's SVD plus a
// reflection fix (fr3d/geometry/superpositions.py:53-78).  On the device the
// same optimum is obtained from Horn's quaternion form: the rotation is the
// eigenvector of the largest eigenvalue of a symmetric 4x4 matrix built from
// the 3x3 covariance, found here by cyclic Jacobi sweeps kept entirely in
// registers.  It is always a proper rotation, so the reference's det test is
// implicit, and it agrees with the SVD route to ~1e-15 whenever the optimum is
// unique (the reference is itself arbitrary when it is not).
#pragma once

#include <math.h>

namespace fr3d {

// One Jacobi rotation on the symmetric 4x4 matrix a (full storage) and the
// accumulated eigenvector matrix v, zeroing a[P][Q].
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(double (&a)[4][4], double (&v)[4][4]) {
    const double apq = a[P][Q];
    if (apq == 0.0) return;
    const double app = a[P][P], aqq = a[Q][Q];
    // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)) with theta = d / h, d = aqq - app, h = 2 apq, written as
    // sgn(d) h / (|d| + sqrt(d^2 + h^2)): one division instead of two, no overflow for tiny apq
    const double d = aqq - app, h = 2.0 * apq;
    const double t = copysign(1.0, d) * h / (fabs(d) + sqrt(fma(d, d, h * h)));
    const double c = rsqrt(fma(t, t, 1.0));
    const double s = t * c;
    a[P][P] = app - t * apq;
    a[Q][Q] = aqq + t * apq;
    a[P][Q] = 0.0;
    a[Q][P] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r != P && r != Q) {
            const double arp = a[r][P], arq = a[r][Q];
            const double np_ = c * arp - s * arq;
            const double nq_ = s * arp + c * arq;
            a[r][P] = np_; a[P][r] = np_;
            a[r][Q] = nq_; a[Q][r] = nq_;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double vrp = v[r][P], vrq = v[r][Q];
        v[r][P] = c * vrp - s * vrq;
        v[r][Q] = s * vrp + c * vrq;
    }
}

// Proper rotation Q (row major 3x3) maximising sum_k b_k . (Q a_k), where
// M[i][j] = sum_k a_k[i] * b_k[j].
__device__ __forceinline__ void horn_rotation(const double (&M)[3][3], double (&Qm)[3][3]) {
    double a[4][4], v[4][4];
    const double Sxx = M[0][0], Sxy = M[0][1], Sxz = M[0][2];
    const double Syx = M[1][0], Syy = M[1][1], Syz = M[1][2];
    const double Szx = M[2][0], Szy = M[2][1], Szz = M[2][2];
    a[0][0] = Sxx + Syy + Szz;
    a[0][1] = Syz - Szy; a[0][2] = Szx - Sxz; a[0][3] = Sxy - Syx;
    a[1][1] = Sxx - Syy - Szz; a[1][2] = Sxy + Syx; a[1][3] = Szx + Sxz;
    a[2][2] = -Sxx + Syy - Szz; a[2][3] = Syz + Szy;
    a[3][3] = -Sxx - Syy + Szz;
    a[1][0] = a[0][1]; a[2][0] = a[0][2]; a[3][0] = a[0][3];
    a[2][1] = a[1][2]; a[3][1] = a[1][3]; a[3][2] = a[2][3];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[r][c] = (r == c) ? 1.0 : 0.0;

    // Cyclic sweeps; quadratic convergence makes 4x4 reach fp64 round-off in <= 6.
    for (int sweep = 0; sweep < 10; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[0][3]) + fabs(a[1][2]) + fabs(a[1][3]) +
                           fabs(a[2][3]);
        const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]) + fabs(a[3][3]);
        if (off <= 1e-18 * diag || off == 0.0) break;
        jacobi_rotate<0, 1>(a, v);
        jacobi_rotate<0, 2>(a, v);
        jacobi_rotate<0, 3>(a, v);
        jacobi_rotate<1, 2>(a, v);
        jacobi_rotate<1, 3>(a, v);
        jacobi_rotate<2, 3>(a, v);
    }
    // largest eigenvalue
    int best = 0;
    double lam = a[0][0];
    if (a[1][1] > lam) { lam = a[1][1]; best = 1; }
    if (a[2][2] > lam) { lam = a[2][2]; best = 2; }
    if (a[3][3] > lam) { lam = a[3][3]; best = 3; }
    double q0, q1, q2, q3;
    if (best == 0)      { q0 = v[0][0]; q1 = v[1][0]; q2 = v[2][0]; q3 = v[3][0]; }
    else if (best == 1) { q0 = v[0][1]; q1 = v[1][1]; q2 = v[2][1]; q3 = v[3][1]; }
    else if (best == 2) { q0 = v[0][2]; q1 = v[1][2]; q2 = v[2][2]; q3 = v[3][2]; }
    else                { q0 = v[0][3]; q1 = v[1][3]; q2 = v[2][3]; q3 = v[3][3]; }
    const double nrm = rsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 *= nrm; q1 *= nrm; q2 *= nrm; q3 *= nrm;
    Qm[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
    Qm[0][1] = 2.0 * (q1 * q2 - q0 * q3);
    Qm[0][2] = 2.0 * (q1 * q3 + q0 * q2);
    Qm[1][0] = 2.0 * (q1 * q2 + q0 * q3);
    Qm[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
    Qm[1][2] = 2.0 * (q2 * q3 - q0 * q1);
    Qm[2][0] = 2.0 * (q1 * q3 - q0 * q2);
    Qm[2][1] = 2.0 * (q2 * q3 + q0 * q1);
    Qm[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
}

// 3x3 determinant
__device__ __forceinline__ double det3(double a, double b, double c, double d, double e, double f, double g,
                                       double h, double i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// Cofactors of row 0 of the symmetric 4x4 matrix A (= a column of adj(A)): for singular A = N - lambda I
// this is the null vector, i.e. the eigenvector of lambda.
__device__ __forceinline__ void cof_row0(const double (&A)[4][4], double &q0, double &q1, double &q2, double &q3) {
    q0 = det3(A[1][1], A[1][2], A[1][3], A[2][1], A[2][2], A[2][3], A[3][1], A[3][2], A[3][3]);
    q1 = -det3(A[1][0], A[1][2], A[1][3], A[2][0], A[2][2], A[2][3], A[3][0], A[3][2], A[3][3]);
    q2 = det3(A[1][0], A[1][1], A[1][3], A[2][0], A[2][1], A[2][3], A[3][0], A[3][1], A[3][3]);
    q3 = -det3(A[1][0], A[1][1], A[1][2], A[2][0], A[2][1], A[2][2], A[3][0], A[3][1], A[3][2]);
}

// Same optimum as horn_rotation, ~8x fewer flops, for the all-vs-all discrepancy kernel: the largest
// root of the characteristic quartic lambda^4 + C2 lambda^2 + C1 lambda + C0 of Horn's matrix is found
// by Newton's iteration from the upper bound e0 = (sum|a|^2 + sum|b|^2) / 2 (monotone convergence from
// the right of the largest root), its eigenvector is a column of adj(N - lambda I), and one Rayleigh-
// quotient step (quadratically accurate eigenvalue) followed by a second adjugate removes the error the
// polynomial root carries.  Falls back to the Jacobi solver when the adjugate column is (near) zero,
// i.e. when the optimum is (nearly) not unique.
__device__ __forceinline__ void horn_rotation_fast(const double (&M)[3][3], double e0, double (&Qm)[3][3]) {
    const double Sxx = M[0][0], Sxy = M[0][1], Sxz = M[0][2];
    const double Syx = M[1][0], Syy = M[1][1], Syz = M[1][2];
    const double Szx = M[2][0], Szy = M[2][1], Szz = M[2][2];
    double N[4][4];
    N[0][0] = Sxx + Syy + Szz;
    N[0][1] = Syz - Szy; N[0][2] = Szx - Sxz; N[0][3] = Sxy - Syx;
    N[1][1] = Sxx - Syy - Szz; N[1][2] = Sxy + Syx; N[1][3] = Szx + Sxz;
    N[2][2] = -Sxx + Syy - Szz; N[2][3] = Syz + Szy;
    N[3][3] = -Sxx - Syy + Szz;
    N[1][0] = N[0][1]; N[2][0] = N[0][2]; N[3][0] = N[0][3];
    N[2][1] = N[1][2]; N[3][1] = N[1][3]; N[3][2] = N[2][3];
    const double C2 = -2.0 * (Sxx * Sxx + Sxy * Sxy + Sxz * Sxz + Syx * Syx + Syy * Syy + Syz * Syz + Szx * Szx +
                              Szy * Szy + Szz * Szz);
    const double C1 = -8.0 * det3(Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz);
    double c0, c1, c2, c3;
    cof_row0(N, c0, c1, c2, c3);
    const double C0 = N[0][0] * c0 + N[0][1] * c1 + N[0][2] * c2 + N[0][3] * c3;      // det(N)
    double lam = e0;
    bool ok = e0 > 0.0;
    for (int it = 0; it < 40 && ok; ++it) {
        const double x2 = lam * lam;
        const double b = (x2 + C2) * lam;
        const double a = b + C1;
        const double fp = 2.0 * x2 * lam + b + a;
        if (fp == 0.0) { ok = false; break; }
        const double delta = (a * lam + C0) / fp;
        lam -= delta;
        if (fabs(delta) <= 1e-13 * fabs(lam)) break;
    }
    double q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    if (ok) {
        double A[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) A[r][c] = N[r][c] - (r == c ? lam : 0.0);
        cof_row0(A, q0, q1, q2, q3);
        double nn = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
        const double scale = e0 * e0 * e0;
        if (!(nn > 1e-10 * scale * scale)) ok = false;
        if (ok) {
            // Rayleigh quotient, then the adjugate once more
            const double y0 = N[0][0] * q0 + N[0][1] * q1 + N[0][2] * q2 + N[0][3] * q3;
            const double y1 = N[1][0] * q0 + N[1][1] * q1 + N[1][2] * q2 + N[1][3] * q3;
            const double y2 = N[2][0] * q0 + N[2][1] * q1 + N[2][2] * q2 + N[2][3] * q3;
            const double y3 = N[3][0] * q0 + N[3][1] * q1 + N[3][2] * q2 + N[3][3] * q3;
            const double lam2 = (q0 * y0 + q1 * y1 + q2 * y2 + q3 * y3) / nn;
#pragma unroll
            for (int r = 0; r < 4; ++r) A[r][r] = N[r][r] - lam2;
            double p0, p1, p2, p3;
            cof_row0(A, p0, p1, p2, p3);
            const double n2 = p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3;
            if (n2 > 1e-10 * scale * scale) { q0 = p0; q1 = p1; q2 = p2; q3 = p3; nn = n2; }
            const double nrm = rsqrt(nn);
            q0 *= nrm; q1 *= nrm; q2 *= nrm; q3 *= nrm;
        }
    }
    if (!ok) { horn_rotation(M, Qm); return; }
    Qm[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
    Qm[0][1] = 2.0 * (q1 * q2 - q0 * q3);
    Qm[0][2] = 2.0 * (q1 * q3 + q0 * q2);
    Qm[1][0] = 2.0 * (q1 * q2 + q0 * q3);
    Qm[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
    Qm[1][2] = 2.0 * (q2 * q3 - q0 * q1);
    Qm[2][0] = 2.0 * (q1 * q3 - q0 * q2);
    Qm[2][1] = 2.0 * (q2 * q3 + q0 * q1);
    Qm[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
}

// Column K of adj(A) for symmetric A (= row K of cofactors): adj(N - lambda I) = c q q^T for a simple eigenvalue, so any
// column is the eigenvector times c q_K.
template <int K>
__device__ __forceinline__ void cof_col(const double (&A)[4][4], double &q0, double &q1, double &q2, double &q3) {
    constexpr int r0 = K == 0 ? 1 : 0, r1 = K <= 1 ? 2 : 1, r2 = K <= 2 ? 3 : 2;
    const double m0 = det3(A[r0][1], A[r0][2], A[r0][3], A[r1][1], A[r1][2], A[r1][3], A[r2][1], A[r2][2], A[r2][3]);
    const double m1 = det3(A[r0][0], A[r0][2], A[r0][3], A[r1][0], A[r1][2], A[r1][3], A[r2][0], A[r2][2], A[r2][3]);
    const double m2 = det3(A[r0][0], A[r0][1], A[r0][3], A[r1][0], A[r1][1], A[r1][3], A[r2][0], A[r2][1], A[r2][3]);
    const double m3 = det3(A[r0][0], A[r0][1], A[r0][2], A[r1][0], A[r1][1], A[r1][2], A[r2][0], A[r2][1], A[r2][2]);
    constexpr double sg = (K & 1) ? -1.0 : 1.0;
    q0 = sg * m0; q1 = -sg * m1; q2 = sg * m2; q3 = -sg * m3;
}

__device__ __forceinline__ void cof_col_dyn(const double (&A)[4][4], int k, double &q0, double &q1, double &q2, double &q3) {
    if (k == 0) cof_col<0>(A, q0, q1, q2, q3);
    else if (k == 1) cof_col<1>(A, q0, q1, q2, q3);
    else if (k == 2) cof_col<2>(A, q0, q1, q2, q3);
    else cof_col<3>(A, q0, q1, q2, q3);
}

// horn_rotation_fast with a pivoted adjugate column, for the frame fit (K1): the column of adj(N - lambda I) with the
// largest diagonal entry (c q_K^2, q_K^2 >= 1/4) is used, so the eigenvector keeps full relative accuracy whatever
// the rotation is (column 0 alone loses digits near 180-degree rotations, where q_0 -> 0).  e0 = upper bound of the
// largest eigenvalue ((sum |a|^2 + sum |b|^2) / 2).  Falls back to the Jacobi solver when the largest eigenvalue is
// (nearly) not simple.
__device__ __forceinline__ void horn_rotation_pivot(const double (&M)[3][3], double e0, double (&Qm)[3][3]) {
    const double Sxx = M[0][0], Sxy = M[0][1], Sxz = M[0][2];
    const double Syx = M[1][0], Syy = M[1][1], Syz = M[1][2];
    const double Szx = M[2][0], Szy = M[2][1], Szz = M[2][2];
    double N[4][4];
    N[0][0] = Sxx + Syy + Szz;
    N[0][1] = Syz - Szy; N[0][2] = Szx - Sxz; N[0][3] = Sxy - Syx;
    N[1][1] = Sxx - Syy - Szz; N[1][2] = Sxy + Syx; N[1][3] = Szx + Sxz;
    N[2][2] = -Sxx + Syy - Szz; N[2][3] = Syz + Szy;
    N[3][3] = -Sxx - Syy + Szz;
    N[1][0] = N[0][1]; N[2][0] = N[0][2]; N[3][0] = N[0][3];
    N[2][1] = N[1][2]; N[3][1] = N[1][3]; N[3][2] = N[2][3];
    const double C2 = -2.0 * (Sxx * Sxx + Sxy * Sxy + Sxz * Sxz + Syx * Syx + Syy * Syy + Syz * Syz + Szx * Szx +
                              Szy * Szy + Szz * Szz);
    const double C1 = -8.0 * det3(Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz);
    double c0, c1, c2, c3;
    cof_row0(N, c0, c1, c2, c3);
    const double C0 = N[0][0] * c0 + N[0][1] * c1 + N[0][2] * c2 + N[0][3] * c3;      // det(N)
    double lam = e0;
    bool ok = e0 > 0.0;
    for (int it = 0; it < 40 && ok; ++it) {
        const double x2 = lam * lam;
        const double b = (x2 + C2) * lam;
        const double a = b + C1;
        const double fp = 2.0 * x2 * lam + b + a;
        if (fp == 0.0) { ok = false; break; }
        const double delta = (a * lam + C0) / fp;
        lam -= delta;
        if (fabs(delta) <= 1e-13 * fabs(lam)) break;
    }
    double q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    if (ok) {
        double A[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) A[r][c] = N[r][c] - (r == c ? lam : 0.0);
        // diagonal of the adjugate: c q_k^2
        const double d0 = det3(A[1][1], A[1][2], A[1][3], A[2][1], A[2][2], A[2][3], A[3][1], A[3][2], A[3][3]);
        const double d1 = det3(A[0][0], A[0][2], A[0][3], A[2][0], A[2][2], A[2][3], A[3][0], A[3][2], A[3][3]);
        const double d2 = det3(A[0][0], A[0][1], A[0][3], A[1][0], A[1][1], A[1][3], A[3][0], A[3][1], A[3][3]);
        const double d3 = det3(A[0][0], A[0][1], A[0][2], A[1][0], A[1][1], A[1][2], A[2][0], A[2][1], A[2][2]);
        int k = 0;
        double dm = fabs(d0);
        if (fabs(d1) > dm) { dm = fabs(d1); k = 1; }
        if (fabs(d2) > dm) { dm = fabs(d2); k = 2; }
        if (fabs(d3) > dm) { dm = fabs(d3); k = 3; }
        // c = product of the gaps to the other three eigenvalues; tiny relative to e0^3: the optimum is not unique
        if (!(dm > 1e-7 * e0 * e0 * e0)) ok = false;
        if (ok) {
            cof_col_dyn(A, k, q0, q1, q2, q3);
            double nn = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
            // Rayleigh quotient (quadratically accurate eigenvalue), then the same adjugate column once more
            const double y0 = N[0][0] * q0 + N[0][1] * q1 + N[0][2] * q2 + N[0][3] * q3;
            const double y1 = N[1][0] * q0 + N[1][1] * q1 + N[1][2] * q2 + N[1][3] * q3;
            const double y2 = N[2][0] * q0 + N[2][1] * q1 + N[2][2] * q2 + N[2][3] * q3;
            const double y3 = N[3][0] * q0 + N[3][1] * q1 + N[3][2] * q2 + N[3][3] * q3;
            const double lam2 = (q0 * y0 + q1 * y1 + q2 * y2 + q3 * y3) / nn;
#pragma unroll
            for (int r = 0; r < 4; ++r) A[r][r] = N[r][r] - lam2;
            double p0, p1, p2, p3;
            cof_col_dyn(A, k, p0, p1, p2, p3);
            const double n2 = p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3;
            if (n2 > 0.25 * nn) { q0 = p0; q1 = p1; q2 = p2; q3 = p3; nn = n2; }
            const double nrm = rsqrt(nn);
            q0 *= nrm; q1 *= nrm; q2 *= nrm; q3 *= nrm;
        }
    }
    if (!ok) { horn_rotation(M, Qm); return; }
    Qm[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
    Qm[0][1] = 2.0 * (q1 * q2 - q0 * q3);
    Qm[0][2] = 2.0 * (q1 * q3 + q0 * q2);
    Qm[1][0] = 2.0 * (q1 * q2 + q0 * q3);
    Qm[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
    Qm[1][2] = 2.0 * (q2 * q3 - q0 * q1);
    Qm[2][0] = 2.0 * (q1 * q3 - q0 * q2);
    Qm[2][1] = 2.0 * (q2 * q3 + q0 * q1);
    Qm[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
}

// angle_of_rotation (fr3d/geometry/angleofrotation.py:4-7) given the trace.
__device__ __forceinline__ double angle_from_trace(double tr) {
    double value = (tr - 1.0) / 2.0;
    value = fmin(1.0, fmax(-1.0, value));
    return acos(value);
}

}  // namespace fr3d
