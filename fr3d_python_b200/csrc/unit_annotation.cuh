// Per-nucleotide annotations: glycosidic bond orientation.
//
// Replaces the arithmetic of annotate_bond_orientation (fr3d/classifiers/NA_unit_annotation.py:48-171): the torsion
// chi = O4'-C1'-N9-C4 (purines) / O4'-C1'-N1-C2 (pyrimidines) with the IUPAC sign, and its class.  One thread per
// nt, fp64, the reference's own sequence of operations (two unit normals, their cross product, atan2 of
// (+-|cross|, dot)), so chi agrees with numpy to ~1e-13 degree; the host formats it with "%0.3f" like the reference.
#pragma once

#include "classify.cuh"

namespace fr3d {

struct RingSlots { int32_t s[FR3D_N_PARENTS]; };     // base slot of C4 (purines) / C2 (pyrimidines) per parent

__global__ void __launch_bounds__(128)
bond_orientation_kernel(fr3d_nts_t nts, RingSlots ring, double *__restrict__ chi_degree, int32_t *__restrict__ orientation) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nts.n_nt) return;
    const uint32_t bm = nts.base_mask[n], bb = nts.bb_mask[n];
    const int parent = (int)((bm >> 16) & 15u) - 1;
    double chi = __longlong_as_double(0x7FF8000000000000LL);
    int cls = 0;                                                                    // NA
    if (parent >= 0 && parent < FR3D_N_PARENTS) {
        const int rs = ring.s[parent];
        // N9 / N1 is base slot 0; O4' and C1' are backbone slots 4 and 7 (NA_unit_annotation.py:110-113)
        if ((bm & 1u) && ((bm >> rs) & 1u) && ((bb >> 4) & 1u) && ((bb >> 7) & 1u)) {
            const double *B = nts.base_xyz + (size_t)n * FR3D_BASE_SLOTS * 3;
            const double *K = nts.bb_xyz + (size_t)n * FR3D_BB_SLOTS * 3;
            const double a0 = K[21] - K[12], a1 = K[22] - K[13], a2 = K[23] - K[14];          // O4P_C1P = C1' - O4'
            const double b0 = K[21] - B[0], b1 = K[22] - B[1], b2 = K[23] - B[2];             // N1N9_C1P = C1' - N
            const double c0 = B[0] - B[rs * 3], c1 = B[1] - B[rs * 3 + 1], c2 = B[2] - B[rs * 3 + 2];   // C2C4_N1N9 = N - C
            double s0 = a1 * b2 - a2 * b1, s1 = a2 * b0 - a0 * b2, s2 = a0 * b1 - a1 * b0;    // perp_to_sugar
            const double ns = sqrt(s0 * s0 + s1 * s1 + s2 * s2);
            if (ns != 0.0) { s0 /= ns; s1 /= ns; s2 /= ns; }
            double p0 = (-b1) * c2 - (-b2) * c1, p1 = (-b2) * c0 - (-b0) * c2, p2 = (-b0) * c1 - (-b1) * c0;   // perp_to_base
            const double np_ = sqrt(p0 * p0 + p1 * p1 + p2 * p2);
            p0 /= np_; p1 /= np_; p2 /= np_;
            const double x0 = p1 * s2 - p2 * s1, x1 = p2 * s0 - p0 * s2, x2 = p0 * s1 - p1 * s0;   // cross_cross_chi
            const double cos_chi = s0 * p0 + s1 * p1 + s2 * p2;
            const double nx = sqrt(x0 * x0 + x1 * x1 + x2 * x2);
            const double sin_chi = (x0 * b0 + x1 * b1 + x2 * b2 > 0.0) ? nx : -nx;
            chi = 180.0 * atan2(sin_chi, cos_chi) / 3.141592653589793;
            if (chi > -90.0 && chi < -45.0) cls = 3;                                  // int_syn
            else if (chi >= -45.0 && chi < 90.0) cls = 2;                             // syn
            else cls = 1;                                                             // anti
        }
    }
    chi_degree[n] = chi;
    orientation[n] = cls;
}

}  // namespace fr3d
