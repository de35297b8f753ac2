// sqpnp.cuh -- rows S, S1-S3, U1 of SURVEY.md 8a: batched SqPnP::solve_robot_pose, one warp per problem.
//
// Reference: /root/reference/crates/chalkydri_sqpnp/src/lib.rs (solve_robot_pose :297-377, solve :248-295,
// build_linear_system :124-180, solve_rotation_candidates :396-428, nearest_so3 :42-59, optimization :463-479,
// solve_newton :98-115, constraints_and_jacobian :62-95, compute_std_devs :224-246).  nalgebra's dense kernels
// (symmetric_eigen, 3x3 SVD, LU, try_inverse, Rotation3::from_matrix) are restated with their published
// algorithms: cyclic Jacobi, one-sided Jacobi, partial-pivot LU with reciprocal-pivot multipliers, cofactors,
// Mueller's iterative rotation extraction.
//
// B200 mapping (FP64-issue bound, ~0.3 KB of traffic per problem): everything lives in shared memory / registers
// of one warp.  Accumulations whose value depends on summation order (Omega, the Jacobi convergence sum, Newton
// step norms) are evaluated in the reference's order, element-parallel across lanes, so the eigen-basis picked in
// the rank-deficient single-tag case does not depend on the lane mapping.  The six Newton refinements run two at a
// time, one per half-warp, with the 15x15 KKT system of each in shared memory.
#pragma once
#include "common.cuh"

namespace cb {

constexpr int SQ_MAX_TAGS = 32;          // tags per problem (the reference solves with every visible field tag; an FRC field has 22)
constexpr int SQ_MAX_PTS = SQ_MAX_TAGS * 4;
constexpr int SQ_WARPS = 4;
constexpr int SQ_KP = 17;          // row pitch (doubles) of the 15x15 KKT system: odd, so that column accesses by 16 lanes are bank-conflict free

struct SqWarpShared {
    double omega[81];       // column-major 9x9
    double a[81];           // Jacobi working copy
    double v[81];           // eigenvectors (columns)
    double q_rr[81];
    double q_rt[27];        // column-major 9x3
    double temp[27];
    double q_tt[9], q_tt_inv[9];
    double kkt[2][15 * SQ_KP]; // row-major 15x15 (pitch SQ_KP) per half-warp
    double rhs[2][16];
    double r[2][9];
    double cand_r[6][9];
    double cand_e[6];
    double pw[SQ_MAX_PTS][3];   // world corner points
    double pb[SQ_MAX_PTS][3];   // bearings
};

// what sq_prepare_kernel needs of the above (no KKT systems, no candidates): 9.3 KB per warp
struct SqPrepShared {
    double omega[81], a[81], v[81], q_rr[81];
    double q_rt[27], temp[27];
    double q_tt[9], q_tt_inv[9];
    double pw[SQ_MAX_PTS][3], pb[SQ_MAX_PTS][3];
};

struct SqParams {
    int max_iter;
    double tol_sq;
    double sign_change_error;
};

__device__ __forceinline__ long long total_key(double x)
{
    long long b = __double_as_longlong(x);
    b ^= (long long)(((unsigned long long)(b >> 63)) >> 1);
    return b;
}

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 vscale(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 vcross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Quat { double w, x, y, z; };
__device__ __forceinline__ V3 quat_rotate(const Quat &q, V3 v)
{
    V3 qv = v3(q.x, q.y, q.z);
    V3 t = vscale(vcross(qv, v), 2.0);
    V3 c = vcross(qv, t);
    return vadd(vadd(vscale(t, q.w), c), v);
}
__device__ __forceinline__ Quat quat_mul(const Quat &a, const Quat &b)
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return r;
}
// column-major 3x3: m[c*3 + r]
__device__ __forceinline__ void quat_to_mat(const Quat &q, double *m)
{
    const double i = q.x, j = q.y, k = q.z, w = q.w;
    const double ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    const double ij = i * j * 2, wk = w * k * 2, wj = w * j * 2, ik = i * k * 2, jk = j * k * 2, wi = w * i * 2;
    m[0] = ww + ii - jj - kk; m[3] = ij - wk; m[6] = wj + ik;
    m[1] = wk + ij; m[4] = ww - ii + jj - kk; m[7] = jk - wi;
    m[2] = ik - wj; m[5] = wi + jk; m[8] = ww - ii - jj + kk;
}
#define SQM(m, r, c) (m)[(c) * 3 + (r)]
__device__ __forceinline__ Quat quat_from_mat(const double *m)
{
    const double tr = SQM(m, 0, 0) + SQM(m, 1, 1) + SQM(m, 2, 2);
    Quat q;
    if (tr > 0) {
        const double denom = sqrt(tr + 1.0) * 2.0;
        q.w = 0.25 * denom; q.x = (SQM(m, 2, 1) - SQM(m, 1, 2)) / denom; q.y = (SQM(m, 0, 2) - SQM(m, 2, 0)) / denom; q.z = (SQM(m, 1, 0) - SQM(m, 0, 1)) / denom;
    } else if (SQM(m, 0, 0) > SQM(m, 1, 1) && SQM(m, 0, 0) > SQM(m, 2, 2)) {
        const double denom = sqrt(1.0 + SQM(m, 0, 0) - SQM(m, 1, 1) - SQM(m, 2, 2)) * 2.0;
        q.w = (SQM(m, 2, 1) - SQM(m, 1, 2)) / denom; q.x = 0.25 * denom; q.y = (SQM(m, 0, 1) + SQM(m, 1, 0)) / denom; q.z = (SQM(m, 0, 2) + SQM(m, 2, 0)) / denom;
    } else if (SQM(m, 1, 1) > SQM(m, 2, 2)) {
        const double denom = sqrt(1.0 + SQM(m, 1, 1) - SQM(m, 0, 0) - SQM(m, 2, 2)) * 2.0;
        q.w = (SQM(m, 0, 2) - SQM(m, 2, 0)) / denom; q.x = (SQM(m, 0, 1) + SQM(m, 1, 0)) / denom; q.y = 0.25 * denom; q.z = (SQM(m, 1, 2) + SQM(m, 2, 1)) / denom;
    } else {
        const double denom = sqrt(1.0 + SQM(m, 2, 2) - SQM(m, 0, 0) - SQM(m, 1, 1)) * 2.0;
        q.w = (SQM(m, 1, 0) - SQM(m, 0, 1)) / denom; q.x = (SQM(m, 0, 2) + SQM(m, 2, 0)) / denom; q.y = (SQM(m, 1, 2) + SQM(m, 2, 1)) / denom; q.z = 0.25 * denom;
    }
    return q;
}
__device__ __forceinline__ void mat3_mul(const double *a, const double *b, double *out)
{
    double t[9];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += SQM(a, r, k) * SQM(b, k, c);
            t[c * 3 + r] = acc;
        }
    for (int i = 0; i < 9; i++) out[i] = t[i];
}
__device__ __forceinline__ V3 mat3_mulv(const double *a, V3 v)
{
    return v3(SQM(a, 0, 0) * v.x + SQM(a, 0, 1) * v.y + SQM(a, 0, 2) * v.z,
              SQM(a, 1, 0) * v.x + SQM(a, 1, 1) * v.y + SQM(a, 1, 2) * v.z,
              SQM(a, 2, 0) * v.x + SQM(a, 2, 1) * v.y + SQM(a, 2, 2) * v.z);
}
__device__ __forceinline__ double mat3_det(const double *m)
{
    const double m11 = SQM(m, 0, 0), m12 = SQM(m, 0, 1), m13 = SQM(m, 0, 2);
    const double m21 = SQM(m, 1, 0), m22 = SQM(m, 1, 1), m23 = SQM(m, 1, 2);
    const double m31 = SQM(m, 2, 0), m32 = SQM(m, 2, 1), m33 = SQM(m, 2, 2);
    const double a = m22 * m33 - m32 * m23, b = m21 * m33 - m31 * m23, c = m21 * m32 - m31 * m22;
    return m11 * a - m12 * b + m13 * c;
}
__device__ __forceinline__ bool mat3_try_inverse(const double *m, double *out)
{
    const double m11 = SQM(m, 0, 0), m12 = SQM(m, 0, 1), m13 = SQM(m, 0, 2);
    const double m21 = SQM(m, 1, 0), m22 = SQM(m, 1, 1), m23 = SQM(m, 1, 2);
    const double m31 = SQM(m, 2, 0), m32 = SQM(m, 2, 1), m33 = SQM(m, 2, 2);
    const double minor_m12_m23 = m22 * m33 - m32 * m23;
    const double minor_m11_m23 = m21 * m33 - m31 * m23;
    const double minor_m11_m22 = m21 * m32 - m31 * m22;
    const double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (det == 0) return false;
    SQM(out, 0, 0) = minor_m12_m23 / det;
    SQM(out, 0, 1) = (m13 * m32 - m33 * m12) / det;
    SQM(out, 0, 2) = (m12 * m23 - m22 * m13) / det;
    SQM(out, 1, 0) = -minor_m11_m23 / det;
    SQM(out, 1, 1) = (m11 * m33 - m31 * m13) / det;
    SQM(out, 1, 2) = (m13 * m21 - m23 * m11) / det;
    SQM(out, 2, 0) = minor_m11_m22 / det;
    SQM(out, 2, 1) = (m12 * m31 - m32 * m11) / det;
    SQM(out, 2, 2) = (m11 * m22 - m21 * m12) / det;
    return true;
}

// 3x3 SVD by one-sided Jacobi (scalar; a handful of rotations), singular values descending
__device__ void svd3(const double *m, double *U, double *V)
{
    double a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++) a[i] = m[i];
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int k = 0; k < 3; k++) {
                    alpha += a[p * 3 + k] * a[p * 3 + k];
                    beta += a[q * 3 + k] * a[q * 3 + k];
                    gamma += a[p * 3 + k] * a[q * 3 + k];
                }
                if (gamma == 0 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                if (zeta < 0) t = -t;
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 3; k++) {
                    double x = a[p * 3 + k], y = a[q * 3 + k];
                    a[p * 3 + k] = c * x - s * y;
                    a[q * 3 + k] = s * x + c * y;
                    x = v[p * 3 + k]; y = v[q * 3 + k];
                    v[p * 3 + k] = c * x - s * y;
                    v[q * 3 + k] = s * x + c * y;
                }
            }
        if (!rotated) break;
    }
    double sv[3];
    int ord[3] = {0, 1, 2};
    for (int j = 0; j < 3; j++) sv[j] = sqrt(a[j * 3] * a[j * 3] + a[j * 3 + 1] * a[j * 3 + 1] + a[j * 3 + 2] * a[j * 3 + 2]);
    for (int i = 1; i < 3; i++)
        for (int j = i; j > 0 && sv[ord[j - 1]] < sv[ord[j]]; j--) { const int t = ord[j - 1]; ord[j - 1] = ord[j]; ord[j] = t; }
    int rank = 0;
    for (int j = 0; j < 3; j++) {
        const int o = ord[j];
        for (int k = 0; k < 3; k++) V[j * 3 + k] = v[o * 3 + k];
        if (sv[o] > 1e-300 && sv[o] > 1e-15 * sv[ord[0]]) {
            for (int k = 0; k < 3; k++) U[j * 3 + k] = a[o * 3 + k] / sv[o];
            rank = j + 1;
        } else {
            for (int k = 0; k < 3; k++) U[j * 3 + k] = 0;
        }
    }
    if (rank == 0) { for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? 1 : 0; rank = 3; }
    if (rank == 1) {
        const V3 u0 = v3(U[0], U[1], U[2]);
        int j = 0;
        double best = fabs(U[0]);
        for (int k = 1; k < 3; k++) if (fabs(U[k]) < best) { best = fabs(U[k]); j = k; }
        const V3 e = v3(j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0);
        V3 u1 = vsub(e, vscale(u0, vdot(e, u0)));
        const double nrm = sqrt(vdot(u1, u1));
        u1 = vscale(u1, 1.0 / nrm);
        U[3] = u1.x; U[4] = u1.y; U[5] = u1.z;
        rank = 2;
    }
    if (rank == 2) {
        const V3 u2 = vcross(v3(U[0], U[1], U[2]), v3(U[3], U[4], U[5]));
        U[6] = u2.x; U[7] = u2.y; U[8] = u2.z;
    }
}

// lib.rs:42-59
__device__ void nearest_so3(const double *r_vec, double *out)
{
    double U[9], V[9], Vt[9], rot[9];
    svd3(r_vec, U, V);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) SQM(Vt, r, c) = SQM(V, c, r);
    mat3_mul(U, Vt, rot);
    if (mat3_det(rot) < 0.0) {
        U[6] = -U[6]; U[7] = -U[7]; U[8] = -U[8];
        mat3_mul(U, Vt, rot);
    }
    for (int i = 0; i < 9; i++) out[i] = rot[i];
}

__device__ void axis_angle_mat(V3 u, double ang, double *R)
{
    const double s = sin(ang), c = cos(ang), one_c = 1.0 - c;
    const double ux = u.x, uy = u.y, uz = u.z;
    SQM(R, 0, 0) = ux * ux * one_c + c;      SQM(R, 0, 1) = ux * uy * one_c - uz * s; SQM(R, 0, 2) = ux * uz * one_c + uy * s;
    SQM(R, 1, 0) = ux * uy * one_c + uz * s; SQM(R, 1, 1) = uy * uy * one_c + c;      SQM(R, 1, 2) = uy * uz * one_c - ux * s;
    SQM(R, 2, 0) = ux * uz * one_c - uy * s; SQM(R, 2, 1) = uy * uz * one_c + ux * s; SQM(R, 2, 2) = uz * uz * one_c + c;
}

// Rotation3::from_matrix_eps(m, EPSILON, unlimited, identity) (lib.rs:289,370)
__device__ void rot3_from_matrix(const double *m, double *out)
{
    const double eps = 2.220446049250313e-16;
    double rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int perturb_axis = 0;
    for (int it = 0; it < 10000; it++) {
        V3 axis = v3(0, 0, 0);
        double denom = 0;
        for (int c = 0; c < 3; c++) {
            const V3 rc = v3(rot[c * 3], rot[c * 3 + 1], rot[c * 3 + 2]), mc = v3(m[c * 3], m[c * 3 + 1], m[c * 3 + 2]);
            axis = vadd(axis, vcross(rc, mc));
            denom += vdot(rc, mc);
        }
        const V3 axisangle = vscale(axis, 1.0 / (fabs(denom) + eps));
        const double angle = sqrt(vdot(axisangle, axisangle));
        if (angle > eps) {
            double R[9];
            axis_angle_mat(vscale(axisangle, 1.0 / angle), angle, R);
            mat3_mul(R, rot, rot);
        } else {
            double nsq = 0;
            for (int k = 0; k < 9; k++) nsq += (m[k] - rot[k]) * (m[k] - rot[k]);
            double pert[9];
            for (int k = 0; k < 9; k++) pert[k] = rot[k];
            const V3 ax = v3(perturb_axis == 0 ? 1.0 : 0.0, perturb_axis == 1 ? 1.0 : 0.0, perturb_axis == 2 ? 1.0 : 0.0);
            double R[9], nsq2 = nsq;
            axis_angle_mat(ax, sqrt(eps), R);
            for (int tries = 0; tries < 64; tries++) {
                mat3_mul(pert, R, pert);
                nsq2 = 0;
                for (int k = 0; k < 9; k++) nsq2 += (m[k] - pert[k]) * (m[k] - pert[k]);
                if (fabs(nsq - nsq2) > eps) break;
            }
            if (nsq <= nsq2) break;
            perturb_axis = (perturb_axis + 1) % 3;
            for (int k = 0; k < 9; k++) rot[k] = pert[k];
        }
    }
    for (int k = 0; k < 9; k++) out[k] = rot[k];
}

// Newton refinement of one start by one half-warp (16 lanes, lane hl = KKT row).  lib.rs:463-479 with :62-115.
__device__ double sq_optimize_half(SqWarpShared &S, int half, int hl, uint32_t hmask, const SqParams &prm)
{
    double *M = S.kkt[half], *b = S.rhs[half], *r = S.r[half];
    const double *om = S.omega;
    for (int it = 0; it < prm.max_iter; it++) {
        // ---- build the KKT system: row hl of [Omega J^T; J 0] and rhs = [-Omega r; -h] ----
        if (hl < 15) {
            for (int c = 0; c < 15; c++) M[hl * SQ_KP + c] = 0;
            if (hl < 9) {
                double acc = 0;
                for (int j = 0; j < 9; j++) { const double o = om[j * 9 + hl]; M[hl * SQ_KP + j] = o; acc += o * r[j]; }
                b[hl] = -acc;
            }
        }
        __syncwarp(hmask);
        if (hl < 6) {
            const double *c1 = r, *c2 = r + 3, *c3 = r + 6;
            double hval;
            double jr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            switch (hl) {
                case 0: hval = (c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]) - 1.0; for (int k = 0; k < 3; k++) jr[k] = 2.0 * c1[k]; break;
                case 1: hval = (c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2]) - 1.0; for (int k = 0; k < 3; k++) jr[3 + k] = 2.0 * c2[k]; break;
                case 2: hval = (c3[0] * c3[0] + c3[1] * c3[1] + c3[2] * c3[2]) - 1.0; for (int k = 0; k < 3; k++) jr[6 + k] = 2.0 * c3[k]; break;
                case 3: hval = c1[0] * c2[0] + c1[1] * c2[1] + c1[2] * c2[2]; for (int k = 0; k < 3; k++) { jr[k] = c2[k]; jr[3 + k] = c1[k]; } break;
                case 4: hval = c1[0] * c3[0] + c1[1] * c3[1] + c1[2] * c3[2]; for (int k = 0; k < 3; k++) { jr[k] = c3[k]; jr[6 + k] = c1[k]; } break;
                default: hval = c2[0] * c3[0] + c2[1] * c3[1] + c2[2] * c3[2]; for (int k = 0; k < 3; k++) { jr[3 + k] = c3[k]; jr[6 + k] = c2[k]; } break;
            }
            for (int j = 0; j < 9; j++) { M[(9 + hl) * SQ_KP + j] = jr[j]; M[j * SQ_KP + 9 + hl] = jr[j]; }
            b[9 + hl] = -hval;
        }
        __syncwarp(hmask);
        // ---- LU with partial pivoting (first maximum in row order), reciprocal-pivot multipliers ----
        bool singular = false;
        for (int i = 0; i < 15; i++) {
            double best = (hl >= i && hl < 15) ? fabs(M[hl * SQ_KP + i]) : -1.0;
            int piv = hl;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(hmask, best, o, 16);
                const int op = __shfl_xor_sync(hmask, piv, o, 16);
                if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
            }
            const double diag = M[piv * SQ_KP + i];
            if (diag == 0) continue;
            if (piv != i) {
                if (hl < 15) { const double t = M[i * SQ_KP + hl]; M[i * SQ_KP + hl] = M[piv * SQ_KP + hl]; M[piv * SQ_KP + hl] = t; }
                else { const double t = b[i]; b[i] = b[piv]; b[piv] = t; }
            }
            __syncwarp(hmask);
            const double inv_diag = 1.0 / diag;
            if (hl > i && hl < 15) {
                const double coeff = M[hl * SQ_KP + i] * inv_diag;
                M[hl * SQ_KP + i] = coeff;
                for (int c = i + 1; c < 15; c++) M[hl * SQ_KP + c] -= coeff * M[i * SQ_KP + c];
            }
            __syncwarp(hmask);
        }
        // forward substitution L y = b (unit diagonal), then back substitution U x = y
        for (int i = 0; i < 15; i++) {
            if (hl > i && hl < 15) b[hl] -= M[hl * SQ_KP + i] * b[i];
            __syncwarp(hmask);
        }
        for (int i = 14; i >= 0; i--) {
            const double diag = M[i * SQ_KP + i];
            if (diag == 0) { singular = true; break; }
            if (hl == i) b[i] = b[i] / diag;
            __syncwarp(hmask);
            if (hl < i) b[hl] -= M[hl * SQ_KP + i] * b[i];
            __syncwarp(hmask);
        }
        if (singular) break;
        double nsq = 0;
        for (int k = 0; k < 9; k++) nsq += b[k] * b[k];
        __syncwarp(hmask);
        if (hl < 9) r[hl] += b[hl];
        __syncwarp(hmask);
        if (nsq < prm.tol_sq) break;
    }
    // energy = r . (Omega r)
    double e = 0;
    for (int i = 0; i < 9; i++) {
        double acc = 0;
        for (int j = 0; j < 9; j++) acc += om[j * 9 + i] * r[j];
        e += r[i] * acc;
    }
    return e;
}

__device__ __forceinline__ double quad_form9(const double *om, const double *r)
{
    double tmp[9];
    for (int i = 0; i < 9; i++) {
        double acc = 0;
        for (int j = 0; j < 9; j++) acc += om[j * 9 + i] * r[j];
        tmp[i] = acc;
    }
    double e = 0;
    for (int i = 0; i < 9; i++) e += r[i] * tmp[i];
    return e;
}

__global__ void __launch_bounds__(SQ_WARPS * 32, 4)
sqpnp_kernel(const cb_iso3 *__restrict__ tags, const double *__restrict__ bearings, const int32_t *__restrict__ n_tags, int max_tags,
             const cb_iso3 *__restrict__ robot_to_cam_p, const double *__restrict__ gyro_arr, long long nprob, cb_pose *__restrict__ out,
             uint8_t *__restrict__ ok, SqParams prm)
{
    extern __shared__ __align__(16) unsigned char sq_smem[];
    SqWarpShared &S = reinterpret_cast<SqWarpShared *>(sq_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t full = 0xffffffffu;
    const long long warps_total = (long long)gridDim.x * SQ_WARPS;
    const double XY_STD_DEV_SCALAR = 5.0, THETA_STD_DEV_SCALAR = 2.0, MAX_TRUSTABLE_RMS = 0.1, MAX_GYRO_DELTA = 30.0;
    const double TAG_SIZE = 0.1651, CORNER_DISTANCE = TAG_SIZE / 2.0, PI = 3.14159265358979323846;
    const cb_iso3 r2c_in = *robot_to_cam_p;
    Quat r2c_q; r2c_q.w = r2c_in.q[0]; r2c_q.x = r2c_in.q[1]; r2c_q.y = r2c_in.q[2]; r2c_q.z = r2c_in.q[3];
    const V3 r2c_t = v3(r2c_in.t[0], r2c_in.t[1], r2c_in.t[2]);
    double r2c_m[9];
    quat_to_mat(r2c_q, r2c_m);
    const V3 fwd_in_cam = v3(r2c_m[0], r2c_m[1], r2c_m[2]);

    for (long long prob = (long long)blockIdx.x * SQ_WARPS + (threadIdx.x >> 5); prob < nprob; prob += warps_total) {
        __syncwarp();
        const int nt = n_tags[prob];
        const int n = nt * 4;
        const double gyro = gyro_arr[prob];
        if (nt < 1 || nt > max_tags || nt > SQ_MAX_TAGS) { if (lane == 0) ok[prob] = 0; continue; }   // <3 points -> None
        const cb_iso3 *ptags = tags + prob * max_tags;
        const double *pbear = bearings + prob * max_tags * 12;
        // ---- corner_points_from_center (lib.rs:379-394) ----
        for (int i = lane; i < n; i += 32) {
            const cb_iso3 iso = ptags[i >> 2];
            Quat q; q.w = iso.q[0]; q.x = iso.q[1]; q.y = iso.q[2]; q.z = iso.q[3];
            const int c = i & 3;
            const double Sx = CORNER_DISTANCE;
            const V3 corner = v3(0.0, (c == 0 || c == 3) ? -Sx : Sx, (c < 2) ? -Sx : Sx);
            const V3 p = vadd(quat_rotate(q, corner), v3(iso.t[0], iso.t[1], iso.t[2]));
            S.pw[i][0] = p.x; S.pw[i][1] = p.y; S.pw[i][2] = p.z;
            S.pb[i][0] = pbear[i * 3]; S.pb[i][1] = pbear[i * 3 + 1]; S.pb[i][2] = pbear[i * 3 + 2];
        }
        __syncwarp();
        V3 centroid = v3(0, 0, 0);
        for (int i = 0; i < n; i++) centroid = vadd(centroid, v3(S.pw[i][0], S.pw[i][1], S.pw[i][2]));
        centroid = v3(centroid.x / (double)n, centroid.y / (double)n, centroid.z / (double)n);
        // ---- build_linear_system (lib.rs:124-180): one accumulator per lane-owned entry, points in order ----
        for (int e = lane; e < 117; e += 32) {
            // e in [0,9): q_tt; [9,36): q_rt; [36,117): q_rr
            int kind, r_, c_, a_ = 0, b_ = 0;
            if (e < 9) { kind = 0; r_ = e % 3; c_ = e / 3; }
            else if (e < 36) { kind = 1; const int k = e - 9; const int col = k / 9, row = k % 9; a_ = row / 3; r_ = row % 3; c_ = col; }
            else { kind = 2; const int k = e - 36; const int col = k / 9, row = k % 9; a_ = row / 3; r_ = row % 3; b_ = col / 3; c_ = col % 3; }
            double acc = 0;
            for (int i = 0; i < n; i++) {
                const double vx = S.pb[i][0], vy = S.pb[i][1], vz = S.pb[i][2];
                const double sq_norm = vx * vx + vy * vy + vz * vz;
                const double inv_norm = 1.0 / sq_norm;
                const double vr = r_ == 0 ? vx : (r_ == 1 ? vy : vz), vc = c_ == 0 ? vx : (c_ == 1 ? vy : vz);
                const double P = (r_ == c_ ? 1.0 : 0.0) - (vr * vc) * inv_norm;
                const double X0 = S.pw[i][0] - centroid.x, X1 = S.pw[i][1] - centroid.y, X2 = S.pw[i][2] - centroid.z;
                double add;
                if (kind == 0) add = P;
                else if (kind == 1) add = P * (a_ == 0 ? X0 : (a_ == 1 ? X1 : X2));
                else {
                    const int lo = a_ < b_ ? a_ : b_, hi = a_ < b_ ? b_ : a_;
                    const double Xlo = lo == 0 ? X0 : (lo == 1 ? X1 : X2), Xhi = hi == 0 ? X0 : (hi == 1 ? X1 : X2);
                    add = (P * Xlo) * Xhi;
                }
                acc += add;
            }
            if (kind == 0) S.q_tt[e] = acc;
            else if (kind == 1) S.q_rt[e - 9] = acc;
            else S.q_rr[e - 36] = acc;
        }
        __syncwarp();
        if (lane == 0) {
            if (!mat3_try_inverse(S.q_tt, S.q_tt_inv)) for (int k = 0; k < 9; k++) S.q_tt_inv[k] = 0;
        }
        __syncwarp();
        if (lane < 27) {
            const int c = lane / 9, r_ = lane % 9;
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += S.q_rt[k * 9 + r_] * SQM(S.q_tt_inv, k, c);
            S.temp[c * 9 + r_] = acc;
        }
        __syncwarp();
        for (int e = lane; e < 81; e += 32) {
            const int c = e / 9, r_ = e % 9;
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += S.temp[k * 9 + r_] * S.q_rt[k * 9 + c];
            S.omega[e] = S.q_rr[e] - acc;
        }
        __syncwarp();
        // ---- symmetric eigen: cyclic Jacobi on omega / max|omega| ----
        double amax = 0;
        for (int e = lane; e < 81; e += 32) amax = fmax(amax, fabs(S.omega[e]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(full, amax, o));
        for (int e = lane; e < 81; e += 32) { S.a[e] = amax == 0 ? 0.0 : S.omega[e] / amax; S.v[e] = (e % 10 == 0) ? 1.0 : 0.0; }
        __syncwarp();
        if (amax != 0) {
            for (int sweep = 0; sweep < 40; sweep++) {
                double off = 0;
                for (int q = 1; q < 9; q++)
                    for (int p = 0; p < q; p++) off += S.a[q * 9 + p] * S.a[q * 9 + p];
                if (off <= 1e-34) break;
                for (int p = 0; p < 8; p++)
                    for (int q = p + 1; q < 9; q++) {
                        const double apq = S.a[q * 9 + p];
                        if (apq == 0) continue;
                        const double app = S.a[p * 9 + p], aqq = S.a[q * 9 + q];
                        const double theta = (aqq - app) / (2.0 * apq);
                        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                        if (theta < 0) t = -t;
                        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                        __syncwarp();
                        if (lane < 9) {            // columns p,q of A
                            const double akp = S.a[p * 9 + lane], akq = S.a[q * 9 + lane];
                            S.a[p * 9 + lane] = c * akp - s * akq;
                            S.a[q * 9 + lane] = s * akp + c * akq;
                        } else if (lane < 18) {    // V <- V J
                            const int k = lane - 9;
                            const double vkp = S.v[p * 9 + k], vkq = S.v[q * 9 + k];
                            S.v[p * 9 + k] = c * vkp - s * vkq;
                            S.v[q * 9 + k] = s * vkp + c * vkq;
                        }
                        __syncwarp();
                        if (lane < 9) {            // rows p,q of A
                            const double apk = S.a[lane * 9 + p], aqk = S.a[lane * 9 + q];
                            S.a[lane * 9 + p] = c * apk - s * aqk;
                            S.a[lane * 9 + q] = s * apk + c * aqk;
                        }
                        __syncwarp();
                        if (lane == 0) { S.a[q * 9 + p] = 0; S.a[p * 9 + q] = 0; }
                        __syncwarp();
                    }
            }
        }
        // eigenvalue order: stable sort by f64::total_cmp (lib.rs:400-401)
        int idx[9];
        {
            double ev[9];
            for (int i = 0; i < 9; i++) { ev[i] = S.a[i * 9 + i] * amax; idx[i] = i; }
            for (int i = 1; i < 9; i++)
                for (int j = i; j > 0 && total_key(ev[idx[j]]) < total_key(ev[idx[j - 1]]); j--) { const int t = idx[j - 1]; idx[j - 1] = idx[j]; idx[j] = t; }
        }
        // ---- six starts: nearest_so3(+-e), Newton refinement two at a time (one per half-warp) ----
        const int half = lane >> 4, hl = lane & 15;
        const uint32_t hmask = half ? 0xffff0000u : 0x0000ffffu;
        const double gyro_cos = cos(gyro), gyro_sin = sin(gyro);
        for (int t = 0; t < 3; t++) {
            __syncwarp();
            if (hl == 0) {
                const double sign = half == 0 ? -1.0 : 1.0;
                double guess[9], r0[9];
                for (int k = 0; k < 9; k++) guess[k] = S.v[idx[t] * 9 + k] * sign;
                nearest_so3(guess, r0);
                for (int k = 0; k < 9; k++) S.r[half][k] = r0[k];
            }
            __syncwarp();
            double energy = sq_optimize_half(S, half, hl, hmask, prm);
            __syncwarp();
            if (hl == 0) {
                const double *r = S.r[half];
                const double fx = r[0] * fwd_in_cam.x + r[1] * fwd_in_cam.y + r[2] * fwd_in_cam.z;
                const double fy = r[3] * fwd_in_cam.x + r[4] * fwd_in_cam.y + r[5] * fwd_in_cam.z;
                const double dt = (fx * gyro_cos) + (fy * gyro_sin);
                const double angle_error = fmax(1.0 - dt, 0.0);
                energy += prm.sign_change_error * angle_error;
                const int ci = t * 2 + half;
                for (int k = 0; k < 9; k++) S.cand_r[ci][k] = r[k];
                S.cand_e[ci] = energy;
            }
        }
        __syncwarp();
        // candidates.sort_by(total_cmp) (stable) and the selection loop of solve() (lib.rs:267-294)
        int cord[6] = {0, 1, 2, 3, 4, 5};
        for (int i = 1; i < 6; i++)
            for (int j = i; j > 0 && total_key(S.cand_e[cord[j]]) < total_key(S.cand_e[cord[j - 1]]); j--) { const int t = cord[j - 1]; cord[j - 1] = cord[j]; cord[j] = t; }
        int chosen = -1;
        V3 t_sel = v3(0, 0, 0);
        for (int k = 0; k < 6 && chosen < 0; k++) {
            const double *r = S.cand_r[cord[k]];
            double qr[3];
            for (int c = 0; c < 3; c++) {
                double acc = 0;
                for (int j = 0; j < 9; j++) acc += S.q_rt[c * 9 + j] * r[j];
                qr[c] = acc;
            }
            V3 tl = mat3_mulv(S.q_tt_inv, v3(qr[0], qr[1], qr[2]));
            tl = vscale(tl, -1.0);
            const V3 tt = vsub(tl, mat3_mulv(r, centroid));
            bool front = true;
            for (int i = lane; i < n; i += 32) {
                const V3 pc = vadd(mat3_mulv(r, v3(S.pw[i][0], S.pw[i][1], S.pw[i][2])), tt);
                if (!(pc.z > 0.0)) front = false;
            }
            if (__all_sync(full, front) && S.cand_e[cord[k]] < 1.7976931348623157e308) { chosen = cord[k]; t_sel = tt; }
        }
        if (chosen < 0) { if (lane == 0) ok[prob] = 0; continue; }
        if (lane == 0) {
            const double *r = S.cand_r[chosen];
            const double pure_energy = quad_form9(S.omega, r);
            double rot_w2c[9];
            rot3_from_matrix(r, rot_w2c);
            cb_pose o;
            // compute_std_devs (lib.rs:224-246)
            {
                const double distance = sqrt(vdot(t_sel, t_sel));
                const double n_points = (double)(nt * 4);
                const double rms_error = sqrt(pure_energy / n_points);
                if (rms_error > MAX_TRUSTABLE_RMS) { o.std_devs[0] = o.std_devs[1] = o.std_devs[2] = 1.7976931348623157e308; }
                else {
                    const double distance_multiplier = 1.0 + (distance / TAG_SIZE);
                    const double base_xy_std = rms_error * distance_multiplier;
                    double xy_std = (base_xy_std / sqrt((double)nt)) * XY_STD_DEV_SCALAR;
                    xy_std = fmin(fmax(xy_std, 0.01), 10.0);
                    const double base_theta_std = rms_error / TAG_SIZE;
                    const double val = (base_theta_std * distance_multiplier / sqrt((double)nt)) * THETA_STD_DEV_SCALAR;
                    const double theta_std = fmin(fmax(val, 0.05), PI);
                    o.std_devs[0] = xy_std; o.std_devs[1] = xy_std; o.std_devs[2] = theta_std;
                }
            }
            // world_to_cam^-1 * robot_to_cam (lib.rs:328-337)
            const Quat qw = quat_from_mat(rot_w2c);
            Quat qi; qi.w = qw.w; qi.x = -qw.x; qi.y = -qw.y; qi.z = -qw.z;
            const V3 ti = vscale(quat_rotate(qi, t_sel), -1.0);
            const V3 robot_pos = vadd(quat_rotate(qi, r2c_t), ti);
            const Quat qr_ = quat_mul(qi, r2c_q);
            double robot_rot[9];
            quat_to_mat(qr_, robot_rot);
            V3 tag_centroid = v3(0, 0, 0);
            for (int i = 0; i < nt; i++) tag_centroid = vadd(tag_centroid, v3(ptags[i].t[0], ptags[i].t[1], ptags[i].t[2]));
            tag_centroid = v3(tag_centroid.x / (double)nt, tag_centroid.y / (double)nt, tag_centroid.z / (double)nt);
            const double vision_yaw = atan2(SQM(robot_rot, 1, 0), SQM(robot_rot, 0, 0));
            double delta_yaw = gyro - vision_yaw;
            {
                const double a = delta_yaw + PI, bb = 2.0 * PI;
                double rr = fmod(a, bb);
                if (rr < 0.0) rr += bb;
                delta_yaw = rr - PI;
            }
            const double delta_deg = fabs(delta_yaw) * (180.0 / PI);
            double weight = fmin(fmax(delta_deg / MAX_GYRO_DELTA, 0.0), 1.0);
            weight = weight * weight * (3.0 - 2.0 * weight);
            const double applied = delta_yaw * weight;
            const double cos_dt = cos(applied), sin_dt = sin(applied);
            double rot_z[9];
            SQM(rot_z, 0, 0) = cos_dt; SQM(rot_z, 0, 1) = -sin_dt; SQM(rot_z, 0, 2) = 0;
            SQM(rot_z, 1, 0) = sin_dt; SQM(rot_z, 1, 1) = cos_dt;  SQM(rot_z, 1, 2) = 0;
            SQM(rot_z, 2, 0) = 0;      SQM(rot_z, 2, 1) = 0;       SQM(rot_z, 2, 2) = 1;
            double rot_z_rot3[9];
            rot3_from_matrix(rot_z, rot_z_rot3);
            const V3 rel = vsub(robot_pos, tag_centroid);
            const V3 piv = vadd(tag_centroid, mat3_mulv(rot_z, rel));
            mat3_mul(rot_z_rot3, robot_rot, o.rot);
            o.pos[0] = piv.x; o.pos[1] = piv.y; o.pos[2] = piv.z;
            out[prob] = o;
            ok[prob] = 1;
        }
    }
}

// ---- three-phase form (default since round 2) ----------------------------------------------------------------------------
// ncu of the one-kernel form above: 185 k warp instructions per problem, 70 % of them in the 15x15 LU and the two triangular
// solves of the Newton refinement, which a half-warp executes with one row per lane -- pivot search by shuffles, two
// __syncwarp per pivot step, a division on one lane while fifteen wait -- and everything that follows the refinement runs on
// lane 0 alone.  The arithmetic is a few hundred flops per system; the rest is the price of spreading ONE small dense
// system over lanes.  So the solve is split where the parallel shape changes, with the intermediates (1.4 KB per problem) in
// global memory:
//   sq_prepare_kernel   one WARP per problem (as before): corner points, Omega with lane-owned entries accumulated in point
//                       order, cyclic Jacobi, eigen order; the six nearest_so3 starts on six lanes at once;
//   sq_newton_kernel    one THREAD per (problem, start): the 15x15 KKT system of the thread lives in shared memory in
//                       [element][lane] layout (bank = lane: conflict-free whatever row a thread's pivoting visits), the
//                       pivot row of a step is cached in registers, no shuffles, no barriers, a division costs one warp
//                       instruction for 32 systems.  Every element sees the operations of the reference's LU
//                       (lib.rs:98-115 on nalgebra's partial-pivot LU) in the same order, so the results are bit for bit
//                       those of the one-kernel form;
//   sq_finish_kernel    one THREAD per problem: gyro penalty, candidate order, cheirality test, pose and std-devs.
constexpr int SQ_NEWTON_SMEM = 240 * 32 * (int)sizeof(double);      // 225 matrix entries + 15 right-hand side, 32 systems
struct SqScratch {          // per problem, structure of arrays; r: the six starts in, the six refined rotations out
    double *omega, *q_rt, *q_tt_inv, *centroid, *r, *energy;
    uint8_t *valid;
};

__global__ void __launch_bounds__(SQ_WARPS * 32)
sq_prepare_kernel(const cb_iso3 *__restrict__ tags, const double *__restrict__ bearings, const int32_t *__restrict__ n_tags, int max_tags,
                  long long nprob, SqScratch sc)
{
    extern __shared__ __align__(16) unsigned char sq_smem[];
    SqPrepShared &S = reinterpret_cast<SqPrepShared *>(sq_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t full = 0xffffffffu;
    const long long warps_total = (long long)gridDim.x * SQ_WARPS;
    const double TAG_SIZE = 0.1651, CORNER_DISTANCE = TAG_SIZE / 2.0;
    for (long long prob = (long long)blockIdx.x * SQ_WARPS + (threadIdx.x >> 5); prob < nprob; prob += warps_total) {
        __syncwarp();
        const int nt = n_tags[prob];
        const int n = nt * 4;
        if (nt < 1 || nt > max_tags || nt > SQ_MAX_TAGS) { if (lane == 0) sc.valid[prob] = 0; continue; }   // <3 points -> None
        const cb_iso3 *ptags = tags + prob * max_tags;
        const double *pbear = bearings + prob * max_tags * 12;
        // ---- corner_points_from_center (lib.rs:379-394) ----
        for (int i = lane; i < n; i += 32) {
            const cb_iso3 iso = ptags[i >> 2];
            Quat q; q.w = iso.q[0]; q.x = iso.q[1]; q.y = iso.q[2]; q.z = iso.q[3];
            const int c = i & 3;
            const double Sx = CORNER_DISTANCE;
            const V3 corner = v3(0.0, (c == 0 || c == 3) ? -Sx : Sx, (c < 2) ? -Sx : Sx);
            const V3 p = vadd(quat_rotate(q, corner), v3(iso.t[0], iso.t[1], iso.t[2]));
            S.pw[i][0] = p.x; S.pw[i][1] = p.y; S.pw[i][2] = p.z;
            S.pb[i][0] = pbear[i * 3]; S.pb[i][1] = pbear[i * 3 + 1]; S.pb[i][2] = pbear[i * 3 + 2];
        }
        __syncwarp();
        V3 centroid = v3(0, 0, 0);
        for (int i = 0; i < n; i++) centroid = vadd(centroid, v3(S.pw[i][0], S.pw[i][1], S.pw[i][2]));
        centroid = v3(centroid.x / (double)n, centroid.y / (double)n, centroid.z / (double)n);
        // ---- build_linear_system (lib.rs:124-180): one accumulator per lane-owned entry, points in order ----
        for (int e = lane; e < 117; e += 32) {
            int kind, r_, c_, a_ = 0, b_ = 0;
            if (e < 9) { kind = 0; r_ = e % 3; c_ = e / 3; }
            else if (e < 36) { kind = 1; const int k = e - 9; const int col = k / 9, row = k % 9; a_ = row / 3; r_ = row % 3; c_ = col; }
            else { kind = 2; const int k = e - 36; const int col = k / 9, row = k % 9; a_ = row / 3; r_ = row % 3; b_ = col / 3; c_ = col % 3; }
            double acc = 0;
            for (int i = 0; i < n; i++) {
                const double vx = S.pb[i][0], vy = S.pb[i][1], vz = S.pb[i][2];
                const double sq_norm = vx * vx + vy * vy + vz * vz;
                const double inv_norm = 1.0 / sq_norm;
                const double vr = r_ == 0 ? vx : (r_ == 1 ? vy : vz), vc = c_ == 0 ? vx : (c_ == 1 ? vy : vz);
                const double P = (r_ == c_ ? 1.0 : 0.0) - (vr * vc) * inv_norm;
                const double X0 = S.pw[i][0] - centroid.x, X1 = S.pw[i][1] - centroid.y, X2 = S.pw[i][2] - centroid.z;
                double add;
                if (kind == 0) add = P;
                else if (kind == 1) add = P * (a_ == 0 ? X0 : (a_ == 1 ? X1 : X2));
                else {
                    const int lo = a_ < b_ ? a_ : b_, hi = a_ < b_ ? b_ : a_;
                    const double Xlo = lo == 0 ? X0 : (lo == 1 ? X1 : X2), Xhi = hi == 0 ? X0 : (hi == 1 ? X1 : X2);
                    add = (P * Xlo) * Xhi;
                }
                acc += add;
            }
            if (kind == 0) S.q_tt[e] = acc;
            else if (kind == 1) S.q_rt[e - 9] = acc;
            else S.q_rr[e - 36] = acc;
        }
        __syncwarp();
        if (lane == 0) {
            if (!mat3_try_inverse(S.q_tt, S.q_tt_inv)) for (int k = 0; k < 9; k++) S.q_tt_inv[k] = 0;
        }
        __syncwarp();
        if (lane < 27) {
            const int c = lane / 9, r_ = lane % 9;
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += S.q_rt[k * 9 + r_] * SQM(S.q_tt_inv, k, c);
            S.temp[c * 9 + r_] = acc;
        }
        __syncwarp();
        for (int e = lane; e < 81; e += 32) {
            const int c = e / 9, r_ = e % 9;
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += S.temp[k * 9 + r_] * S.q_rt[k * 9 + c];
            S.omega[e] = S.q_rr[e] - acc;
        }
        __syncwarp();
        // what the later phases read
        for (int e = lane; e < 81; e += 32) sc.omega[prob * 81 + e] = S.omega[e];
        if (lane < 27) sc.q_rt[prob * 27 + lane] = S.q_rt[lane];
        if (lane < 9) sc.q_tt_inv[prob * 9 + lane] = S.q_tt_inv[lane];
        if (lane < 3) sc.centroid[prob * 3 + lane] = lane == 0 ? centroid.x : (lane == 1 ? centroid.y : centroid.z);
        if (lane == 0) sc.valid[prob] = 1;
        // ---- symmetric eigen: cyclic Jacobi on omega / max|omega| ----
        double amax = 0;
        for (int e = lane; e < 81; e += 32) amax = fmax(amax, fabs(S.omega[e]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(full, amax, o));
        for (int e = lane; e < 81; e += 32) { S.a[e] = amax == 0 ? 0.0 : S.omega[e] / amax; S.v[e] = (e % 10 == 0) ? 1.0 : 0.0; }
        __syncwarp();
        if (amax != 0) {
            for (int sweep = 0; sweep < 40; sweep++) {
                double off = 0;
                for (int q = 1; q < 9; q++)
                    for (int p = 0; p < q; p++) off += S.a[q * 9 + p] * S.a[q * 9 + p];
                if (off <= 1e-34) break;
                for (int p = 0; p < 8; p++)
                    for (int q = p + 1; q < 9; q++) {
                        const double apq = S.a[q * 9 + p];
                        if (apq == 0) continue;
                        const double app = S.a[p * 9 + p], aqq = S.a[q * 9 + q];
                        const double theta = (aqq - app) / (2.0 * apq);
                        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                        if (theta < 0) t = -t;
                        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                        __syncwarp();
                        if (lane < 9) {            // columns p,q of A
                            const double akp = S.a[p * 9 + lane], akq = S.a[q * 9 + lane];
                            S.a[p * 9 + lane] = c * akp - s * akq;
                            S.a[q * 9 + lane] = s * akp + c * akq;
                        } else if (lane < 18) {    // V <- V J
                            const int k = lane - 9;
                            const double vkp = S.v[p * 9 + k], vkq = S.v[q * 9 + k];
                            S.v[p * 9 + k] = c * vkp - s * vkq;
                            S.v[q * 9 + k] = s * vkp + c * vkq;
                        }
                        __syncwarp();
                        if (lane < 9) {            // rows p,q of A
                            const double apk = S.a[lane * 9 + p], aqk = S.a[lane * 9 + q];
                            S.a[lane * 9 + p] = c * apk - s * aqk;
                            S.a[lane * 9 + q] = s * apk + c * aqk;
                        }
                        __syncwarp();
                        if (lane == 0) { S.a[q * 9 + p] = 0; S.a[p * 9 + q] = 0; }
                        __syncwarp();
                    }
            }
        }
        // eigenvalue order: stable sort by f64::total_cmp (lib.rs:400-401)
        int idx[9];
        {
            double ev[9];
            for (int i = 0; i < 9; i++) { ev[i] = S.a[i * 9 + i] * amax; idx[i] = i; }
            for (int i = 1; i < 9; i++)
                for (int j = i; j > 0 && total_key(ev[idx[j]]) < total_key(ev[idx[j - 1]]); j--) { const int t = idx[j - 1]; idx[j - 1] = idx[j]; idx[j] = t; }
        }
        // ---- six starts nearest_so3(+-e_t), t = 0..2 (lib.rs:402-428): candidate ci = 2 t + (sign > 0) on lane ci ----
        if (lane < 6) {
            const int t = lane >> 1;
            const double sign = (lane & 1) == 0 ? -1.0 : 1.0;
            const int col = t == 0 ? idx[0] : (t == 1 ? idx[1] : idx[2]);
            double guess[9], r0[9];
            for (int k = 0; k < 9; k++) guess[k] = S.v[col * 9 + k] * sign;
            nearest_so3(guess, r0);
            for (int k = 0; k < 9; k++) sc.r[(prob * 6 + lane) * 9 + k] = r0[k];
        }
    }
}

#define SQK(r, c) kkt[((r) * 15 + (c)) * 32]
#define SQB(r) kkt[(225 + (r)) * 32]
// one pivot step of the LU, I a compile-time constant so that the cached pivot row is indexed statically
template <int I, int U>
__device__ __forceinline__ void sq_lu_step(double *kkt)
{
    // partial pivoting: first maximum of |M[r][I]|, r >= I, in row order
    double best = fabs(SQK(I, I));
    int piv = I;
#pragma unroll
    for (int r = I + 1; r < 15; r++) {
        const double v = fabs(SQK(r, I));
        if (v > best) { best = v; piv = r; }
    }
    const double diag = SQK(piv, I);
    if (diag == 0) return;
    if (piv != I) {
#pragma unroll
        for (int c = 0; c < 15; c++) { const double t = SQK(I, c); SQK(I, c) = SQK(piv, c); SQK(piv, c) = t; }
        const double t = SQB(I); SQB(I) = SQB(piv); SQB(piv) = t;
    }
    const double inv_diag = 1.0 / diag;
    double prow[15];
#pragma unroll
    for (int c = I + 1; c < 15; c++) prow[c] = SQK(I, c);
#pragma unroll U
    for (int r = I + 1; r < 15; r++) {
        const double coeff = SQK(r, I) * inv_diag;
        SQK(r, I) = coeff;
#pragma unroll
        for (int c = I + 1; c < 15; c++) SQK(r, c) -= coeff * prow[c];
    }
}

template <int U>
__global__ void __launch_bounds__(32)
sq_newton_kernel(SqScratch sc, long long nsys, SqParams prm)
{
    extern __shared__ __align__(16) unsigned char sq_smem[];
    double *kkt = reinterpret_cast<double *>(sq_smem) + threadIdx.x;
    const long long sys = (long long)blockIdx.x * 32 + threadIdx.x;
    if (sys >= nsys) return;
    const long long prob = sys / 6;
    if (!sc.valid[prob]) return;
    const double *__restrict__ om = sc.omega + prob * 81;
    double r[9];
#pragma unroll
    for (int k = 0; k < 9; k++) r[k] = sc.r[sys * 9 + k];
    for (int it = 0; it < prm.max_iter; it++) {
        // ---- the KKT system [Omega J^T; J 0] [delta; lambda] = [-Omega r; -h] (lib.rs:62-115) ----
#pragma unroll
        for (int row = 0; row < 9; row++) {
            double acc = 0;
#pragma unroll
            for (int j = 0; j < 9; j++) { const double o = __ldg(om + j * 9 + row); SQK(row, j) = o; acc += o * r[j]; }
            SQB(row) = -acc;
        }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            // constraint k: |c1|^2 - 1, |c2|^2 - 1, |c3|^2 - 1, c1.c2, c1.c3, c2.c3 and its gradient row
            const int ca = k < 3 ? k : (k == 5 ? 1 : 0), cb_ = k < 3 ? k : (k == 3 ? 1 : 2);
            double jr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            double hval;
            if (k < 3) {
                hval = (r[3 * ca] * r[3 * ca] + r[3 * ca + 1] * r[3 * ca + 1] + r[3 * ca + 2] * r[3 * ca + 2]) - 1.0;
#pragma unroll
                for (int m = 0; m < 3; m++) jr[3 * ca + m] = 2.0 * r[3 * ca + m];
            } else {
                hval = r[3 * ca] * r[3 * cb_] + r[3 * ca + 1] * r[3 * cb_ + 1] + r[3 * ca + 2] * r[3 * cb_ + 2];
#pragma unroll
                for (int m = 0; m < 3; m++) { jr[3 * ca + m] = r[3 * cb_ + m]; jr[3 * cb_ + m] = r[3 * ca + m]; }
            }
#pragma unroll
            for (int j = 0; j < 9; j++) { SQK(9 + k, j) = jr[j]; SQK(j, 9 + k) = jr[j]; }
#pragma unroll
            for (int j = 9; j < 15; j++) SQK(9 + k, j) = 0;
            SQB(9 + k) = -hval;
        }
        // ---- LU with partial pivoting, reciprocal-pivot multipliers ----
        sq_lu_step<0, U>(kkt); sq_lu_step<1, U>(kkt); sq_lu_step<2, U>(kkt); sq_lu_step<3, U>(kkt); sq_lu_step<4, U>(kkt);
        sq_lu_step<5, U>(kkt); sq_lu_step<6, U>(kkt); sq_lu_step<7, U>(kkt); sq_lu_step<8, U>(kkt); sq_lu_step<9, U>(kkt);
        sq_lu_step<10, U>(kkt); sq_lu_step<11, U>(kkt); sq_lu_step<12, U>(kkt); sq_lu_step<13, U>(kkt); sq_lu_step<14, U>(kkt);
        // forward substitution L y = b (unit diagonal), then back substitution U x = y
        double b[15];
#pragma unroll
        for (int i = 0; i < 15; i++) b[i] = SQB(i);
#pragma unroll
        for (int i = 0; i < 15; i++)
#pragma unroll
            for (int rr = i + 1; rr < 15; rr++) b[rr] -= SQK(rr, i) * b[i];
        bool singular = false;
#pragma unroll
        for (int i = 14; i >= 0; i--) {
            if (!singular) {
                const double diag = SQK(i, i);
                if (diag == 0) singular = true;
                else {
                    b[i] = b[i] / diag;
#pragma unroll
                    for (int rr = 0; rr < i; rr++) b[rr] -= SQK(rr, i) * b[i];
                }
            }
        }
        if (singular) break;
        double nsq = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) nsq += b[k] * b[k];
#pragma unroll
        for (int k = 0; k < 9; k++) r[k] += b[k];
        if (nsq < prm.tol_sq) break;
    }
    // energy = r . (Omega r)
    double e = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        double acc = 0;
#pragma unroll
        for (int j = 0; j < 9; j++) acc += __ldg(om + j * 9 + i) * r[j];
        e += r[i] * acc;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) sc.r[sys * 9 + k] = r[k];
    sc.energy[sys] = e;
}
#undef SQK
#undef SQB

// ---- Newton refinement for a handful of problems (the reference's call shape: one frame -> one problem, lib.rs:293-379) ----------
// With a thread per system one problem is six lanes of one warp running 15 iterations of a scalar 15x15 LU: 223 us, more than half of
// the pose chain of a single frame.  Here a CTA of two warps takes one problem and every system gets SQS_G = 8 lanes (three systems
// per warp).  The augmented system [K | b] (15 x 16) lives in shared memory; lane m of a group owns the columns m and m + 8 (the
// right-hand side is column 15).  A pivot step is two phases separated by __syncwarp: (A) every lane of the group reads column I,
// finds the pivot and forms the 14 - I multipliers itself (redundant, no broadcast); (B) it swaps and updates its own columns --
// with uniform control flow: the lanes of a warp own different columns, so a branch on the column would serialise them; invalid
// columns are computed on a clamped address and not stored.  The forward substitution is folded into the elimination (the
// right-hand side is just another column: b[r] -= l[r][i] * b[i] happens with the same operands in the same order as in L y = b
// afterwards), so L is never stored; the back substitution runs redundantly on every lane.  Each element sees the operations of
// sq_newton_kernel in the same order: bit for bit the same rotations
// (tests/test_gpu_sqpnp.py::test_small_batch_newton_is_bit_identical).
constexpr int SQS_G = 8, SQS_PER_WARP = 3, SQS_PITCH = 248;           // pitch (doubles) between the systems of a warp: 16 banks apart
template <int I>
__device__ __forceinline__ void sqs_lu_step(double *__restrict__ M, int m, bool active, bool &singular)
{
    double cf[15];                                   // column I, rows I .. 14; then the multipliers of the rows below the pivot row
    int piv = I;
    // (finished groups and the idle lanes run along on their own system without storing anything: no branch, no divergence)
#pragma unroll
    for (int r = I; r < 15; r++) cf[r] = M[r * 16 + I];
    // partial pivoting: first maximum of |K[r][I]|, r >= I, in row order
    double best = fabs(cf[I]), diag = cf[I];
#pragma unroll
    for (int r = I + 1; r < 15; r++) {
        const double v = fabs(cf[r]);
        const bool gt = v > best;
        best = gt ? v : best; piv = gt ? r : piv; diag = gt ? cf[r] : diag;
    }
    if (active && diag == 0) singular = true;
    const bool go = active && diag != 0;
    __syncwarp();
    {
        const double inv_diag = 1.0 / diag;
        const double old_top = cf[I];
#pragma unroll
        for (int r = I + 1; r < 15; r++) cf[r] = (r == piv ? old_top : cf[r]) * inv_diag;      // rows after the swap
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (8 * j + 7 < I) continue;             // (compile time: these columns hold L, which nothing reads)
            const int c = m + SQS_G * j;
            const bool valid = go && c >= I;
            const int cc = valid ? c : 15;
            const double top_old = M[I * 16 + cc], top = M[piv * 16 + cc];      // the pivot row after the swap: `top`
            double v[15];
#pragma unroll
            for (int r = I + 1; r < 15; r++) v[r] = M[r * 16 + cc];
#pragma unroll
            for (int r = I + 1; r < 15; r++) v[r] = (r == piv ? top_old : v[r]) - cf[r] * top;
            if (valid) {
                M[I * 16 + cc] = top;                // (column I: U[I][I] = diag)
                if (c > I) {
#pragma unroll
                    for (int r = I + 1; r < 15; r++) M[r * 16 + cc] = v[r];
                }
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(64)
sq_newton_small_kernel(SqScratch sc, long long nprob, SqParams prm)
{
    __shared__ __align__(16) double s_sys[2 * SQS_PER_WARP * SQS_PITCH];
    const long long prob = blockIdx.x;
    if (prob >= nprob || !sc.valid[prob]) return;
    const uint32_t full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, grp = min(lane / SQS_G, SQS_PER_WARP - 1), m = lane % SQS_G;
    const long long sys = prob * 6 + wid * SQS_PER_WARP + grp;
    double *M = s_sys + (wid * SQS_PER_WARP + grp) * SQS_PITCH;
    const double *__restrict__ om = sc.omega + prob * 81;
    bool done = lane >= SQS_PER_WARP * SQS_G;        // lanes 24 .. 31 only keep the barriers company
    double r[9];
#pragma unroll
    for (int k = 0; k < 9; k++) r[k] = sc.r[sys * 9 + k];
    for (int it = 0; it < prm.max_iter; it++) {
        if (__all_sync(full, done)) break;
        const bool active = !done;
        if (active) {
            // ---- the KKT system [Omega J^T; J 0] [delta; lambda] = [-Omega r; -h] (lib.rs:62-115): lane m fills row m of the Omega
            //      block (lane 0 also row 8) and, for m < 6, constraint m ----
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int row = m + SQS_G * h;
                if (row < 9) {
                    double acc = 0;
#pragma unroll
                    for (int j = 0; j < 9; j++) { const double o = __ldg(om + j * 9 + row); M[row * 16 + j] = o; acc += o * r[j]; }
                    M[row * 16 + 15] = -acc;
                }
            }
#pragma unroll
            for (int k = 0; k < 6; k++) {
                if (k != m) continue;
                const int ca = k < 3 ? k : (k == 5 ? 1 : 0), cb_ = k < 3 ? k : (k == 3 ? 1 : 2);
                double jr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                double hval;
                if (k < 3) {
                    hval = (r[3 * ca] * r[3 * ca] + r[3 * ca + 1] * r[3 * ca + 1] + r[3 * ca + 2] * r[3 * ca + 2]) - 1.0;
#pragma unroll
                    for (int q = 0; q < 3; q++) jr[3 * ca + q] = 2.0 * r[3 * ca + q];
                } else {
                    hval = r[3 * ca] * r[3 * cb_] + r[3 * ca + 1] * r[3 * cb_ + 1] + r[3 * ca + 2] * r[3 * cb_ + 2];
#pragma unroll
                    for (int q = 0; q < 3; q++) { jr[3 * ca + q] = r[3 * cb_ + q]; jr[3 * cb_ + q] = r[3 * ca + q]; }
                }
#pragma unroll
                for (int j = 0; j < 9; j++) { M[(9 + k) * 16 + j] = jr[j]; M[j * 16 + 9 + k] = jr[j]; }
#pragma unroll
                for (int j = 9; j < 15; j++) M[(9 + k) * 16 + j] = 0;
                M[(9 + k) * 16 + 15] = -hval;
            }
        }
        __syncwarp();
        bool singular = false;
        sqs_lu_step<0>(M, m, active, singular); sqs_lu_step<1>(M, m, active, singular); sqs_lu_step<2>(M, m, active, singular);
        sqs_lu_step<3>(M, m, active, singular); sqs_lu_step<4>(M, m, active, singular); sqs_lu_step<5>(M, m, active, singular);
        sqs_lu_step<6>(M, m, active, singular); sqs_lu_step<7>(M, m, active, singular); sqs_lu_step<8>(M, m, active, singular);
        sqs_lu_step<9>(M, m, active, singular); sqs_lu_step<10>(M, m, active, singular); sqs_lu_step<11>(M, m, active, singular);
        sqs_lu_step<12>(M, m, active, singular); sqs_lu_step<13>(M, m, active, singular); sqs_lu_step<14>(M, m, active, singular);
        if (active) {
            if (singular) done = true;
            else {
                // back substitution U x = y (y: the eliminated right-hand side), every lane of the group for itself
                double b[15];
#pragma unroll
                for (int i = 0; i < 15; i++) b[i] = M[i * 16 + 15];
#pragma unroll
                for (int i = 14; i >= 0; i--) {
                    b[i] = b[i] / M[i * 16 + i];
#pragma unroll
                    for (int rr = 0; rr < i; rr++) b[rr] -= M[rr * 16 + i] * b[i];
                }
                double nsq = 0;
#pragma unroll
                for (int k = 0; k < 9; k++) nsq += b[k] * b[k];
#pragma unroll
                for (int k = 0; k < 9; k++) r[k] += b[k];
                if (nsq < prm.tol_sq) done = true;
            }
        }
        __syncwarp();                                // the next iteration's fill overwrites what the slower lanes still read
    }
    if (lane < SQS_PER_WARP * SQS_G && m == 0) {
        // energy = r . (Omega r)
        double e = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            double acc = 0;
#pragma unroll
            for (int j = 0; j < 9; j++) acc += __ldg(om + j * 9 + i) * r[j];
            e += r[i] * acc;
        }
#pragma unroll
        for (int k = 0; k < 9; k++) sc.r[sys * 9 + k] = r[k];
        sc.energy[sys] = e;
    }
}

__global__ void __launch_bounds__(128)
sq_finish_kernel(const cb_iso3 *__restrict__ tags, const int32_t *__restrict__ n_tags, int max_tags, const cb_iso3 *__restrict__ robot_to_cam_p,
                 const double *__restrict__ gyro_arr, long long nprob, SqScratch sc, cb_pose *__restrict__ out, uint8_t *__restrict__ ok, SqParams prm)
{
    const long long prob = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= nprob) return;
    if (!sc.valid[prob]) { ok[prob] = 0; return; }
    const double XY_STD_DEV_SCALAR = 5.0, THETA_STD_DEV_SCALAR = 2.0, MAX_TRUSTABLE_RMS = 0.1, MAX_GYRO_DELTA = 30.0;
    const double TAG_SIZE = 0.1651, CORNER_DISTANCE = TAG_SIZE / 2.0, PI = 3.14159265358979323846;
    const cb_iso3 r2c_in = *robot_to_cam_p;
    Quat r2c_q; r2c_q.w = r2c_in.q[0]; r2c_q.x = r2c_in.q[1]; r2c_q.y = r2c_in.q[2]; r2c_q.z = r2c_in.q[3];
    const V3 r2c_t = v3(r2c_in.t[0], r2c_in.t[1], r2c_in.t[2]);
    double r2c_m[9];
    quat_to_mat(r2c_q, r2c_m);
    const V3 fwd_in_cam = v3(r2c_m[0], r2c_m[1], r2c_m[2]);
    const int nt = n_tags[prob];
    const double gyro = gyro_arr[prob];
    const cb_iso3 *ptags = tags + prob * max_tags;
    const double *q_rt = sc.q_rt + prob * 27, *q_tt_inv = sc.q_tt_inv + prob * 9, *om = sc.omega + prob * 81;
    const V3 centroid = v3(sc.centroid[prob * 3], sc.centroid[prob * 3 + 1], sc.centroid[prob * 3 + 2]);
    const double gyro_cos = cos(gyro), gyro_sin = sin(gyro);
    // energies with the gyro penalty (lib.rs:414-424), candidates.sort_by(total_cmp) (stable), selection loop of solve() (lib.rs:267-294)
    double cand_e[6];
    int cord[6] = {0, 1, 2, 3, 4, 5};
    for (int ci = 0; ci < 6; ci++) {
        const double *r = sc.r + (prob * 6 + ci) * 9;
        double energy = sc.energy[prob * 6 + ci];
        const double fx = r[0] * fwd_in_cam.x + r[1] * fwd_in_cam.y + r[2] * fwd_in_cam.z;
        const double fy = r[3] * fwd_in_cam.x + r[4] * fwd_in_cam.y + r[5] * fwd_in_cam.z;
        const double dt = (fx * gyro_cos) + (fy * gyro_sin);
        const double angle_error = fmax(1.0 - dt, 0.0);
        energy += prm.sign_change_error * angle_error;
        cand_e[ci] = energy;
    }
    for (int i = 1; i < 6; i++)
        for (int j = i; j > 0 && total_key(cand_e[cord[j]]) < total_key(cand_e[cord[j - 1]]); j--) { const int t = cord[j - 1]; cord[j - 1] = cord[j]; cord[j] = t; }
    int chosen = -1;
    V3 t_sel = v3(0, 0, 0);
    double rsel[9];
    for (int k = 0; k < 6 && chosen < 0; k++) {
        double r[9];
        for (int j = 0; j < 9; j++) r[j] = sc.r[(prob * 6 + cord[k]) * 9 + j];
        double qr[3];
        for (int c = 0; c < 3; c++) {
            double acc = 0;
            for (int j = 0; j < 9; j++) acc += q_rt[c * 9 + j] * r[j];
            qr[c] = acc;
        }
        V3 tl = mat3_mulv(q_tt_inv, v3(qr[0], qr[1], qr[2]));
        tl = vscale(tl, -1.0);
        const V3 tt = vsub(tl, mat3_mulv(r, centroid));
        bool front = true;
        for (int i = 0; i < nt * 4 && front; i++) {
            const cb_iso3 iso = ptags[i >> 2];
            Quat q; q.w = iso.q[0]; q.x = iso.q[1]; q.y = iso.q[2]; q.z = iso.q[3];
            const int c = i & 3;
            const double Sx = CORNER_DISTANCE;
            const V3 corner = v3(0.0, (c == 0 || c == 3) ? -Sx : Sx, (c < 2) ? -Sx : Sx);
            const V3 pw = vadd(quat_rotate(q, corner), v3(iso.t[0], iso.t[1], iso.t[2]));
            const V3 pc = vadd(mat3_mulv(r, pw), tt);
            if (!(pc.z > 0.0)) front = false;
        }
        if (front && cand_e[cord[k]] < 1.7976931348623157e308) {
            chosen = cord[k]; t_sel = tt;
            for (int j = 0; j < 9; j++) rsel[j] = r[j];
        }
    }
    if (chosen < 0) { ok[prob] = 0; return; }
    {
        const double *r = rsel;
        const double pure_energy = quad_form9(om, r);
        double rot_w2c[9];
        rot3_from_matrix(r, rot_w2c);
        cb_pose o;
        // compute_std_devs (lib.rs:224-246)
        {
            const double distance = sqrt(vdot(t_sel, t_sel));
            const double n_points = (double)(nt * 4);
            const double rms_error = sqrt(pure_energy / n_points);
            if (rms_error > MAX_TRUSTABLE_RMS) { o.std_devs[0] = o.std_devs[1] = o.std_devs[2] = 1.7976931348623157e308; }
            else {
                const double distance_multiplier = 1.0 + (distance / TAG_SIZE);
                const double base_xy_std = rms_error * distance_multiplier;
                double xy_std = (base_xy_std / sqrt((double)nt)) * XY_STD_DEV_SCALAR;
                xy_std = fmin(fmax(xy_std, 0.01), 10.0);
                const double base_theta_std = rms_error / TAG_SIZE;
                const double val = (base_theta_std * distance_multiplier / sqrt((double)nt)) * THETA_STD_DEV_SCALAR;
                const double theta_std = fmin(fmax(val, 0.05), PI);
                o.std_devs[0] = xy_std; o.std_devs[1] = xy_std; o.std_devs[2] = theta_std;
            }
        }
        // world_to_cam^-1 * robot_to_cam (lib.rs:328-337)
        const Quat qw = quat_from_mat(rot_w2c);
        Quat qi; qi.w = qw.w; qi.x = -qw.x; qi.y = -qw.y; qi.z = -qw.z;
        const V3 ti = vscale(quat_rotate(qi, t_sel), -1.0);
        const V3 robot_pos = vadd(quat_rotate(qi, r2c_t), ti);
        const Quat qr_ = quat_mul(qi, r2c_q);
        double robot_rot[9];
        quat_to_mat(qr_, robot_rot);
        V3 tag_centroid = v3(0, 0, 0);
        for (int i = 0; i < nt; i++) tag_centroid = vadd(tag_centroid, v3(ptags[i].t[0], ptags[i].t[1], ptags[i].t[2]));
        tag_centroid = v3(tag_centroid.x / (double)nt, tag_centroid.y / (double)nt, tag_centroid.z / (double)nt);
        const double vision_yaw = atan2(SQM(robot_rot, 1, 0), SQM(robot_rot, 0, 0));
        double delta_yaw = gyro - vision_yaw;
        {
            const double a = delta_yaw + PI, bb = 2.0 * PI;
            double rr = fmod(a, bb);
            if (rr < 0.0) rr += bb;
            delta_yaw = rr - PI;
        }
        const double delta_deg = fabs(delta_yaw) * (180.0 / PI);
        double weight = fmin(fmax(delta_deg / MAX_GYRO_DELTA, 0.0), 1.0);
        weight = weight * weight * (3.0 - 2.0 * weight);
        const double applied = delta_yaw * weight;
        const double cos_dt = cos(applied), sin_dt = sin(applied);
        double rot_z[9];
        SQM(rot_z, 0, 0) = cos_dt; SQM(rot_z, 0, 1) = -sin_dt; SQM(rot_z, 0, 2) = 0;
        SQM(rot_z, 1, 0) = sin_dt; SQM(rot_z, 1, 1) = cos_dt;  SQM(rot_z, 1, 2) = 0;
        SQM(rot_z, 2, 0) = 0;      SQM(rot_z, 2, 1) = 0;       SQM(rot_z, 2, 2) = 1;
        double rot_z_rot3[9];
        rot3_from_matrix(rot_z, rot_z_rot3);
        const V3 rel = vsub(robot_pos, tag_centroid);
        const V3 piv = vadd(tag_centroid, mat3_mulv(rot_z, rel));
        mat3_mul(rot_z_rot3, robot_rot, o.rot);
        o.pos[0] = piv.x; o.pos[1] = piv.y; o.pos[2] = piv.z;
        out[prob] = o;
        ok[prob] = 1;
    }
}

// OpenCV-5 un-projection (crates/apriltags/src/lib.rs:316-321): pixel -> bearing (x, y, 1) by fixed-point undistortion
__device__ __forceinline__ bool unproject_opencv5(const double *__restrict__ params, double u, double v, double out[3])
{
    const double fx = params[0], fy = params[1], cx = params[2], cy = params[3], k1 = params[4], k2 = params[5], p1 = params[6],
                 p2 = params[7], k3 = params[8];
    const double xd = (u - cx) / fx, yd = (v - cy) / fy;
    double x = xd, y = yd;
    bool good = false;
    for (int it = 0; it < 100; it++) {
        const double r2 = x * x + y * y;
        const double radial = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3));
        const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
        const double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
        if (radial == 0) break;
        const double xn = (xd - dx) / radial, yn = (yd - dy) / radial;
        const double e = (xn - x) * (xn - x) + (yn - y) * (yn - y);
        x = xn; y = yn;
        if (e < 1e-24) { good = true; break; }
    }
    if (!good || !isfinite(x) || !isfinite(y)) { out[0] = out[1] = out[2] = 0; return false; }
    out[0] = x; out[1] = y; out[2] = 1.0;
    return true;
}

// one thread per pixel
__global__ void unproject_opencv5_kernel(const double *__restrict__ params, const double *__restrict__ px, long long n,
                                         double *__restrict__ bearings, uint8_t *__restrict__ ok)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double b[3];
    ok[i] = unproject_opencv5(params, px[2 * i], px[2 * i + 1], b) ? 1 : 0;
    bearings[3 * i] = b[0]; bearings[3 * i + 1] = b[1]; bearings[3 * i + 2] = b[2];
}

// The body of AprilTags::process between detect() and solve_robot_pose() (crates/apriltags/src/lib.rs:303-327), one WARP per
// frame: detections in list order; tags missing from the field layout are skipped (:306-308); a tag is used only if all four
// corners un-project (:324-327).  Writes frame b's SQPnP problem (tags, 4 bearings per tag, tag count); no gyro reading (NaN) or
// no usable tag leaves the count at 0, which the solver answers with "None".  At most max_tags tags are used.
// Eight detections per step: lane 4 d + c un-projects corner c of detection d (the fixed-point iteration of one corner is a chain of
// divisions, ~2 us; one thread per frame spent 37 us on a four-tag frame), a ballot tells which detections kept all four corners and
// their rank among the used tags.
constexpr int ASM_WARPS = 4;
__global__ void __launch_bounds__(ASM_WARPS * 32)
assemble_pose_problems_kernel(const cb_detection *__restrict__ dets, const int32_t *__restrict__ counts, int dets_per_frame,
                              const int32_t *__restrict__ field_ids, const cb_iso3 *__restrict__ field_poses, int n_field,
                              const double *__restrict__ cam9, const double *__restrict__ gyro, int max_tags,
                              cb_iso3 *__restrict__ tags, double *__restrict__ bearings, int32_t *__restrict__ n_tags,
                              int frame_base, int n_frames)
{
    const int f = blockIdx.x * ASM_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= n_frames) return;
    const uint32_t full = 0xffffffffu;
    const int b = frame_base + f;
    int used = 0;
    const double gy = gyro[b];
    if (gy == gy) {
        const cb_detection *d = dets + (size_t)f * dets_per_frame;
        const int n = counts[f];
        const int c = lane & 3, first = lane & ~3;
        for (int k0 = 0; k0 < n && used < max_tags; k0 += 8) {
            const int k = k0 + (lane >> 2);
            int t = -1;
            double br[3] = {0, 0, 0};
            bool corner_ok = false;
            if (k < n) {
                const int id = d[k].id;
                for (int q = 0; q < n_field; q++)
                    if (field_ids[q] == id) { t = q; break; }
                if (t >= 0) corner_ok = unproject_opencv5(cam9, d[k].p[c][0], d[k].p[c][1], br);
            }
            const uint32_t okb = __ballot_sync(full, corner_ok);
            const bool tag_ok = ((okb >> first) & 0xfu) == 0xfu;
            const uint32_t lead = __ballot_sync(full, tag_ok && c == 0);
            const int slot = used + __popc(lead & ((1u << first) - 1u));
            if (tag_ok && slot < max_tags) {
                if (c == 0) tags[(size_t)b * max_tags + slot] = field_poses[t];
                double *o = bearings + ((size_t)b * max_tags + slot) * 12 + 3 * c;
                o[0] = br[0]; o[1] = br[1]; o[2] = br[2];
            }
            used = min(max_tags, used + __popc(lead));
        }
    }
    if (lane == 0) n_tags[b] = used;
}

}  // namespace cb
