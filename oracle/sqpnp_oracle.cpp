/*
 * oracle/sqpnp_oracle.cpp -- CPU restatement of chalkydri_sqpnp (TEST INFRASTRUCTURE ONLY).
 *
 * Follows /root/reference/crates/chalkydri_sqpnp/src/lib.rs line by line (citations at each function).
 * The reference delegates its dense kernels to nalgebra 0.34.1 (crates/chalkydri_sqpnp/Cargo.toml:7),
 * which is not vendored; they are restated here with their published algorithms:
 *     symmetric_eigen (lib.rs:398)   -> cyclic Jacobi on the max-abs-scaled matrix
 *     Matrix3::svd    (lib.rs:45)    -> one-sided (Hestenes) Jacobi, singular values sorted descending
 *     lu().solve      (lib.rs:111)   -> partial-pivot LU, multipliers formed with the reciprocal pivot
 *     try_inverse 3x3 (lib.rs:171)   -> cofactor formula
 *     Rot3::from_matrix (lib.rs:289) -> Mueller et al. iterative rotation extraction from the identity
 *
 * PARITY UNPINNED: the reference has no test for the solver.  One behaviour is intrinsically
 * implementation dependent: for a single (planar) tag Omega has a >=4 dimensional null space, so
 * "the three smallest eigenvectors" (lib.rs:400-403) are an arbitrary basis picked by rounding noise
 * inside nalgebra's QR iteration.  The Newton refinement makes the final pose insensitive to that
 * choice in all but rare cases; tests/test_sqpnp_oracle.py measures that sensitivity.
 *
 * Matrices are column-major (nalgebra storage), r_vec = vec(R) by columns, exactly like
 * Mat3::from_column_slice (lib.rs:43,271).
 */
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

const double XY_STD_DEV_SCALAR = 5.0;     /* lib.rs:29 */
const double THETA_STD_DEV_SCALAR = 2.0;  /* lib.rs:30 */
const double MAX_TRUSTABLE_RMS = 0.1;     /* lib.rs:31 */
const double MAX_GYRO_DELTA = 30.0;       /* lib.rs:35 */
const double TAG_SIZE = 0.1651;           /* lib.rs:38 */
const double CORNER_DISTANCE = TAG_SIZE / 2.0;
const double PI = 3.14159265358979323846;

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

/* column-major 3x3 helpers: m[c*3 + r] */
inline double &M3(double *m, int r, int c) { return m[c * 3 + r]; }
inline double M3c(const double *m, int r, int c) { return m[c * 3 + r]; }
void mat3_mul(const double *a, const double *b, double *out)
{
    double t[9];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += M3c(a, r, k) * M3c(b, k, c);
            t[c * 3 + r] = acc;
        }
    memcpy(out, t, sizeof(t));
}
V3 mat3_mulv(const double *a, V3 v)
{
    return {M3c(a, 0, 0) * v.x + M3c(a, 0, 1) * v.y + M3c(a, 0, 2) * v.z,
            M3c(a, 1, 0) * v.x + M3c(a, 1, 1) * v.y + M3c(a, 1, 2) * v.z,
            M3c(a, 2, 0) * v.x + M3c(a, 2, 1) * v.y + M3c(a, 2, 2) * v.z};
}
double mat3_det(const double *m)
{
    double m11 = M3c(m, 0, 0), m12 = M3c(m, 0, 1), m13 = M3c(m, 0, 2);
    double m21 = M3c(m, 1, 0), m22 = M3c(m, 1, 1), m23 = M3c(m, 1, 2);
    double m31 = M3c(m, 2, 0), m32 = M3c(m, 2, 1), m33 = M3c(m, 2, 2);
    double minor_m12_m23 = m22 * m33 - m32 * m23;
    double minor_m11_m23 = m21 * m33 - m31 * m23;
    double minor_m11_m22 = m21 * m32 - m31 * m22;
    return m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
}
/* nalgebra try_inverse for 3x3 (cofactors / determinant); false when det == 0 */
bool mat3_try_inverse(const double *m, double *out)
{
    double m11 = M3c(m, 0, 0), m12 = M3c(m, 0, 1), m13 = M3c(m, 0, 2);
    double m21 = M3c(m, 1, 0), m22 = M3c(m, 1, 1), m23 = M3c(m, 1, 2);
    double m31 = M3c(m, 2, 0), m32 = M3c(m, 2, 1), m33 = M3c(m, 2, 2);
    double minor_m12_m23 = m22 * m33 - m32 * m23;
    double minor_m11_m23 = m21 * m33 - m31 * m23;
    double minor_m11_m22 = m21 * m32 - m31 * m22;
    double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
    if (det == 0) return false;
    M3(out, 0, 0) = minor_m12_m23 / det;
    M3(out, 0, 1) = (m13 * m32 - m33 * m12) / det;
    M3(out, 0, 2) = (m12 * m23 - m22 * m13) / det;
    M3(out, 1, 0) = -minor_m11_m23 / det;
    M3(out, 1, 1) = (m11 * m33 - m31 * m13) / det;
    M3(out, 1, 2) = (m13 * m21 - m23 * m11) / det;
    M3(out, 2, 0) = minor_m11_m22 / det;
    M3(out, 2, 1) = (m12 * m31 - m32 * m11) / det;
    M3(out, 2, 2) = (m11 * m22 - m21 * m12) / det;
    return true;
}

/* ---- quaternion / isometry (nalgebra conventions; q = w,x,y,z) ---- */
struct Quat { double w, x, y, z; };
inline V3 quat_rotate(const Quat &q, V3 v)
{
    V3 qv{q.x, q.y, q.z};
    V3 t = cross(qv, v) * 2.0;
    V3 c = cross(qv, t);
    return t * q.w + c + v;
}
inline Quat quat_mul(const Quat &a, const Quat &b)
{
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
            a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
            a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
void quat_to_mat(const Quat &q, double *m)
{
    double i = q.x, j = q.y, k = q.z, w = q.w;
    double ww = w * w, ii = i * i, jj = j * j, kk = k * k;
    double ij = i * j * 2, wk = w * k * 2, wj = w * j * 2, ik = i * k * 2, jk = j * k * 2, wi = w * i * 2;
    M3(m, 0, 0) = ww + ii - jj - kk; M3(m, 0, 1) = ij - wk; M3(m, 0, 2) = wj + ik;
    M3(m, 1, 0) = wk + ij; M3(m, 1, 1) = ww - ii + jj - kk; M3(m, 1, 2) = jk - wi;
    M3(m, 2, 0) = ik - wj; M3(m, 2, 1) = wi + jk; M3(m, 2, 2) = ww - ii - jj + kk;
}
Quat quat_from_mat(const double *m)
{
    double tr = M3c(m, 0, 0) + M3c(m, 1, 1) + M3c(m, 2, 2);
    Quat q;
    if (tr > 0) {
        double denom = sqrt(tr + 1.0) * 2.0;
        q = {0.25 * denom, (M3c(m, 2, 1) - M3c(m, 1, 2)) / denom, (M3c(m, 0, 2) - M3c(m, 2, 0)) / denom,
             (M3c(m, 1, 0) - M3c(m, 0, 1)) / denom};
    } else if (M3c(m, 0, 0) > M3c(m, 1, 1) && M3c(m, 0, 0) > M3c(m, 2, 2)) {
        double denom = sqrt(1.0 + M3c(m, 0, 0) - M3c(m, 1, 1) - M3c(m, 2, 2)) * 2.0;
        q = {(M3c(m, 2, 1) - M3c(m, 1, 2)) / denom, 0.25 * denom, (M3c(m, 0, 1) + M3c(m, 1, 0)) / denom,
             (M3c(m, 0, 2) + M3c(m, 2, 0)) / denom};
    } else if (M3c(m, 1, 1) > M3c(m, 2, 2)) {
        double denom = sqrt(1.0 + M3c(m, 1, 1) - M3c(m, 0, 0) - M3c(m, 2, 2)) * 2.0;
        q = {(M3c(m, 0, 2) - M3c(m, 2, 0)) / denom, (M3c(m, 0, 1) + M3c(m, 1, 0)) / denom, 0.25 * denom,
             (M3c(m, 1, 2) + M3c(m, 2, 1)) / denom};
    } else {
        double denom = sqrt(1.0 + M3c(m, 2, 2) - M3c(m, 0, 0) - M3c(m, 1, 1)) * 2.0;
        q = {(M3c(m, 1, 0) - M3c(m, 0, 1)) / denom, (M3c(m, 0, 2) + M3c(m, 2, 0)) / denom,
             (M3c(m, 1, 2) + M3c(m, 2, 1)) / denom, 0.25 * denom};
    }
    return q;
}
struct Iso { V3 t; Quat q; };
inline Iso iso_from(const orc_iso3 &o) { return {{o.t[0], o.t[1], o.t[2]}, {o.q[0], o.q[1], o.q[2], o.q[3]}}; }
inline V3 iso_apply(const Iso &i, V3 p) { return quat_rotate(i.q, p) + i.t; }
inline Iso iso_inverse(const Iso &i)
{
    Quat qi{i.q.w, -i.q.x, -i.q.y, -i.q.z};
    V3 t = quat_rotate(qi, i.t) * -1.0;
    return {t, qi};
}
inline Iso iso_mul(const Iso &a, const Iso &b) { return {quat_rotate(a.q, b.t) + a.t, quat_mul(a.q, b.q)}; }

/* ---- symmetric eigen 9x9: cyclic Jacobi (restates nalgebra symmetric_eigen's contract: A = V diag(d) V^T) ---- */
void sym_eigen9(const double *a_in, double *d, double *v /* col-major */)
{
    const int n = 9;
    double a[81];
    double amax = 0;
    for (int i = 0; i < 81; i++) amax = std::max(amax, fabs(a_in[i]));
    for (int i = 0; i < 81; i++) v[i] = 0;
    for (int i = 0; i < n; i++) v[i * n + i] = 1;
    if (amax == 0) { for (int i = 0; i < n; i++) d[i] = 0; return; }
    for (int i = 0; i < 81; i++) a[i] = a_in[i] / amax;   /* nalgebra unscales by camax first */
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0;
        for (int q = 1; q < n; q++)
            for (int p = 0; p < q; p++) off += a[q * n + p] * a[q * n + p];
        if (off <= 1e-34) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = a[q * n + p];
                if (apq == 0) continue;
                double app = a[p * n + p], aqq = a[q * n + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                if (theta < 0) t = -t;
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                /* A <- J^T A J with J = [[c, s], [-s, c]] on (p,q) */
                for (int k = 0; k < n; k++) {   /* columns p,q */
                    double akp = a[p * n + k], akq = a[q * n + k];
                    a[p * n + k] = c * akp - s * akq;
                    a[q * n + k] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {   /* rows p,q */
                    double apk = a[k * n + p], aqk = a[k * n + q];
                    a[k * n + p] = c * apk - s * aqk;
                    a[k * n + q] = s * apk + c * aqk;
                }
                a[q * n + p] = 0; a[p * n + q] = 0;
                for (int k = 0; k < n; k++) {   /* V <- V J */
                    double vkp = v[p * n + k], vkq = v[q * n + k];
                    v[p * n + k] = c * vkp - s * vkq;
                    v[q * n + k] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) d[i] = a[i * n + i] * amax;
}

/* f64::total_cmp key */
inline int64_t total_key(double x)
{
    int64_t b;
    memcpy(&b, &x, 8);
    b ^= (int64_t)(((uint64_t)(b >> 63)) >> 1);
    return b;
}

/* ---- 3x3 SVD by one-sided Jacobi; returns U, V (col-major) with singular values descending ---- */
void svd3(const double *m, double *U, double *S, double *V)
{
    double a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(a, m, sizeof(a));
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
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                if (zeta < 0) t = -t;
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
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
    for (int i = 1; i < 3; i++)   /* stable insertion sort, descending */
        for (int j = i; j > 0 && sv[ord[j - 1]] < sv[ord[j]]; j--) std::swap(ord[j - 1], ord[j]);
    int rank = 0;
    for (int j = 0; j < 3; j++) {
        int o = ord[j];
        S[j] = sv[o];
        for (int k = 0; k < 3; k++) V[j * 3 + k] = v[o * 3 + k];
        if (sv[o] > 1e-300 && sv[o] > 1e-15 * sv[ord[0]]) {
            for (int k = 0; k < 3; k++) U[j * 3 + k] = a[o * 3 + k] / sv[o];
            rank = j + 1;
        } else {
            for (int k = 0; k < 3; k++) U[j * 3 + k] = 0;
        }
    }
    /* complete U to an orthonormal basis when m is rank deficient (nalgebra returns some orthogonal U;
       which one is implementation defined) */
    if (rank == 0) { for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? 1 : 0; rank = 3; }
    if (rank == 1) {
        V3 u0{U[0], U[1], U[2]};
        int j = 0;
        double best = fabs(U[0]);
        for (int k = 1; k < 3; k++) if (fabs(U[k]) < best) { best = fabs(U[k]); j = k; }
        V3 e{j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0};
        V3 u1 = e - u0 * dot(e, u0);
        double nrm = sqrt(dot(u1, u1));
        u1 = u1 * (1.0 / nrm);
        U[3] = u1.x; U[4] = u1.y; U[5] = u1.z;
        rank = 2;
    }
    if (rank == 2) {
        V3 u2 = cross(V3{U[0], U[1], U[2]}, V3{U[3], U[4], U[5]});
        U[6] = u2.x; U[7] = u2.y; U[8] = u2.z;
    }
}

/* lib.rs:42-59 */
bool nearest_so3(const double *r_vec, double *out)
{
    double U[9], S[3], V[9], Vt[9], rot[9];
    svd3(r_vec, U, S, V);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) M3(Vt, r, c) = M3c(V, c, r);
    mat3_mul(U, Vt, rot);
    if (mat3_det(rot) < 0.0) {
        U[6] = -U[6]; U[7] = -U[7]; U[8] = -U[8];
        mat3_mul(U, Vt, rot);
    }
    memcpy(out, rot, sizeof(rot));
    return true;
}

/* lib.rs:62-95: h (6) and jac (6x9, stored row-major here: jac[row*9+col]) */
void constraints_and_jacobian(const double *r, double *h, double *jac)
{
    const double *c1 = r, *c2 = r + 3, *c3 = r + 6;
    auto d3 = [](const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    h[0] = d3(c1, c1) - 1.0; h[1] = d3(c2, c2) - 1.0; h[2] = d3(c3, c3) - 1.0;
    h[3] = d3(c1, c2); h[4] = d3(c1, c3); h[5] = d3(c2, c3);
    memset(jac, 0, sizeof(double) * 54);
    for (int k = 0; k < 3; k++) {
        jac[0 * 9 + 0 + k] = 2.0 * c1[k];
        jac[1 * 9 + 3 + k] = 2.0 * c2[k];
        jac[2 * 9 + 6 + k] = 2.0 * c3[k];
        jac[3 * 9 + 0 + k] = c2[k]; jac[3 * 9 + 3 + k] = c1[k];
        jac[4 * 9 + 0 + k] = c3[k]; jac[4 * 9 + 6 + k] = c1[k];
        jac[5 * 9 + 3 + k] = c3[k]; jac[5 * 9 + 6 + k] = c2[k];
    }
}

/* nalgebra LU (partial pivoting, reciprocal-pivot multipliers) + solve; false when singular.  a is row-major 15x15 */
bool lu_solve15(double *a, double *b)
{
    const int n = 15;
    for (int i = 0; i < n; i++) {
        int piv = i;
        double best = fabs(a[i * n + i]);
        for (int r = i + 1; r < n; r++)
            if (fabs(a[r * n + i]) > best) { best = fabs(a[r * n + i]); piv = r; }
        double diag = a[piv * n + i];
        if (diag == 0) continue;   /* no non-zero entries on this column */
        if (piv != i) {
            for (int c = 0; c < n; c++) std::swap(a[i * n + c], a[piv * n + c]);
            std::swap(b[i], b[piv]);
        }
        double inv_diag = 1.0 / diag;
        for (int r = i + 1; r < n; r++) {
            double coeff = a[r * n + i] * inv_diag;
            a[r * n + i] = coeff;
            for (int c = i + 1; c < n; c++) a[r * n + c] -= coeff * a[i * n + c];
        }
    }
    /* L y = b (unit diagonal) */
    for (int i = 0; i < n; i++)
        for (int r = i + 1; r < n; r++) b[r] -= a[r * n + i] * b[i];
    /* U x = y */
    for (int i = n - 1; i >= 0; i--) {
        double diag = a[i * n + i];
        if (diag == 0) return false;
        b[i] = b[i] / diag;
        for (int r = 0; r < i; r++) b[r] -= a[r * n + i] * b[i];
    }
    return true;
}

/* lib.rs:98-115 */
bool solve_newton(const double *r, const double *omega /* col-major 9x9, symmetric */, const double *h, const double *jac, double *delta)
{
    double lhs[225], rhs[15];
    memset(lhs, 0, sizeof(lhs));
    for (int i = 0; i < 9; i++)
        for (int j = 0; j < 9; j++) lhs[i * 15 + j] = omega[j * 9 + i];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 9; j++) {
            lhs[j * 15 + 9 + i] = jac[i * 9 + j];
            lhs[(9 + i) * 15 + j] = jac[i * 9 + j];
        }
    for (int i = 0; i < 9; i++) {
        double acc = 0;
        for (int j = 0; j < 9; j++) acc += omega[j * 9 + i] * r[j];
        rhs[i] = -acc;
    }
    for (int i = 0; i < 6; i++) rhs[9 + i] = -h[i];
    if (!lu_solve15(lhs, rhs)) return false;
    for (int i = 0; i < 9; i++) delta[i] = rhs[i];
    return true;
}

double quad_form(const double *omega, const double *r)
{
    double e = 0;
    double tmp[9];
    for (int i = 0; i < 9; i++) {
        double acc = 0;
        for (int j = 0; j < 9; j++) acc += omega[j * 9 + i] * r[j];
        tmp[i] = acc;
    }
    for (int i = 0; i < 9; i++) e += r[i] * tmp[i];
    return e;
}

/* lib.rs:463-479 */
double optimization(double *r, const double *omega, int max_iter, double tol_sq)
{
    for (int it = 0; it < max_iter; it++) {
        double h[6], jac[54], delta[9];
        constraints_and_jacobian(r, h, jac);
        if (!solve_newton(r, omega, h, jac, delta)) break;
        double nsq = 0;
        for (int i = 0; i < 9; i++) { r[i] += delta[i]; nsq += delta[i] * delta[i]; }
        if (nsq < tol_sq) break;
    }
    return quad_form(omega, r);
}

struct LinearSys { double omega[81], q_tt_inv[9], q_rt[27]; /* q_rt col-major 9x3: q_rt[c*9 + r] */ };

/* lib.rs:124-180 */
void build_linear_system(const V3 *p3, const V3 *p2, int n, LinearSys &sys)
{
    double q_rr[81], q_rt[27], q_tt[9];
    memset(q_rr, 0, sizeof(q_rr)); memset(q_rt, 0, sizeof(q_rt)); memset(q_tt, 0, sizeof(q_tt));
    for (int i = 0; i < n; i++) {
        const V3 &v = p2[i];
        double sq_norm = v.x * v.x + v.y * v.y + v.z * v.z;
        double inv_norm = 1.0 / sq_norm;
        double vv[3] = {v.x, v.y, v.z};
        double P[9];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) M3(P, r, c) = (r == c ? 1.0 : 0.0) - (vv[r] * vv[c]) * inv_norm;
        for (int k = 0; k < 9; k++) q_tt[k] += P[k];
        double X[3] = {p3[i].x, p3[i].y, p3[i].z};
        double Pk[3][9];
        for (int a = 0; a < 3; a++) for (int k = 0; k < 9; k++) Pk[a][k] = P[k] * X[a];
        for (int a = 0; a < 3; a++)       /* q_rt block rows 3a..3a+2, cols 0..2 */
            for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) q_rt[c * 9 + 3 * a + r] += M3c(Pk[a], r, c);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                /* reference accumulates px*x, py*y, pz*z on the diagonal and pxy, pxz, pyz (px*y, px*z, py*z) off it */
                int lo = a < b ? a : b, hi = a < b ? b : a;
                for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++)
                    q_rr[(3 * b + c) * 9 + 3 * a + r] += M3c(Pk[lo], r, c) * X[hi];
            }
    }
    if (!mat3_try_inverse(q_tt, sys.q_tt_inv)) memset(sys.q_tt_inv, 0, sizeof(sys.q_tt_inv));
    /* temp = q_rt * q_tt_inv (9x3); omega = q_rr - temp * q_rt^T */
    double temp[27];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 9; r++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += q_rt[k * 9 + r] * M3c(sys.q_tt_inv, k, c);
            temp[c * 9 + r] = acc;
        }
    for (int c = 0; c < 9; c++)
        for (int r = 0; r < 9; r++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += temp[k * 9 + r] * q_rt[k * 9 + c];
            sys.omega[c * 9 + r] = q_rr[c * 9 + r] - acc;
        }
    memcpy(sys.q_rt, q_rt, sizeof(q_rt));
}

/* nalgebra Rotation3::from_matrix_eps(m, EPSILON, unlimited, identity) (lib.rs:289,370) */
void rot3_from_matrix(const double *m, double *out)
{
    const double eps = 2.220446049250313e-16;
    double rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int perturb_axis = 0;
    for (int it = 0; it < 10000; it++) {
        V3 axis{0, 0, 0};
        double denom = 0;
        for (int c = 0; c < 3; c++) {
            V3 rc{rot[c * 3], rot[c * 3 + 1], rot[c * 3 + 2]}, mc{m[c * 3], m[c * 3 + 1], m[c * 3 + 2]};
            axis = axis + cross(rc, mc);
            denom += dot(rc, mc);
        }
        V3 axisangle = axis * (1.0 / (fabs(denom) + eps));
        double angle = sqrt(dot(axisangle, axisangle));
        auto axis_angle_mat = [](V3 u, double ang, double *R) {
            double s = sin(ang), c = cos(ang), one_c = 1.0 - c;
            double ux = u.x, uy = u.y, uz = u.z;
            M3(R, 0, 0) = ux * ux * one_c + c;       M3(R, 0, 1) = ux * uy * one_c - uz * s; M3(R, 0, 2) = ux * uz * one_c + uy * s;
            M3(R, 1, 0) = ux * uy * one_c + uz * s;  M3(R, 1, 1) = uy * uy * one_c + c;      M3(R, 1, 2) = uy * uz * one_c - ux * s;
            M3(R, 2, 0) = ux * uz * one_c - uy * s;  M3(R, 2, 1) = uy * uz * one_c + ux * s; M3(R, 2, 2) = uz * uz * one_c + c;
        };
        if (angle > eps) {
            double R[9];
            axis_angle_mat(axisangle * (1.0 / angle), angle, R);
            mat3_mul(R, rot, rot);
        } else {
            /* stuck: minimum or maximum of |m - rot|? perturb and compare */
            double nsq = 0;
            for (int k = 0; k < 9; k++) nsq += (m[k] - rot[k]) * (m[k] - rot[k]);
            double pert[9];
            memcpy(pert, rot, sizeof(pert));
            V3 ax{perturb_axis == 0 ? 1.0 : 0.0, perturb_axis == 1 ? 1.0 : 0.0, perturb_axis == 2 ? 1.0 : 0.0};
            double R[9], nsq2 = nsq;
            axis_angle_mat(ax, sqrt(eps), R);
            for (int tries = 0; tries < 64; tries++) {
                mat3_mul(pert, R, pert);
                nsq2 = 0;
                for (int k = 0; k < 9; k++) nsq2 += (m[k] - pert[k]) * (m[k] - pert[k]);
                if (fabs(nsq - nsq2) > eps) break;
            }
            if (nsq <= nsq2) break;   /* a minimum: done */
            perturb_axis = (perturb_axis + 1) % 3;
            memcpy(rot, pert, sizeof(rot));
        }
    }
    memcpy(out, rot, sizeof(rot));
}

struct Solver {
    int max_iter = 15;
    double tol_sq = 1e-16;
    double gyro_cos = 0, gyro_sin = 0, sign_change_error = 0;
    V3 fwd_in_cam{0, 0, 1};
    std::vector<V3> buffer;
    struct Cand { double r[9]; double energy; };
    std::vector<Cand> candidates;

    /* lib.rs:379-394 */
    void corner_points_from_center(const Iso *iso, int n)
    {
        const double S = CORNER_DISTANCE;
        const V3 corners[4] = {{0.0, -S, -S}, {0.0, S, -S}, {0.0, S, S}, {0.0, -S, S}};
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 4; c++) buffer.push_back(iso_apply(iso[i], corners[c]));
    }

    /* lib.rs:396-428 */
    void solve_rotation_candidates(const double *omega)
    {
        candidates.clear();
        double evals[9], evecs[81];
        sym_eigen9(omega, evals, evecs);
        int idx[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
        std::stable_sort(idx, idx + 9, [&](int a, int b) { return total_key(evals[a]) < total_key(evals[b]); });
        for (int t = 0; t < 3; t++) {
            const double *e = &evecs[idx[t] * 9];
            for (int si = 0; si < 2; si++) {
                double sign = si == 0 ? -1.0 : 1.0;
                double guess[9], r[9];
                for (int k = 0; k < 9; k++) guess[k] = e[k] * sign;
                if (!nearest_so3(guess, r)) continue;
                double energy = optimization(r, omega, max_iter, tol_sq);
                const V3 &d = fwd_in_cam;
                double robot_fwd_x = r[0] * d.x + r[1] * d.y + r[2] * d.z;
                double robot_fwd_y = r[3] * d.x + r[4] * d.y + r[5] * d.z;
                double dt = (robot_fwd_x * gyro_cos) + (robot_fwd_y * gyro_sin);
                double angle_error = std::max(1.0 - dt, 0.0);
                energy += sign_change_error * angle_error;
                Cand c;
                memcpy(c.r, r, sizeof(c.r));
                c.energy = energy;
                candidates.push_back(c);
            }
        }
        std::stable_sort(candidates.begin(), candidates.end(),
                         [](const Cand &a, const Cand &b) { return total_key(a.energy) < total_key(b.energy); });
    }

    /* lib.rs:248-295 */
    bool solve(const Iso *tags, int n_tags, const V3 *p2, int n2, double *rot_out, V3 &t_out, double &pure_energy)
    {
        corner_points_from_center(tags, n_tags);
        if (buffer.size() < 3 || (int)buffer.size() != n2) return false;
        V3 centroid{0, 0, 0};
        for (auto &p : buffer) centroid = centroid + p;
        centroid = {centroid.x / (double)buffer.size(), centroid.y / (double)buffer.size(), centroid.z / (double)buffer.size()};
        std::vector<V3> local(buffer.size());
        for (size_t i = 0; i < buffer.size(); i++) local[i] = buffer[i] - centroid;
        LinearSys sys;
        build_linear_system(local.data(), p2, (int)local.size(), sys);
        solve_rotation_candidates(sys.omega);
        bool have = false;
        double best_score = 1.7976931348623157e308;
        for (auto &cand : candidates) {
            const double *r = cand.r;
            /* t_local = -(q_tt_inv * (q_rt^T r)) */
            double qr[3];
            for (int c = 0; c < 3; c++) {
                double acc = 0;
                for (int k = 0; k < 9; k++) acc += sys.q_rt[c * 9 + k] * r[k];
                qr[c] = acc;
            }
            V3 tl = mat3_mulv(sys.q_tt_inv, V3{qr[0], qr[1], qr[2]});
            tl = tl * -1.0;
            V3 t = tl - mat3_mulv(r, centroid);
            bool all_in_front = true;
            for (auto &p : buffer) {
                V3 pc = mat3_mulv(r, p) + t;
                if (!(pc.z > 0.0)) { all_in_front = false; break; }
            }
            if (!all_in_front) continue;
            if (cand.energy < best_score) {
                best_score = cand.energy;
                pure_energy = quad_form(sys.omega, r);
                rot3_from_matrix(r, rot_out);
                t_out = t;
                have = true;
            }
        }
        return have;
    }

    /* lib.rs:224-246 */
    void compute_std_devs(double pure_geometric_energy, double distance, int n_tags, double *out) const
    {
        double n_points = (double)(n_tags * 4);
        double rms_error = sqrt(pure_geometric_energy / n_points);
        if (rms_error > MAX_TRUSTABLE_RMS) { out[0] = out[1] = out[2] = 1.7976931348623157e308; return; }
        double distance_multiplier = 1.0 + (distance / TAG_SIZE);
        double base_xy_std = rms_error * distance_multiplier;
        double xy_std = (base_xy_std / sqrt((double)n_tags)) * XY_STD_DEV_SCALAR;
        xy_std = std::min(std::max(xy_std, 0.01), 10.0);
        double base_theta_std = rms_error / TAG_SIZE;
        double val = (base_theta_std * distance_multiplier / sqrt((double)n_tags)) * THETA_STD_DEV_SCALAR;
        double theta_std = std::min(std::max(val, 0.05), PI);
        out[0] = xy_std; out[1] = xy_std; out[2] = theta_std;
    }

    /* lib.rs:297-377 */
    bool solve_robot_pose(const Iso *tags, int n_tags, const V3 *p2, int n2, const Iso &robot_to_cam, double gyro,
                          double sce, orc_robot_pose *out)
    {
        gyro_cos = cos(gyro); gyro_sin = sin(gyro);
        sign_change_error = sce;
        buffer.clear(); candidates.clear();
        double r2c[9];
        quat_to_mat(robot_to_cam.q, r2c);
        fwd_in_cam = V3{r2c[0], r2c[1], r2c[2]};   /* column 0 */
        double rot_w2c[9]; V3 t_w2c; double pure_energy = 0;
        if (!solve(tags, n_tags, p2, n2, rot_w2c, t_w2c, pure_energy)) return false;
        double distance = sqrt(dot(t_w2c, t_w2c));
        compute_std_devs(pure_energy, distance, n_tags, out->std_devs);
        Iso world_to_cam{t_w2c, quat_from_mat(rot_w2c)};
        Iso t_world_robot = iso_mul(iso_inverse(world_to_cam), robot_to_cam);
        V3 robot_pos = t_world_robot.t;
        double robot_rot[9];
        quat_to_mat(t_world_robot.q, robot_rot);
        V3 tag_centroid{0, 0, 0};
        for (int i = 0; i < n_tags; i++) tag_centroid = tag_centroid + tags[i].t;
        tag_centroid = {tag_centroid.x / (double)n_tags, tag_centroid.y / (double)n_tags, tag_centroid.z / (double)n_tags};
        double vision_fwd_x = M3c(robot_rot, 0, 0), vision_fwd_y = M3c(robot_rot, 1, 0);
        double vision_yaw = atan2(vision_fwd_y, vision_fwd_x);
        double delta_yaw = gyro - vision_yaw;
        {   /* rem_euclid */
            double a = delta_yaw + PI, b = 2.0 * PI;
            double r = fmod(a, b);
            if (r < 0.0) r += b;
            delta_yaw = r - PI;
        }
        double delta_deg = fabs(delta_yaw) * (180.0 / PI);
        double weight = std::min(std::max(delta_deg / MAX_GYRO_DELTA, 0.0), 1.0);
        weight = weight * weight * (3.0 - 2.0 * weight);
        double applied = delta_yaw * weight;
        double cos_dt = cos(applied), sin_dt = sin(applied);
        double rot_z[9];
        M3(rot_z, 0, 0) = cos_dt; M3(rot_z, 0, 1) = -sin_dt; M3(rot_z, 0, 2) = 0;
        M3(rot_z, 1, 0) = sin_dt; M3(rot_z, 1, 1) = cos_dt;  M3(rot_z, 1, 2) = 0;
        M3(rot_z, 2, 0) = 0;      M3(rot_z, 2, 1) = 0;       M3(rot_z, 2, 2) = 1;
        double rot_z_rot3[9];
        rot3_from_matrix(rot_z, rot_z_rot3);
        V3 rel = robot_pos - tag_centroid;
        V3 piv = tag_centroid + mat3_mulv(rot_z, rel);
        mat3_mul(rot_z_rot3, robot_rot, out->rot);
        out->pos[0] = piv.x; out->pos[1] = piv.y; out->pos[2] = piv.z;
        return true;
    }
};

}  // namespace

extern "C" {

int orc_sqpnp_solve_robot_pose(const orc_iso3 *tags, int n_tags, const double *bearings, int n_bearings,
                               const orc_iso3 *robot_to_cam, double gyro, double sign_change_error, int max_iter,
                               double tol_sq, orc_robot_pose *out)
{
    Solver s;
    s.max_iter = max_iter; s.tol_sq = tol_sq;
    std::vector<Iso> isos(n_tags);
    for (int i = 0; i < n_tags; i++) isos[i] = iso_from(tags[i]);
    std::vector<V3> p2(n_bearings);
    for (int i = 0; i < n_bearings; i++) p2[i] = {bearings[i * 3], bearings[i * 3 + 1], bearings[i * 3 + 2]};
    return s.solve_robot_pose(isos.data(), n_tags, p2.data(), n_bearings, iso_from(*robot_to_cam), gyro, sign_change_error, out) ? 1 : 0;
}

int orc_sqpnp_batch(const orc_iso3 *tags, const double *bearings, const int32_t *n_tags, int max_tags,
                    const orc_iso3 *robot_to_cam, const double *gyro, double sign_change_error, int64_t n,
                    orc_robot_pose *out, uint8_t *ok, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        const int64_t chunk = 256;
        for (;;) {
            int64_t b0 = next.fetch_add(chunk);
            if (b0 >= n) break;
            int64_t b1 = std::min(n, b0 + chunk);
            for (int64_t i = b0; i < b1; i++) {
                int nt = n_tags[i];
                ok[i] = (uint8_t)orc_sqpnp_solve_robot_pose(tags + i * max_tags, nt, bearings + i * max_tags * 12, nt * 4,
                                                            robot_to_cam, gyro[i], sign_change_error, 15, 1e-16, &out[i]);
            }
        }
    };
    if (nthreads == 1) work();
    else {
        std::vector<std::thread> ts;
        for (int i = 0; i < nthreads; i++) ts.emplace_back(work);
        for (auto &t : ts) t.join();
    }
    return 0;
}

void orc_create_solver_camera_transform(double fwd, double left, double up, double roll_deg, double pitch_deg,
                                        double yaw_deg, orc_iso3 *out)
{
    /* lib.rs:430-461 */
    double roll = roll_deg * (PI / 180.0), pitch = pitch_deg * (PI / 180.0), yaw = yaw_deg * (PI / 180.0);
    double sr = sin(roll * 0.5), cr = cos(roll * 0.5), sp = sin(pitch * 0.5), cp = cos(pitch * 0.5);
    double sy = sin(yaw * 0.5), cy = cos(yaw * 0.5);
    Quat q{cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy};
    Iso robot_pose_of_cam{{fwd, left, up}, q};
    double m[9];
    M3(m, 0, 0) = 0;  M3(m, 0, 1) = 0;  M3(m, 0, 2) = 1;
    M3(m, 1, 0) = -1; M3(m, 1, 1) = 0;  M3(m, 1, 2) = 0;
    M3(m, 2, 0) = 0;  M3(m, 2, 1) = -1; M3(m, 2, 2) = 0;
    Iso nwu_to_cv{{0, 0, 0}, quat_from_mat(m)};
    Iso r = iso_inverse(iso_mul(robot_pose_of_cam, nwu_to_cv));
    out->t[0] = r.t.x; out->t[1] = r.t.y; out->t[2] = r.t.z;
    out->q[0] = r.q.w; out->q[1] = r.q.x; out->q[2] = r.q.y; out->q[3] = r.q.z;
}

void orc_sqpnp_omega(const double *pts3d, const double *bearings, int n, double *omega, double *q_tt_inv, double *q_rt)
{
    std::vector<V3> p3(n), p2(n);
    for (int i = 0; i < n; i++) {
        p3[i] = {pts3d[i * 3], pts3d[i * 3 + 1], pts3d[i * 3 + 2]};
        p2[i] = {bearings[i * 3], bearings[i * 3 + 1], bearings[i * 3 + 2]};
    }
    LinearSys sys;
    build_linear_system(p3.data(), p2.data(), n, sys);
    memcpy(omega, sys.omega, sizeof(sys.omega));
    memcpy(q_tt_inv, sys.q_tt_inv, sizeof(sys.q_tt_inv));
    memcpy(q_rt, sys.q_rt, sizeof(sys.q_rt));
}

void orc_sym_eigen9(const double *a, double *evals, double *evecs) { sym_eigen9(a, evals, evecs); }
void orc_nearest_so3(const double *r9, double *out9) { nearest_so3(r9, out9); }

int orc_unproject_opencv5(const double *k, double u, double v, double *out)
{
    /* OpenCVModel5 (fx,fy,cx,cy,k1,k2,p1,p2,k3): pixel -> normalised, then fixed-point undistortion
       x <- (xd - tangential(x)) / radial(x), the published OpenCV undistortPoints iteration. */
    double fx = k[0], fy = k[1], cx = k[2], cy = k[3], k1 = k[4], k2 = k[5], p1 = k[6], p2 = k[7], k3 = k[8];
    double xd = (u - cx) / fx, yd = (v - cy) / fy;
    double x = xd, y = yd;
    bool ok = false;
    for (int it = 0; it < 100; it++) {
        double r2 = x * x + y * y;
        double radial = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3));
        double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
        double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
        if (radial == 0) return 0;
        double xn = (xd - dx) / radial, yn = (yd - dy) / radial;
        double e = (xn - x) * (xn - x) + (yn - y) * (yn - y);
        x = xn; y = yn;
        if (e < 1e-24) { ok = true; break; }
    }
    if (!ok || !std::isfinite(x) || !std::isfinite(y)) return 0;
    out[0] = x; out[1] = y; out[2] = 1.0;
    return 1;
}

}  // extern "C"
