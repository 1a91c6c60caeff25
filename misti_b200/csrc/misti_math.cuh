// misti_math.cuh -- scalar FP64 numerics of the coalescence-rate ("lambda") correction chain.
//
// Everything here is __host__ __device__ so that the very same code that the K1 kernel runs per
// item can be compiled with g++ into the test-only harness tests/hostsim (CPU unit tests in a
// GPU-less build container).  The product library only ever calls it from device code.
//
// What is restated (reference = Genomics-HSE/MiSTI, scipy 1.18.1 for the third-party parts):
//   * 3x3 matrix exponential / inverse        CorrectLambda.py:55-65  (scipy.linalg.expm / inv)
//   * least_squares(method='trf', jac='2-point', gtol=xtol=1e-10, ftol=1e-8, x_scale=1)
//       scipy/optimize/_lsq/least_squares.py:880-1044, _lsq/trf.py:206-587, _lsq/common.py,
//       _numdiff.py:14-92,147-190,683-768      (iterate-faithful port: same formulas, same
//       comparisons, same termination tests, so that it stops on the same iterate)
//   * the four residual systems               CorrectLambda.py:67-110,135-173,213-264
//   * SolveLambdaSystem / FitSinglePop        CorrectLambda.py:82-92,266-317
//   * CorrectLambdas / SmoothConst            MigrationInference.py:305-405
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define MISTI_HD __host__ __device__
#define MISTI_NOINLINE __noinline__
#else
#define MISTI_HD
#define MISTI_NOINLINE __attribute__((noinline))
#endif
// The residual functors are NOT inlined: the finite-difference Jacobian subtracts two evaluations of the same
// function, and scipy's result (an exactly zero column where the residual does not depend on a variable, which
// then stays at its start value) is only reproduced if both evaluations run the very same instruction sequence --
// inlined copies may be fused into FMAs differently at different call sites.

namespace misti {

constexpr double kEps = 2.220446049250313e-16;
constexpr double kSqrtEps = 1.4901161193847656e-08;  // EPS**0.5: default 2-point relative step
#define MISTI_INF_ (__builtin_huge_val())
constexpr double kInf = MISTI_INF_;

// ------------------------------------------------------------------------------------------
// 3x3 helpers (row-major double[9])
// ------------------------------------------------------------------------------------------
MISTI_HD inline void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

MISTI_HD inline void mat3_vec(const double* A, const double* x, double* y) {
    for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}

// Solve Q X = P (3x3, X overwrites P) by Gaussian elimination with partial pivoting.  The pivot row is picked with
// compares and the row exchange done with selects on fixed indices (no run-time subscripts), so that after inlining the
// two matrices live in registers; the arithmetic is that of the textbook loop (first largest pivot wins), bit for bit.
MISTI_HD inline bool mat3_solve(double* Q, double* P) {
    double q[3][3], p[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { q[i][j] = Q[3 * i + j]; p[i][j] = P[3 * i + j]; }
    // column 0: pivot among rows 0, 1, 2
    {
        int piv = 0;
        double best = fabs(q[0][0]);
        if (fabs(q[1][0]) > best) { best = fabs(q[1][0]); piv = 1; }
        if (fabs(q[2][0]) > best) { best = fabs(q[2][0]); piv = 2; }
        if (best == 0.0) return false;
        for (int j = 0; j < 3; ++j) {
            const double a0 = q[0][j], a1 = q[1][j], a2 = q[2][j];
            q[0][j] = piv == 1 ? a1 : (piv == 2 ? a2 : a0);
            q[1][j] = piv == 1 ? a0 : a1;
            q[2][j] = piv == 2 ? a0 : a2;
            const double b0 = p[0][j], b1 = p[1][j], b2 = p[2][j];
            p[0][j] = piv == 1 ? b1 : (piv == 2 ? b2 : b0);
            p[1][j] = piv == 1 ? b0 : b1;
            p[2][j] = piv == 2 ? b0 : b2;
        }
        const double inv = 1.0 / q[0][0];
        for (int i = 1; i < 3; ++i) {
            const double f = q[i][0] * inv;
            for (int j = 0; j < 3; ++j) q[i][j] -= f * q[0][j];
            for (int j = 0; j < 3; ++j) p[i][j] -= f * p[0][j];
        }
    }
    // column 1: pivot among rows 1, 2
    {
        const bool sw = fabs(q[2][1]) > fabs(q[1][1]);
        const double best = sw ? fabs(q[2][1]) : fabs(q[1][1]);
        if (best == 0.0) return false;
        for (int j = 0; j < 3; ++j) {
            const double a1 = q[1][j], a2 = q[2][j];
            q[1][j] = sw ? a2 : a1;
            q[2][j] = sw ? a1 : a2;
            const double b1 = p[1][j], b2 = p[2][j];
            p[1][j] = sw ? b2 : b1;
            p[2][j] = sw ? b1 : b2;
        }
        const double inv = 1.0 / q[1][1];
        const double f = q[2][1] * inv;
        for (int j = 1; j < 3; ++j) q[2][j] -= f * q[1][j];
        for (int j = 0; j < 3; ++j) p[2][j] -= f * p[1][j];
    }
    if (fabs(q[2][2]) == 0.0) return false;
    for (int k = 2; k >= 0; --k) {
        const double inv = 1.0 / q[k][k];
        for (int j = 0; j < 3; ++j) {
            double v = p[k][j];
            for (int i = k + 1; i < 3; ++i) v -= q[k][i] * p[i][j];
            p[k][j] = v * inv;
        }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { Q[3 * i + j] = q[i][j]; P[3 * i + j] = p[i][j]; }
    return true;
}

MISTI_HD inline bool mat3_inv(const double* A, double* Ainv) {
    double Q[9];
    for (int i = 0; i < 9; ++i) { Q[i] = A[i]; Ainv[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    return mat3_solve(Q, Ainv);
}

// exp(A) for a 3x3 matrix: Pade [m/m] with m in {3,5,7,9,13} chosen from ||A||_1, scaling and
// squaring (Higham 2005 thresholds; same algorithm family as scipy.linalg.expm).
// (inlined into its three callers: the 9-element arrays then live in registers instead of going through the stack)
MISTI_HD inline int mat3_pade(const double* Ain, double* U, double* V) {
    double A[9];
    double nrm = 0.0;
    for (int j = 0; j < 3; ++j) {
        const double c = fabs(Ain[j]) + fabs(Ain[3 + j]) + fabs(Ain[6 + j]);
        nrm = c > nrm ? c : nrm;
    }
    int s = 0;
    if (nrm > 5.371920351148152) {
        s = (int)ceil(log2(nrm / 5.371920351148152));
        if (s < 0) s = 0;
        if (s > 1000) s = 1000;
    }
    const double sc = ldexp(1.0, -s);
    for (int i = 0; i < 9; ++i) A[i] = Ain[i] * sc;
    double A2[9], T1[9];
    mat3_mul(A, A, A2);
    if (s == 0 && nrm <= 1.495585217958292e-2) {
        for (int i = 0; i < 9; ++i) { T1[i] = A2[i]; V[i] = 12.0 * A2[i]; }
        T1[0] += 60.0; T1[4] += 60.0; T1[8] += 60.0;
        V[0] += 120.0; V[4] += 120.0; V[8] += 120.0;
        mat3_mul(A, T1, U);
    } else if (s == 0 && nrm <= 2.539398330063230e-1) {
        double A4[9];
        mat3_mul(A2, A2, A4);
        for (int i = 0; i < 9; ++i) { T1[i] = A4[i] + 420.0 * A2[i]; V[i] = 30.0 * A4[i] + 3360.0 * A2[i]; }
        T1[0] += 15120.0; T1[4] += 15120.0; T1[8] += 15120.0;
        V[0] += 30240.0; V[4] += 30240.0; V[8] += 30240.0;
        mat3_mul(A, T1, U);
    } else if (s == 0 && nrm <= 9.504178996162932e-1) {
        double A4[9], A6[9];
        mat3_mul(A2, A2, A4);
        mat3_mul(A4, A2, A6);
        for (int i = 0; i < 9; ++i) {
            T1[i] = A6[i] + 1512.0 * A4[i] + 277200.0 * A2[i];
            V[i] = 56.0 * A6[i] + 25200.0 * A4[i] + 1995840.0 * A2[i];
        }
        T1[0] += 8648640.0; T1[4] += 8648640.0; T1[8] += 8648640.0;
        V[0] += 17297280.0; V[4] += 17297280.0; V[8] += 17297280.0;
        mat3_mul(A, T1, U);
    } else if (s == 0 && nrm <= 2.097847961257068) {
        double A4[9], A6[9], A8[9];
        mat3_mul(A2, A2, A4);
        mat3_mul(A4, A2, A6);
        mat3_mul(A6, A2, A8);
        for (int i = 0; i < 9; ++i) {
            T1[i] = A8[i] + 3960.0 * A6[i] + 2162160.0 * A4[i] + 302702400.0 * A2[i];
            V[i] = 90.0 * A8[i] + 110880.0 * A6[i] + 30270240.0 * A4[i] + 2075673600.0 * A2[i];
        }
        T1[0] += 8821612800.0; T1[4] += 8821612800.0; T1[8] += 8821612800.0;
        V[0] += 17643225600.0; V[4] += 17643225600.0; V[8] += 17643225600.0;
        mat3_mul(A, T1, U);
    } else {
        double A4[9], A6[9], W1[9], W2[9], Z[9];
        mat3_mul(A2, A2, A4);
        mat3_mul(A4, A2, A6);
        for (int i = 0; i < 9; ++i) {
            W1[i] = A6[i] + 16380.0 * A4[i] + 40840800.0 * A2[i];
            W2[i] = 182.0 * A6[i] + 960960.0 * A4[i] + 1323241920.0 * A2[i];
        }
        mat3_mul(A6, W1, Z);
        for (int i = 0; i < 9; ++i)
            T1[i] = Z[i] + 33522128640.0 * A6[i] + 10559470521600.0 * A4[i] + 1187353796428800.0 * A2[i];
        T1[0] += 32382376266240000.0; T1[4] += 32382376266240000.0; T1[8] += 32382376266240000.0;
        mat3_mul(A, T1, U);
        mat3_mul(A6, W2, Z);
        for (int i = 0; i < 9; ++i)
            V[i] = Z[i] + 670442572800.0 * A6[i] + 129060195264000.0 * A4[i] + 7771770303897600.0 * A2[i];
        V[0] += 64764752532480000.0; V[4] += 64764752532480000.0; V[8] += 64764752532480000.0;
    }
    return s;
}

MISTI_HD inline void mat3_expm(const double* Ain, double* E) {
    double U[9], V[9];
    const int s = mat3_pade(Ain, U, V);
    double T1[9];
    double Q[9];
    for (int i = 0; i < 9; ++i) { Q[i] = V[i] - U[i]; E[i] = V[i] + U[i]; }
    mat3_solve(Q, E);
    for (int k = 0; k < s; ++k) {
        mat3_mul(E, E, T1);
        for (int i = 0; i < 9; ++i) E[i] = T1[i];
    }
}

// Solve Q^T z = (1, 1, 1) (partial pivoting on the rows of Q^T, as mat3_solve): z^T = 1^T Q^-1
MISTI_HD inline bool mat3_solve_ones_t(const double* Q, double* z) {
    double q[3][3], p[3] = {1.0, 1.0, 1.0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) q[i][j] = Q[3 * j + i];
    {
        int piv = 0;
        double best = fabs(q[0][0]);
        if (fabs(q[1][0]) > best) { best = fabs(q[1][0]); piv = 1; }
        if (fabs(q[2][0]) > best) { best = fabs(q[2][0]); piv = 2; }
        if (best == 0.0) return false;
        for (int j = 0; j < 3; ++j) {
            const double a0 = q[0][j], a1 = q[1][j], a2 = q[2][j];
            q[0][j] = piv == 1 ? a1 : (piv == 2 ? a2 : a0);
            q[1][j] = piv == 1 ? a0 : a1;
            q[2][j] = piv == 2 ? a0 : a2;
        }
        const double inv = 1.0 / q[0][0];
        for (int i = 1; i < 3; ++i) {
            const double f = q[i][0] * inv;
            for (int j = 0; j < 3; ++j) q[i][j] -= f * q[0][j];
            p[i] -= f * p[0];  // (the right-hand side is all ones: the row exchange leaves it as it is)
        }
    }
    {
        const bool sw = fabs(q[2][1]) > fabs(q[1][1]);
        const double best = sw ? fabs(q[2][1]) : fabs(q[1][1]);
        if (best == 0.0) return false;
        for (int j = 0; j < 3; ++j) {
            const double a1 = q[1][j], a2 = q[2][j];
            q[1][j] = sw ? a2 : a1;
            q[2][j] = sw ? a1 : a2;
        }
        const double b1 = p[1], b2 = p[2];
        p[1] = sw ? b2 : b1;
        p[2] = sw ? b1 : b2;
        const double inv = 1.0 / q[1][1];
        const double f = q[2][1] * inv;
        for (int j = 1; j < 3; ++j) q[2][j] -= f * q[1][j];
        p[2] -= f * p[1];
    }
    if (fabs(q[2][2]) == 0.0) return false;
    for (int k = 2; k >= 0; --k) {
        double v = p[k];
        for (int i = k + 1; i < 3; ++i) v -= q[k][i] * p[i];
        p[k] = v * (1.0 / q[k][k]);
    }
    z[0] = p[0]; z[1] = p[1]; z[2] = p[2];
    return true;
}

// w = 1^T exp(A): what the cpfit residuals need of the exponential (the probability not to have coalesced, summed over the
// three states).  Without squarings (||A||_1 <= 5.37, the rule on the stretched unit interval) that is z^T (V + U) with
// Q^T z = 1, Q = V - U: one right-hand side instead of three and no matrix-vector products afterwards; with squarings the
// column sums of the full exponential.
MISTI_HD inline void mat3_expm_ones(const double* Ain, double* w) {
    double U[9], V[9];
    const int s = mat3_pade(Ain, U, V);
    if (s == 0) {
        double Q[9], Nn[9], z[3];
        for (int i = 0; i < 9; ++i) { Q[i] = V[i] - U[i]; Nn[i] = V[i] + U[i]; }
        mat3_solve_ones_t(Q, z);
        for (int j = 0; j < 3; ++j) w[j] = (z[0] * Nn[j] + z[1] * Nn[3 + j]) + z[2] * Nn[6 + j];
        return;
    }
    double Q[9], E[9], T1[9];
    for (int i = 0; i < 9; ++i) { Q[i] = V[i] - U[i]; E[i] = V[i] + U[i]; }
    mat3_solve(Q, E);
    for (int k = 0; k < s; ++k) {
        mat3_mul(E, E, T1);
        for (int i = 0; i < 9; ++i) E[i] = T1[i];
    }
    for (int j = 0; j < 3; ++j) w[j] = (E[j] + E[3 + j]) + E[6 + j];
}

// ------------------------------------------------------------------------------------------
// Thin SVD of an m x N matrix (N = 1 or 2), one-sided Jacobi.  B is column-major: B[c*MR + r].
// Returns singular values (descending), V (row-major N x N) and suf_i = s_i * (u_i . f).
// ------------------------------------------------------------------------------------------
template <int N, int MR>
MISTI_HD inline void thin_svd(const double* B, int m, const double* f, double* s, double* V, double* uf) {
    if constexpr (N == 1) {
        double n2 = 0, d = 0;
        for (int r = 0; r < m; ++r) { n2 += B[r] * B[r]; d += B[r] * f[r]; }
        s[0] = sqrt(n2);
        V[0] = 1.0;
        uf[0] = s[0] > 0 ? d / s[0] : 0.0;
    } else {
    double c1[MR], c2[MR];
    for (int r = 0; r < m; ++r) { c1[r] = B[r]; c2[r] = B[MR + r]; }
    double v00 = 1, v01 = 0, v10 = 0, v11 = 1;  // V columns: (v00,v10) and (v01,v11)
    for (int sweep = 0; sweep < 2; ++sweep) {  // the first rotation makes the two columns orthogonal to rounding, the second takes
                                               // the residue out; a third changes nothing but costs two square roots and two divisions
        double a = 0, b = 0, c = 0;
        for (int r = 0; r < m; ++r) { a += c1[r] * c1[r]; b += c1[r] * c2[r]; c += c2[r] * c2[r]; }
        if (b == 0.0) break;
        const double zeta = (c - a) / (2.0 * b);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int r = 0; r < m; ++r) {
            const double x = c1[r], y = c2[r];
            c1[r] = cs * x - sn * y;
            c2[r] = sn * x + cs * y;
        }
        const double a0 = v00, a1 = v10, b0 = v01, b1 = v11;
        v00 = cs * a0 - sn * b0; v10 = cs * a1 - sn * b1;
        v01 = sn * a0 + cs * b0; v11 = sn * a1 + cs * b1;
    }
    double n1 = 0, n2 = 0, d1 = 0, d2 = 0;
    for (int r = 0; r < m; ++r) { n1 += c1[r] * c1[r]; n2 += c2[r] * c2[r]; d1 += c1[r] * f[r]; d2 += c2[r] * f[r]; }
    n1 = sqrt(n1); n2 = sqrt(n2);
    if (n1 >= n2) {
        s[0] = n1; s[1] = n2;
        V[0] = v00; V[2] = v10; V[1] = v01; V[3] = v11;
        uf[0] = n1 > 0 ? d1 / n1 : 0.0; uf[1] = n2 > 0 ? d2 / n2 : 0.0;
    } else {
        s[0] = n2; s[1] = n1;
        V[0] = v01; V[2] = v11; V[1] = v00; V[3] = v10;
        uf[0] = n2 > 0 ? d2 / n2 : 0.0; uf[1] = n1 > 0 ? d1 / n1 : 0.0;
    }
    }
}

template <int N>
MISTI_HD inline double vnorm(const double* x) {
    double s = 0;
    for (int i = 0; i < N; ++i) s += x[i] * x[i];
    return sqrt(s);
}

// scipy/optimize/_lsq/common.py:57-168
template <int N>
MISTI_HD inline void solve_lsq_trust_region(int m, const double* uf, const double* s, const double* V, double Delta,
                                            double* alpha_io, double* p) {
    double suf[N];
    for (int i = 0; i < N; ++i) suf[i] = s[i] * uf[i];
    bool full_rank = false;
    if (m >= N) full_rank = s[N - 1] > kEps * m * s[0];
    if (full_rank) {
        double q[N];
        for (int i = 0; i < N; ++i) q[i] = uf[i] / s[i];
        for (int i = 0; i < N; ++i) {
            double v = 0;
            for (int j = 0; j < N; ++j) v += V[i * N + j] * q[j];
            p[i] = -v;
        }
        if (vnorm<N>(p) <= Delta) { *alpha_io = 0.0; return; }
    }
    double alpha_upper = vnorm<N>(suf) / Delta;
    double alpha_lower = 0.0;
    if (full_rank) {
        // phi_and_derivative(0.0, ...)
        double t[N], pn = 0, sum = 0;
        for (int i = 0; i < N; ++i) { const double den = s[i] * s[i]; t[i] = suf[i] / den; pn += t[i] * t[i]; }
        pn = sqrt(pn);
        for (int i = 0; i < N; ++i) { const double den = s[i] * s[i]; sum += suf[i] * suf[i] / (den * den * den); }
        const double phi = pn - Delta, phi_prime = -sum / pn;
        alpha_lower = -phi / phi_prime;
    }
    double alpha = *alpha_io;
    if (!full_rank && alpha == 0.0) {
        const double a = 0.001 * alpha_upper, b = sqrt(alpha_lower * alpha_upper);
        alpha = a > b ? a : b;
    }
    for (int it = 0; it < 10; ++it) {
        if (alpha < alpha_lower || alpha > alpha_upper) {
            const double a = 0.001 * alpha_upper, b = sqrt(alpha_lower * alpha_upper);
            alpha = a > b ? a : b;
        }
        double pn = 0, sum = 0;
        for (int i = 0; i < N; ++i) {
            const double den = s[i] * s[i] + alpha;
            const double t = suf[i] / den;
            pn += t * t;
            sum += suf[i] * suf[i] / (den * den * den);
        }
        pn = sqrt(pn);
        const double phi = pn - Delta, phi_prime = -sum / pn;
        if (phi < 0) alpha_upper = alpha;
        const double ratio = phi / phi_prime;
        const double cand = alpha - ratio;
        alpha_lower = alpha_lower > cand ? alpha_lower : cand;
        alpha -= (phi + Delta) * ratio / Delta;
        if (fabs(phi) < 0.01 * Delta) break;
    }
    double q[N];
    for (int i = 0; i < N; ++i) q[i] = suf[i] / (s[i] * s[i] + alpha);
    for (int i = 0; i < N; ++i) {
        double v = 0;
        for (int j = 0; j < N; ++j) v += V[i * N + j] * q[j];
        p[i] = -v;
    }
    const double sc = Delta / vnorm<N>(p);
    for (int i = 0; i < N; ++i) p[i] *= sc;
    *alpha_io = alpha;
}

// J is row-major m x N with m == N here.
template <int N>
MISTI_HD inline void jdot(const double* J, const double* s, double* out) {
    for (int r = 0; r < N; ++r) {
        double v = 0;
        for (int c = 0; c < N; ++c) v += J[r * N + c] * s[c];
        out[r] = v;
    }
}

// evaluate_quadratic (common.py), diag may be null
template <int N>
MISTI_HD inline double evaluate_quadratic(const double* J, const double* g, const double* s, const double* diag) {
    double Js[N];
    jdot<N>(J, s, Js);
    double q = 0, l = 0;
    for (int i = 0; i < N; ++i) q += Js[i] * Js[i];
    if (diag) {
        double e = 0;
        for (int i = 0; i < N; ++i) e += s[i] * diag[i] * s[i];
        q += e;
    }
    for (int i = 0; i < N; ++i) l += s[i] * g[i];
    return 0.5 * q + l;
}

// build_quadratic_1d (common.py)
template <int N>
MISTI_HD inline void build_quadratic_1d(const double* J, const double* g, const double* s, const double* diag,
                                        const double* s0, double* a, double* b, double* c) {
    double v[N];
    jdot<N>(J, s, v);
    double aa = 0;
    for (int i = 0; i < N; ++i) aa += v[i] * v[i];
    if (diag) {
        double e = 0;
        for (int i = 0; i < N; ++i) e += s[i] * diag[i] * s[i];
        aa += e;
    }
    aa *= 0.5;
    double bb = 0;
    for (int i = 0; i < N; ++i) bb += g[i] * s[i];
    if (s0) {
        double u[N];
        jdot<N>(J, s0, u);
        double uv = 0, uu = 0, gs0 = 0;
        for (int i = 0; i < N; ++i) { uv += u[i] * v[i]; uu += u[i] * u[i]; gs0 += g[i] * s0[i]; }
        bb += uv;
        double cc = 0.5 * uu + gs0;
        if (diag) {
            double e1 = 0, e2 = 0;
            for (int i = 0; i < N; ++i) { e1 += s0[i] * diag[i] * s[i]; e2 += s0[i] * diag[i] * s0[i]; }
            bb += e1;
            cc += 0.5 * e2;
        }
        *c = cc;
    }
    *a = aa;
    *b = bb;
}

// minimize_quadratic_1d (common.py): candidates lb, ub, interior extremum; first minimum wins.
MISTI_HD inline void minimize_quadratic_1d(double a, double b, double lb, double ub, double c, double* t_out, double* y_out) {
    double t = lb, y = lb * (a * lb + b) + c;
    const double yu = ub * (a * ub + b) + c;
    if (yu < y) { y = yu; t = ub; }
    if (a != 0) {
        const double ex = -0.5 * b / a;
        if (lb < ex && ex < ub) {
            const double ye = ex * (a * ex + b) + c;
            if (ye < y) { y = ye; t = ex; }
        }
    }
    *t_out = t;
    *y_out = y;
}

// step_size_to_bound with ub = +inf (common.py)
template <int N>
MISTI_HD inline double step_size_to_bound(const double* x, const double* s, double lb, int* hits) {
    double steps[N];
    double mn = kInf;
    for (int i = 0; i < N; ++i) {
        if (s[i] == 0) steps[i] = kInf;
        else if (s[i] > 0) steps[i] = kInf;           // max((lb-x)/s, +inf)
        else steps[i] = (lb - x[i]) / s[i];          // max((lb-x)/s, -inf)
        if (steps[i] < mn) mn = steps[i];
    }
    for (int i = 0; i < N; ++i) hits[i] = (steps[i] == mn) ? (s[i] > 0 ? 1 : (s[i] < 0 ? -1 : 0)) : 0;
    return mn;
}

// make_strictly_feasible with ub = +inf (common.py:440-465)
template <int N>
MISTI_HD inline void make_strictly_feasible(double* x, double lb, double rstep) {
    for (int i = 0; i < N; ++i) {
        if (rstep == 0) {
            if (x[i] <= lb) x[i] = nextafter(lb, kInf);
        } else {
            const double thr = rstep * (fabs(lb) > 1 ? fabs(lb) : 1.0);
            if (x[i] - lb <= thr) x[i] = lb + thr;
        }
        // tight bounds can not happen with ub = +inf
    }
}

// ------------------------------------------------------------------------------------------
// least_squares(fun, x0, bounds=(lb, +inf) or unbounded, method='trf', jac='2-point',
//               ftol=1e-8, xtol=gtol=1e-10).  m == n == N.  Fun: bool operator()(const double*, double*)
// returns false if it cannot be evaluated (treated like a non-finite residual).
// Return value: scipy termination status (0..4), or -1 when scipy would have raised.
// ------------------------------------------------------------------------------------------
template <int N, class Fun>
MISTI_HD inline bool fd_jacobian(Fun& fun, const double* x, const double* f0, bool bounded, double lb, double* J, int* njev_calls) {
    for (int i = 0; i < N; ++i) {
        double h = kSqrtEps * (x[i] >= 0 ? 1.0 : -1.0) * (fabs(x[i]) > 1.0 ? fabs(x[i]) : 1.0);
        if (bounded) {  // _adjust_scheme_to_bounds, '1-sided', ub = +inf
            const double xt = x[i] + h;
            const bool violated = xt < lb;
            // upper_dist = +inf  =>  fitting is always true
            if (violated) h = -h;
        }
        double x1[N], f1[N];
        for (int k = 0; k < N; ++k) x1[k] = x[k];
        x1[i] = x[i] + h;
        const double dx = x1[i] - x[i];
        fun(x1, f1);
        ++*njev_calls;
        for (int r = 0; r < N; ++r) J[r * N + i] = (f1[r] - f0[r]) / dx;
    }
    return true;
}

template <int N>
MISTI_HD inline bool all_finite(const double* f) {
    for (int i = 0; i < N; ++i)
        if (!(fabs(f[i]) <= DBL_MAX)) return false;
    return true;
}

// Residuals f(x) and their 2-point Jacobian.  COOP (device only): the item is run by FOUR lanes in lock step (all of them
// execute the same code on the same data); here lane 0 of the quad evaluates f(x) and lanes 1..N the shifted points, at
// the same time, and the results are exchanged by shuffles -- the N + 1 sequential evaluations of a solver round become
// one.  Every evaluation still runs the out-of-line functor, so the numbers are those of the one-thread path, bit for bit.
template <int N, bool COOP, class Fun>
MISTI_HD inline void eval_fj(Fun& fun, const double* x, bool bounded, double lb, double* f, double* J, int* njev_calls) {
#if defined(__CUDA_ARCH__)
    if constexpr (COOP) {
        const unsigned lane = threadIdx.x & 31u, q = lane & 3u, base = lane & ~3u;
        const unsigned mask = 0xFu << base;
        double xq[N], fq[N], dx[N];
        for (int i = 0; i < N; ++i) {
            double h = kSqrtEps * (x[i] >= 0 ? 1.0 : -1.0) * (fabs(x[i]) > 1.0 ? fabs(x[i]) : 1.0);
            if (bounded && x[i] + h < lb) h = -h;
            xq[i] = x[i];
            const double xs = x[i] + h;
            dx[i] = xs - x[i];
            if (q == (unsigned)(i + 1)) xq[i] = xs;
        }
        fun(xq, fq);
        for (int r = 0; r < N; ++r) f[r] = __shfl_sync(mask, fq[r], base);
        for (int i = 0; i < N; ++i)
            for (int r = 0; r < N; ++r) {
                const double f1 = __shfl_sync(mask, fq[r], base + i + 1);
                J[r * N + i] = (f1 - f[r]) / dx[i];
            }
        *njev_calls += N;
        return;
    }
#endif
    fun(x, f);
    fd_jacobian<N>(fun, x, f, bounded, lb, J, njev_calls);
}

template <int N, bool COOP = false, class Fun>
MISTI_HD inline int least_squares_trf(Fun& fun, double* x, bool bounded, double lb, int* nfev_out) {
    const double ftol = 1e-8, xtol = 1e-10, gtol = 1e-10;
    const int max_nfev = 100 * N;
    int fd_calls = 0;
    if (bounded) {
        for (int i = 0; i < N; ++i)
            if (!(x[i] >= lb)) { *nfev_out = 0; return -1; }  // "Initial guess is outside of provided bounds"
        make_strictly_feasible<N>(x, lb, 1e-10);
    }
    double f[N], J[N * N], g[N];
    [[maybe_unused]] double Jn[N * N];  // COOP: Jacobian at the trial point, evaluated together with its residuals
    if constexpr (COOP) {
        eval_fj<N, true>(fun, x, bounded, lb, f, J, &fd_calls);
        if (!all_finite<N>(f)) { *nfev_out = 1; return -1; }
    } else {
        fun(x, f);
        if (!all_finite<N>(f)) { *nfev_out = 1; return -1; }  // "Residuals are not finite in the initial point"
        fd_jacobian<N>(fun, x, f, bounded, lb, J, &fd_calls);
    }
    int nfev = 1;
    double cost = 0;
    for (int i = 0; i < N; ++i) cost += f[i] * f[i];
    cost *= 0.5;
    for (int c = 0; c < N; ++c) {
        double v = 0;
        for (int r = 0; r < N; ++r) v += J[r * N + c] * f[r];
        g[c] = v;
    }
    double v[N], dv[N];
    double Delta;
    {
        double t = 0;
        for (int i = 0; i < N; ++i) {
            double vi = 1.0;
            if (bounded && g[i] > 0) vi = x[i] - lb;
            const double q = bounded ? x[i] / sqrt(vi) : x[i];
            t += q * q;
        }
        Delta = sqrt(t);
        if (Delta == 0) Delta = 1.0;
    }
    double alpha = 0.0;
    int status = -2;  // None
    double xn[N], fn[N];
    while (true) {
        double g_norm = 0;
        for (int i = 0; i < N; ++i) {
            v[i] = 1.0; dv[i] = 0.0;
            if (bounded && g[i] > 0) { v[i] = x[i] - lb; dv[i] = 1.0; }
            const double a = fabs(g[i] * v[i]);
            if (a > g_norm) g_norm = a;
        }
        if (g_norm < gtol) status = 1;
        if (status != -2 || nfev == max_nfev) break;

        double d[N], diag_h[N], g_h[N], Jh[N * N];
        for (int i = 0; i < N; ++i) {
            d[i] = bounded ? sqrt(v[i]) : 1.0;
            diag_h[i] = g[i] * dv[i];
            g_h[i] = d[i] * g[i];
        }
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < N; ++c) Jh[r * N + c] = J[r * N + c] * d[c];
        double s[N], V[N * N], uf[N];
        if (bounded) {
            double B[N * 2 * N], fa[2 * N];  // column-major (2N rows)
            for (int c = 0; c < N; ++c) {
                for (int r = 0; r < N; ++r) B[c * 2 * N + r] = Jh[r * N + c];
                for (int r = 0; r < N; ++r) B[c * 2 * N + N + r] = (r == c) ? sqrt(diag_h[c]) : 0.0;
            }
            for (int r = 0; r < N; ++r) { fa[r] = f[r]; fa[N + r] = 0.0; }
            thin_svd<N, 2 * N>(B, 2 * N, fa, s, V, uf);
        } else {
            double B[N * N];
            for (int c = 0; c < N; ++c)
                for (int r = 0; r < N; ++r) B[c * N + r] = Jh[r * N + c];
            thin_svd<N, N>(B, N, f, s, V, uf);
        }
        const double theta = (0.995 > 1 - g_norm) ? 0.995 : 1 - g_norm;
        double actual_reduction = -1;
        double cost_new = cost;
        while (actual_reduction <= 0 && nfev < max_nfev) {
            double p_h[N], step[N], step_h[N];
            solve_lsq_trust_region<N>(N, uf, s, V, Delta, &alpha, p_h);
            double predicted;
            if (!bounded) {
                predicted = -evaluate_quadratic<N>(Jh, g_h, p_h, nullptr);
                for (int i = 0; i < N; ++i) { step_h[i] = p_h[i]; step[i] = d[i] * p_h[i]; xn[i] = x[i] + step[i]; }
            } else {
                // ---- select_step (trf.py:129-203) with ub = +inf
                double p[N];
                for (int i = 0; i < N; ++i) p[i] = d[i] * p_h[i];
                bool inb = true;
                for (int i = 0; i < N; ++i)
                    if (!(x[i] + p[i] >= lb)) inb = false;
                if (inb) {
                    predicted = -evaluate_quadratic<N>(Jh, g_h, p_h, diag_h);
                    for (int i = 0; i < N; ++i) { step[i] = p[i]; step_h[i] = p_h[i]; }
                } else {
                    int hits[N];
                    const double p_stride = step_size_to_bound<N>(x, p, lb, hits);
                    double r_h[N], r[N], xb[N];
                    for (int i = 0; i < N; ++i) {
                        r_h[i] = hits[i] != 0 ? -p_h[i] : p_h[i];
                        r[i] = d[i] * r_h[i];
                        p[i] *= p_stride;
                        p_h[i] *= p_stride;
                        xb[i] = x[i] + p[i];
                    }
                    // intersect_trust_region(p_h, r_h, Delta) -> positive root
                    double to_tr;
                    {
                        double a = 0, b = 0, c = 0;
                        for (int i = 0; i < N; ++i) { a += r_h[i] * r_h[i]; b += p_h[i] * r_h[i]; c += p_h[i] * p_h[i]; }
                        c -= Delta * Delta;
                        const double dd = sqrt(b * b - a * c);
                        const double q = -(b + (b >= 0 ? dd : -dd));   // copysign(d, b); -0.0 is not expected here
                        const double t1 = q / a, t2 = c / q;
                        to_tr = t1 < t2 ? t2 : t1;
                    }
                    int hits2[N];
                    const double to_bound = step_size_to_bound<N>(xb, r, lb, hits2);
                    double r_stride = to_bound < to_tr ? to_bound : to_tr;
                    double r_stride_l, r_stride_u;
                    if (r_stride > 0) {
                        r_stride_l = (1 - theta) * p_stride / r_stride;
                        r_stride_u = (r_stride == to_bound) ? theta * to_bound : to_tr;
                    } else {
                        r_stride_l = 0; r_stride_u = -1;
                    }
                    double r_value;
                    if (r_stride_l <= r_stride_u) {
                        double a, b, c;
                        build_quadratic_1d<N>(Jh, g_h, r_h, diag_h, p_h, &a, &b, &c);
                        minimize_quadratic_1d(a, b, r_stride_l, r_stride_u, c, &r_stride, &r_value);
                        for (int i = 0; i < N; ++i) { r_h[i] = r_h[i] * r_stride + p_h[i]; r[i] = r_h[i] * d[i]; }
                    } else {
                        r_value = kInf;
                    }
                    for (int i = 0; i < N; ++i) { p[i] *= theta; p_h[i] *= theta; }
                    const double p_value = evaluate_quadratic<N>(Jh, g_h, p_h, diag_h);
                    double ag_h[N], ag[N];
                    for (int i = 0; i < N; ++i) { ag_h[i] = -g_h[i]; ag[i] = d[i] * ag_h[i]; }
                    double to_tr2 = Delta / vnorm<N>(ag_h);
                    int hits3[N];
                    const double to_bound2 = step_size_to_bound<N>(x, ag, lb, hits3);
                    double ag_stride = (to_bound2 < to_tr2) ? theta * to_bound2 : to_tr2;
                    double a, b, cdum, ag_value;
                    build_quadratic_1d<N>(Jh, g_h, ag_h, diag_h, nullptr, &a, &b, &cdum);
                    minimize_quadratic_1d(a, b, 0.0, ag_stride, 0.0, &ag_stride, &ag_value);
                    for (int i = 0; i < N; ++i) { ag_h[i] *= ag_stride; ag[i] *= ag_stride; }
                    if (p_value < r_value && p_value < ag_value) {
                        for (int i = 0; i < N; ++i) { step[i] = p[i]; step_h[i] = p_h[i]; }
                        predicted = -p_value;
                    } else if (r_value < p_value && r_value < ag_value) {
                        for (int i = 0; i < N; ++i) { step[i] = r[i]; step_h[i] = r_h[i]; }
                        predicted = -r_value;
                    } else {
                        for (int i = 0; i < N; ++i) { step[i] = ag[i]; step_h[i] = ag_h[i]; }
                        predicted = -ag_value;
                    }
                }
                for (int i = 0; i < N; ++i) xn[i] = x[i] + step[i];
                make_strictly_feasible<N>(xn, lb, 0.0);
            }
            if constexpr (COOP) eval_fj<N, true>(fun, xn, bounded, lb, fn, Jn, &fd_calls);
            else fun(xn, fn);
            ++nfev;
            const double step_h_norm = vnorm<N>(step_h);
            if (!all_finite<N>(fn)) {
                Delta = 0.25 * step_h_norm;
                continue;
            }
            cost_new = 0;
            for (int i = 0; i < N; ++i) cost_new += fn[i] * fn[i];
            cost_new *= 0.5;
            actual_reduction = cost - cost_new;
            // update_tr_radius (common.py:222-248)
            double ratio;
            if (predicted > 0) ratio = actual_reduction / predicted;
            else if (predicted == 0 && actual_reduction == 0) ratio = 1;
            else ratio = 0;
            double Delta_new = Delta;
            if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
            else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = Delta * 2.0;
            const double step_norm = vnorm<N>(step);
            // check_termination (common.py:705-717)
            const bool ftol_ok = actual_reduction < ftol * cost && ratio > 0.25;
            const bool xtol_ok = step_norm < xtol * (xtol + vnorm<N>(x));
            if (ftol_ok && xtol_ok) status = 4;
            else if (ftol_ok) status = 2;
            else if (xtol_ok) status = 3;
            if (status != -2) break;
            alpha *= Delta / Delta_new;
            Delta = Delta_new;
        }
        if (actual_reduction > 0) {
            for (int i = 0; i < N; ++i) { x[i] = xn[i]; f[i] = fn[i]; }
            cost = cost_new;
            if constexpr (COOP) { for (int i = 0; i < N * N; ++i) J[i] = Jn[i]; }
            else fd_jacobian<N>(fun, x, f, bounded, lb, J, &fd_calls);
            for (int c = 0; c < N; ++c) {
                double t = 0;
                for (int r = 0; r < N; ++r) t += J[r * N + c] * f[r];
                g[c] = t;
            }
        }
    }
    if (status == -2) status = 0;
    *nfev_out = nfev;
    return status;
}

// ------------------------------------------------------------------------------------------
// One interval of the correction (CorrectLambda.py)
// ------------------------------------------------------------------------------------------
struct IntervalState {
    double lh[2];     // PSMC-apparent rates of the two genomes on this interval
    double T;         // interval length
    double mu[2];     // migration rates
    double P0[2][3];  // per genome: P(both lineages in deme 0 / deme 1 / one each), not yet coalesced
    double nch[2];    // cpfit target of the interval, exp(-lh T) (sum of P0): the same in every residual evaluation
};
constexpr int kNoSolve = -9;  // "termination status" of an interval that has a closed form (no least-squares solve)

MISTI_HD inline void corr_matrix(const double* l, const double* mu, double T, double* M) {
    // CorrectLambda.SetMatrix (CorrectLambda.py:55-56), times T
    M[0] = (-2 * mu[0] - l[0]) * T; M[1] = 0.0;                       M[2] = mu[1] * T;
    M[3] = 0.0;                       M[4] = (-2 * mu[1] - l[1]) * T; M[5] = mu[0] * T;
    M[6] = 2 * mu[0] * T;             M[7] = 2 * mu[1] * T;           M[8] = (-mu[0] - mu[1]) * T;
}

MISTI_HD inline double one_pop_time(double lam, double T) {  // ExpectedCoalTimeOnePop, CorrectLambda.py:67-72
    double r;
    if (lam > 100) r = 0;
    else r = T / (exp(lam * T) - 1);
    return 1.0 / lam - r;
}

MISTI_HD inline double one_pop_time_noncond(double lam, double T) {  // CorrectLambda.py:79-80
    return (1 - exp(-lam * T) * (1 + lam * T)) / lam;
}

// cpfit residuals: LambdaSystem1 / LambdaEquation (CorrectLambda.py:135-144,169-173)
struct ResidualProb {
    const IntervalState* st;
    MISTI_HD MISTI_NOINLINE bool operator()(const double* l, double* out) const {
        double M[9], w[3];
        corr_matrix(l, st->mu, st->T, M);
        mat3_expm_ones(M, w);
        for (int k = 0; k < 2; ++k) {
            const double* P = st->P0[k];
            out[k] = ((w[0] * P[0] + w[1] * P[1]) + w[2] * P[2]) - st->nch[k];
        }
        return true;
    }
};

// default-mode residuals: LambdaSystem / ExpectedCoalTimeTwoPop (CorrectLambda.py:94-110,151-157)
struct ResidualTime {
    const IntervalState* st;
    MISTI_HD MISTI_NOINLINE bool operator()(const double* l, double* out) const {
        const double T = st->T;
        double M[9], MT[9], E[9], Minv[9];
        corr_matrix(l, st->mu, 1.0, M);
        for (int i = 0; i < 9; ++i) MT[i] = M[i] * T;
        mat3_expm(MT, E);
        if (!mat3_inv(M, Minv)) { out[0] = out[1] = kInf - kInf; return false; }
        for (int k = 0; k < 2; ++k) {
            const double* P = st->P0[k];
            const double sP = (P[0] + P[1]) + P[2];
            double Pn[3] = {P[0] / sP, P[1] / sP, P[2] / sP};
            double EmI[9];
            for (int i = 0; i < 9; ++i) EmI[i] = E[i] - ((i % 4 == 0) ? 1.0 : 0.0);
            double v1[3], t1[3], v2[3], t2[3];
            mat3_vec(EmI, Pn, t1);
            mat3_vec(Minv, t1, v1);
            mat3_vec(Minv, v1, t1);      // Minv Minv (E-I) Pn
            mat3_vec(E, Pn, v2);
            const double pnc = (v2[0] + v2[1]) + v2[2];
            mat3_vec(Minv, v2, t2);
            const double vec0 = T * t2[0] - t1[0], vec1 = T * t2[1] - t1[1];
            const double ct = (l[0] * vec0 + l[1] * vec1) / (1 - pnc);
            const double lam = st->lh[k];
            const double pn1 = exp(-lam * T);                       // ExpectedCoalTimeOnePopTmp :74-77
            const double Tc = 1.0 / lam - T / (1.0 / pn1 - 1.0);
            out[k] = ct - Tc;
        }
        return true;
    }
};

// default mode, zero migration: LambdaSystemNoMigration (CorrectLambda.py:237-251)
struct ResidualNoMig {
    const IntervalState* st;
    double pr0[2][3];
    MISTI_HD MISTI_NOINLINE bool operator()(const double* l, double* out) const {
        const double T = st->T;
        for (int i = 0; i < 2; ++i) {
            const double pnc = (pr0[i][0] * exp(-l[0] * T) + pr0[i][1] * exp(-l[1] * T)) + pr0[i][2];
            const double ct = (pr0[i][0] * one_pop_time_noncond(l[0], T) + pr0[i][1] * one_pop_time_noncond(l[1], T)) / (1 - pnc);
            out[i] = ct - one_pop_time(st->lh[i], T);
        }
        return true;
    }
};

// post-split single rate: EPSFromExpectedCoalTime (CorrectLambda.py:82-86)
struct ResidualSingle {
    double T, Te;
    MISTI_HD MISTI_NOINLINE bool operator()(const double* lam, double* out) const {
        out[0] = one_pop_time(lam[0], T) - Te;
        return true;
    }
};

// Per-interval constants of a grid, shared by every item that uses the grid (computed once on the host when the grid
// is registered): exp(-lh_g T), 1 - exp(-lh_g T) for the two genomes, and 1/T.
constexpr int kGridAux = 5;
MISTI_HD inline void grid_aux_row(const double* lh2, double T, double* out) {
    out[0] = exp(-lh2[0] * T); out[1] = exp(-lh2[1] * T);
    out[2] = -expm1(-lh2[0] * T); out[3] = -expm1(-lh2[1] * T);
    out[4] = 1.0 / T;
}

// The interval WITH migration (CorrectLambda.py:276-317): 2-unknown trust-region solve on the 3-state chains.
// (Inlined: an out-of-line version measured 5 % slower on B200.)
template <bool COOP = false>
MISTI_HD inline bool solve_interval_mig(IntervalState* st, bool cpfit, double* lc, int* nfev, int* status_out = nullptr) {
    const double T = st->T;
    double (*P0)[3] = st->P0;
    double lh[2] = {st->lh[0], st->lh[1]};
    {
        double nV0 = 0, nV1 = 0, nD = 0;
        for (int i = 0; i < 3; ++i) {
            nV0 += P0[0][i] * P0[0][i]; nV1 += P0[1][i] * P0[1][i];
            const double d = P0[0][i] - P0[1][i]; nD += d * d;
        }
        nV0 = sqrt(nV0); nV1 = sqrt(nV1); nD = sqrt(nD);
        if (nD < 0.02 * (nV0 < nV1 ? nV0 : nV1)) { const double a = (lh[0] + lh[1]) / 2.0; lh[0] = lh[1] = a; }
    }
    // "stretch" to the unit interval (:293-298)
    IntervalState u = *st;
    u.T = T / T;
    u.mu[0] = st->mu[0] * T; u.mu[1] = st->mu[1] * T;
    u.lh[0] = lh[0] * T; u.lh[1] = lh[1] * T;
    double x[2] = {u.lh[0], u.lh[1]};
    int nf = 0, status;
    for (int k = 0; k < 2; ++k) u.nch[k] = exp(-u.lh[k] * u.T) * ((P0[k][0] + P0[k][1]) + P0[k][2]);
    if (cpfit) { ResidualProb fun; fun.st = &u; status = least_squares_trf<2, COOP>(fun, x, false, -kInf, &nf); }
    else { ResidualTime fun; fun.st = &u; status = least_squares_trf<2, COOP>(fun, x, false, -kInf, &nf); }
    *nfev += nf;
    if (status_out) *status_out = status;
    if (status < 0) return false;
    // un-stretch exactly as the reference does: mu*T/T, x/T
    const double mu_back[2] = {u.mu[0] / T, u.mu[1] / T};
    lc[0] = x[0] / T; lc[1] = x[1] / T;
    double M[9], E[9];
    corr_matrix(lc, mu_back, T, M);
    mat3_expm(M, E);
    for (int k = 0; k < 2; ++k) {
        double p[3];
        mat3_vec(E, P0[k], p);
        P0[k][0] = p[0]; P0[k][1] = p[1]; P0[k][2] = p[2];
    }
    return lc[0] > 0 && lc[1] > 0;
}

// SolveLambdaSystem (CorrectLambda.py:266-317).  On return lc[2] and st->P0 (advanced through the
// interval).  Returns false when the reference would report a failed correction or crash.
// `ga` (nullable): grid_aux_row of this interval.
template <bool COOP = false>
MISTI_HD inline bool solve_interval(IntervalState* st, bool cpfit, double mixtureTH, double* lc, int* nfev, const double* ga = nullptr,
                                    int* status_out = nullptr) {  // status_out (nullable): scipy's termination status of the solve, kNoSolve if none
    const double T = st->T;
    double (*P0)[3] = st->P0;
    const double s0 = (P0[0][0] + P0[0][1]) + P0[0][2], s1 = (P0[1][0] + P0[1][1]) + P0[1][2];
    if (status_out) *status_out = kNoSolve;
    if (mixtureTH > 0) {  // sqrt(mix) >= 0: the test can only fire for a positive threshold
        double mix = 0;
        for (int i = 0; i < 3; ++i) { const double d = P0[0][i] / s0 - P0[1][i] / s1; mix += d * d; }
        if (sqrt(mix) < mixtureTH) { lc[0] = lc[1] = -1; return false; }
    }
    if (st->mu[0] + st->mu[1] < 1e-10) {
        if (cpfit) {
            // SolveNoMigration1 (:213-235): find (a0, a1) = (exp(-l0 T), exp(-l1 T)) with
            //   P0[k][0] a0 + P0[k][1] a1 + P0[k][2] = exp(-lh_k T) s_k   for both genomes k.
            // The reference normalises by s_k and inverts the 2x2 matrix entry by entry (12 divisions); this is the same
            // linear system solved with one reciprocal, and the state is advanced with a0, a1 themselves.
            const double E0 = ga ? ga[0] : exp(-st->lh[0] * T), E1 = ga ? ga[1] : exp(-st->lh[1] * T);
            const double invT = ga ? ga[4] : 1.0 / T;
            const double Y0 = E0 * s0 - P0[0][2], Y1 = E1 * s1 - P0[1][2];
            const double det = P0[0][0] * P0[1][1] - P0[0][1] * P0[1][0];
            const double rdet = 1.0 / det;
            const double a0 = (P0[1][1] * Y0 - P0[0][1] * Y1) * rdet, a1 = (P0[0][0] * Y1 - P0[1][0] * Y0) * rdet;
            if (!(a0 > 0 && a1 > 0)) { lc[0] = lc[1] = -1; return false; }
            lc[0] = -log(a0) * invT; lc[1] = -log(a1) * invT;
            for (int k = 0; k < 2; ++k) { P0[k][0] *= a0; P0[k][1] *= a1; }
            return lc[0] > 0 && lc[1] > 0;
        }
        // SolveNoMigration (:253-264)
        ResidualNoMig fun;
        fun.st = st;
        for (int i = 0; i < 3; ++i) { fun.pr0[0][i] = P0[0][i] / s0; fun.pr0[1][i] = P0[1][i] / s1; }
        const double lb = 0.01 * (st->lh[0] < st->lh[1] ? st->lh[0] : st->lh[1]);
        double x[2] = {st->lh[0], st->lh[1]};
        int nf = 0;
        const int status = least_squares_trf<2, COOP>(fun, x, true, lb, &nf);
        *nfev += nf;
        if (status_out) *status_out = status;
        if (status < 0) return false;
        lc[0] = x[0]; lc[1] = x[1];
        const double e0 = exp(-lc[0] * T), e1 = exp(-lc[1] * T);
        for (int k = 0; k < 2; ++k) { P0[k][0] *= e0; P0[k][1] *= e1; }
        return lc[0] > 0 && lc[1] > 0;
    }
    return solve_interval_mig<COOP>(st, cpfit, lc, nfev, status_out);
}

// FitSinglePop (CorrectLambda.py:88-92) with P0 = [[exp(nc0),0,0],[exp(nc1),0,0]]
template <bool COOP = false>
MISTI_HD inline bool fit_single_pop(const double* lh, double T, double nc0, double nc1, double* lam, int* nfev, int* status_out = nullptr) {
    double p0 = exp(nc0), p1 = exp(nc1);
    const double sp = p0 + p1;
    p0 = p0 / sp; p1 = p1 / sp;
    ResidualSingle fun;
    fun.T = T;
    fun.Te = p0 * one_pop_time(lh[0], T) + p1 * one_pop_time(lh[1], T);
    double x[1] = {p0 * lh[0] + p1 * lh[1]};
    const double lb = 0.01 * (lh[0] < lh[1] ? lh[0] : lh[1]);
    int nf = 0;
    const int status = least_squares_trf<1, COOP>(fun, x, true, lb, &nf);
    *nfev += nf;
    if (status_out) *status_out = status;
    if (status < 0) return false;
    *lam = x[0];
    return true;
}

}  // namespace misti
