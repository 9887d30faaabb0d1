// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// PINNED BY EXECUTION: the reference (snowiow/SIFT) ships no tests or golden vectors and cannot be built here (Vigra,
// OpenCV-C++, Boost are absent), but its shipped executable carries Vigra's own inverse / linearSolve: oracle/refbin_run.cpp
// calls Sift::_eliminateEdgeResponses and alg::vertexParabola of that executable in place, and tests/test_refbin_pin.py holds
// this file to them on every interior pixel of random DoG stacks (ties, singular Hessians included) and on whole images.
// This file restates the fp32 linear
// algebra of Vigra 1.11 (un-vendored dependency of the reference; SONAME libvigraimpex.so.11)
// that the reference calls from
//   sift.cpp:306        linalg::inverse(neg_sec_deriv, inverse_matrix)
//   sift.cpp:311        linalg::linearSolve(inverse_matrix, deriv, extremum)      (method "QR")
//   sift.cpp:322        linalg::dot(deriv_transpose, extremum)
//   algorithms.cpp:175  linalg::linearSolve(a, b, res)                            (rank deficient)
// The structure (Householder QR, optional column pivoting, rank estimate from the diagonal of
// R for n < 4, minimum-norm least squares when rank < n) follows the published Vigra algorithm
// (vigra/linear_solve.hxx) and was cross-checked against the instruction sequences of the
// reference's own shipped binary bin/arch_x64/sift:
//   norm()                          @0x4174c0  sqrtf of a sequential fp32 sum of squares
//   dot()                           @0x41c1d0  sequential fp32 mul + add
//   qrHouseholderStepImpl           @0x41c9c0  vnorm sign, f = sqrt(vnorm*(vnorm-v0)), col -= (dot*u)
//   qrTransformToTriangularImpl     @0x41e040  pivot = first strict max; norm downdate uses row k;
//                                              tol = (float)m*max|r_kk|*FLT_EPSILON, compared in double
//   linearSolveQRReplace            @0x41f1c0  rank<n: lower-triangular re-factorisation w/o pivoting
//   linearSolveUpper/LowerTriangular@0x41a1b0 / @0x419fa0
//   applyHouseholderColumnReflections @0x41c6f0
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

namespace oracle {
namespace la {

// Strided 2-D view: (i, j) = row i, column j.
struct MV {
    float* p;
    long s0, s1;  // strides
    long n0, n1;  // rows, columns
    MV() : p(nullptr), s0(0), s1(0), n0(0), n1(0) {}
    MV(float* p_, long s0_, long s1_, long n0_, long n1_) : p(p_), s0(s0_), s1(s1_), n0(n0_), n1(n1_) {}
    float& operator()(long i, long j) const { return p[i * s0 + j * s1]; }
    long rows() const { return n0; }
    long cols() const { return n1; }
    MV T() const { return MV{p, s1, s0, n1, n0}; }
    // half-open [i0,i1) x [j0,j1)
    MV sub(long i0, long j0, long i1, long j1) const {
        return MV{p + i0 * s0 + j0 * s1, s0, s1, i1 - i0, j1 - j0};
    }
    // column j, rows [i0, i1)
    MV colv(long i0, long j, long i1) const { return sub(i0, j, i1, j + 1); }
};

// Dense owner, column-major like vigra::Matrix (storage order is unobservable here: every
// reduction below runs along a 1-D vector in index order).
struct Mat {
    long r = 0, c = 0;
    std::vector<float> a;
    Mat() {}
    Mat(long r_, long c_) : r(r_), c(c_), a((size_t)(r_ * c_), 0.0f) {}
    MV view() { return MV{a.data(), 1, r, r, c}; }
    float& operator()(long i, long j) { return a[(size_t)(j * r + i)]; }
    float operator()(long i, long j) const { return a[(size_t)(j * r + i)]; }
};

inline Mat copy_of(const MV& v) {
    Mat m(v.rows(), v.cols());
    for (long j = 0; j < v.cols(); ++j)
        for (long i = 0; i < v.rows(); ++i) m(i, j) = v(i, j);
    return m;
}

// fp32 sequential sum of squares over a vector-shaped view.
inline float squared_norm_vec(const MV& v) {
    float s = 0.0f;
    for (long j = 0; j < v.cols(); ++j)
        for (long i = 0; i < v.rows(); ++i) {
            float e = v(i, j);
            s += e * e;
        }
    return s;
}
inline float norm_vec(const MV& v) { return std::sqrt(squared_norm_vec(v)); }

// dot of two vector-shaped views with equal element count, index order.
inline float dot_vec(const MV& x, const MV& y) {
    long n = x.rows() * x.cols();
    float ret = 0.0f;
    for (long i = 0; i < n; ++i) {
        float xe = (x.cols() == 1) ? x(i, 0) : x(0, i);
        float ye = (y.cols() == 1) ? y(i, 0) : y(0, i);
        ret += xe * ye;
    }
    return ret;
}

// v: column vector view (len x 1), u: len x 1 output.  Returns false for the trivial reflection.
inline bool householder_vector(const MV& v, Mat& u, float& vnorm) {
    vnorm = (v(0, 0) > 0.0f) ? -norm_vec(v) : norm_vec(v);
    float f = std::sqrt(vnorm * (vnorm - v(0, 0)));
    if (f == 0.0f) {
        std::fill(u.a.begin(), u.a.end(), 0.0f);
        return false;
    }
    u(0, 0) = (v(0, 0) - vnorm) / f;
    for (long k = 1; k < u.r; ++k) u(k, 0) = v(k, 0) / f;
    return true;
}

// One Householder step on column i of r; optionally transforms rhs and stores u.
inline bool qr_householder_step(long i, MV r, MV rhs, MV householder) {
    const long m = r.rows(), n = r.cols(), rhs_count = rhs.cols();
    Mat u(m - i, 1);
    float vnorm;
    bool nontrivial = householder_vector(r.colv(i, i, m), u, vnorm);
    r(i, i) = vnorm;
    for (long k = i + 1; k < m; ++k) r(k, i) = 0.0f;
    if (householder.cols() == n)
        for (long k = i; k < m; ++k) householder(k, i) = u(k - i, 0);
    if (nontrivial) {
        MV uv = u.view();
        for (long k = i + 1; k < n; ++k) {
            float d = dot_vec(r.colv(i, k, m), uv);
            for (long l = i; l < m; ++l) r(l, k) -= d * u(l - i, 0);
        }
        for (long k = 0; k < rhs_count; ++k) {
            float d = dot_vec(rhs.colv(i, k, m), uv);
            for (long l = i; l < m; ++l) rhs(l, k) -= d * u(l - i, 0);
        }
    }
    return r(i, i) != 0.0f;
}

// First index of the strict maximum of vals[from..n) (Vigra argMax: starts at -FLT_MAX).
inline long arg_max_from(const std::vector<float>& vals, long from, long n) {
    long best = -1;
    float cur = -FLT_MAX;
    for (long l = from; l < n; ++l)
        if (vals[(size_t)l] > cur) {
            cur = vals[(size_t)l];
            best = l - from;
        }
    return best;
}

inline void swap_columns(MV r, long a, long b) {
    for (long i = 0; i < r.rows(); ++i) std::swap(r(i, a), r(i, b));
}

// Householder QR to upper-triangular form; returns the numerical rank.  n < 4 only (the "simple
// singular value approximation" branch); the reference never calls it with n >= 4.
inline unsigned qr_to_triangular(MV r, MV rhs, MV householder, std::vector<long>& permutation,
                                 double epsilon = 0.0) {
    const long m = r.rows(), n = r.cols();
    const long max_rank = std::min(m, n);
    if (n == 0) return 0;
    bool pivoting = !permutation.empty();

    std::vector<float> col_sq;
    if (pivoting) {
        col_sq.resize((size_t)n);
        for (long k = 0; k < n; ++k) col_sq[(size_t)k] = squared_norm_vec(r.colv(0, k, m));
        long pivot = arg_max_from(col_sq, 0, n);
        if (pivot != 0) {
            swap_columns(r, 0, pivot);
            std::swap(col_sq[0], col_sq[(size_t)pivot]);
            std::swap(permutation[0], permutation[(size_t)pivot]);
        }
    }

    qr_householder_step(0, r, rhs, householder);

    long rank = 1;
    float max_sv = std::fabs(r(0, 0)), min_sv = max_sv;
    double tolerance = (epsilon == 0.0) ? (double)((float)m * max_sv * FLT_EPSILON) : epsilon;
    if ((double)min_sv <= tolerance) {
        rank = 0;
        pivoting = false;
    }

    for (long k = 1; k < max_rank; ++k) {
        if (pivoting) {
            for (long l = k; l < n; ++l) {
                float e = r(k, l);
                col_sq[(size_t)l] -= e * e;
            }
            long pivot = k + arg_max_from(col_sq, k, n);
            if (pivot != k) {
                swap_columns(r, k, pivot);
                std::swap(col_sq[(size_t)k], col_sq[(size_t)pivot]);
                std::swap(permutation[(size_t)k], permutation[(size_t)pivot]);
            }
        }
        qr_householder_step(k, r, rhs, householder);

        float nv = std::fabs(r(k, k));
        max_sv = std::max(nv, max_sv);
        min_sv = std::min(nv, min_sv);
        if (epsilon == 0.0) tolerance = (double)((float)m * max_sv * FLT_EPSILON);
        if ((double)min_sv > tolerance)
            ++rank;
        else
            pivoting = false;
    }
    return (unsigned)rank;
}

inline bool solve_upper_triangular(const MV& r, const MV& b, MV x) {
    const long m = r.rows(), rhs_count = b.cols();
    for (long k = 0; k < rhs_count; ++k)
        for (long i = m - 1; i >= 0; --i) {
            if (r(i, i) == 0.0f) return false;
            float sum = b(i, k);
            for (long j = i + 1; j < m; ++j) sum -= r(i, j) * x(j, k);
            x(i, k) = sum / r(i, i);
        }
    return true;
}

inline bool solve_lower_triangular(const MV& l, const MV& b, MV x) {
    const long m = l.cols(), n = b.cols();
    for (long k = 0; k < n; ++k)
        for (long i = 0; i < m; ++i) {
            if (l(i, i) == 0.0f) return false;
            float sum = b(i, k);
            for (long j = 0; j < i; ++j) sum -= l(i, j) * x(j, k);
            x(i, k) = sum / l(i, i);
        }
    return true;
}

inline void apply_householder_column_reflections(const MV& householder, MV res) {
    const long n = householder.rows(), m = householder.cols(), rhs_count = res.cols();
    for (long k = m - 1; k >= 0; --k) {
        MV u = householder.colv(k, k, n);
        for (long l = 0; l < rhs_count; ++l) {
            float d = dot_vec(res.colv(k, l, n), u);
            for (long i = k; i < n; ++i) res(i, l) -= d * u(i - k, 0);
        }
    }
}

// linalg::inverse for a square matrix: QR without pivoting, false unless full rank, R*res = Q^T.
inline bool inverse(const MV& v, MV res) {
    const long n = v.cols();
    Mat r = copy_of(v), q(n, n);
    for (long i = 0; i < n; ++i) q(i, i) = 1.0f;
    std::vector<long> no_pivoting;
    MV tq = q.view().T();
    unsigned rank = qr_to_triangular(r.view(), tq, MV{}, no_pivoting, 0.0);
    if ((long)rank != n) return false;
    solve_upper_triangular(r.view(), tq, res);
    return true;
}

// linalg::linearSolve(A, b, res, "QR") for m >= n.  Returns rank == n; res is written either way.
inline bool linear_solve(const MV& A_in, const MV& b_in, MV res) {
    Mat A = copy_of(A_in), b = copy_of(b_in);
    const long n = A.c, m = A.r, rhs_count = res.cols();
    std::vector<long> permutation((size_t)n);
    for (long k = 0; k < n; ++k) permutation[(size_t)k] = k;

    long rank = (long)qr_to_triangular(A.view(), b.view(), MV{}, permutation, 0.0);

    Mat permuted(n, rhs_count);
    if (rank < n) {
        Mat hh(n, rank);
        MV asub = A.view().sub(0, 0, rank, n);
        std::vector<long> no_pivoting;
        qr_to_triangular(asub.T(), MV{}, hh.view(), no_pivoting, 0.0);
        solve_lower_triangular(asub.sub(0, 0, rank, rank), b.view().sub(0, 0, rank, rhs_count),
                               permuted.view().sub(0, 0, rank, rhs_count));
        apply_householder_column_reflections(hh.view(), permuted.view());
    } else {
        solve_upper_triangular(A.view().sub(0, 0, rank, rank), b.view().sub(0, 0, rank, rhs_count),
                               permuted.view());
    }
    for (long k = 0; k < n; ++k)
        for (long l = 0; l < rhs_count; ++l) res(permutation[(size_t)k], l) = permuted(k, l);
    (void)m;
    return rank == n;
}

}  // namespace la
}  // namespace oracle
