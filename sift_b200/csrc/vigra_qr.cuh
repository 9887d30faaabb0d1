// Device fp32 Householder QR with Vigra 1.11's rank rule and minimum-norm least squares, as used
// by the reference at sift.cpp:306 (inverse), sift.cpp:311 and algorithms.cpp:175 (linearSolve,
// method "QR").  Vigra is an un-vendored dependency of the reference; the operation order below
// follows its published algorithm (vigra/linear_solve.hxx) as confirmed in the reference's
// shipped binary.  Every multiply and add is a separate IEEE fp32
// operation (this translation unit is compiled with -fmad=false) so results equal the x86
// mulss/addss sequence bit for bit.  At most 3x3.
#pragma once
#include <float.h>

namespace siftgpu {
namespace qr {

struct View {  // (i, j) = row i, column j over caller storage
    float* p;
    int s0, s1, n0, n1;
    __device__ float& operator()(int i, int j) const { return p[i * s0 + j * s1]; }
    __device__ View T() const { return View{p, s1, s0, n1, n0}; }
    __device__ View sub(int i0, int j0, int i1, int j1) const { return View{p + i0 * s0 + j0 * s1, s0, s1, i1 - i0, j1 - j0}; }
};

__device__ inline View make_view(float* p, int rows, int cols) { return View{p, cols, 1, rows, cols}; }  // row-major storage
__device__ inline View empty_view() { return View{nullptr, 0, 0, 0, 0}; }

// One Householder step on column i of r (rows i..m-1); optionally transforms rhs and stores u.
__device__ inline void householder_step(int i, View r, View rhs, View hh) {
    const int m = r.n0, n = r.n1;
    float u[3];
    // vnorm = v0 > 0 ? -|v| : |v|, |v| = sqrtf(sequential sum of squares)
    float ss = 0.0f;
    for (int k = i; k < m; ++k) {
        float e = r(k, i);
        ss = ss + e * e;
    }
    const float nv = sqrtf(ss);
    const float v0 = r(i, i);
    const float vnorm = (v0 > 0.0f) ? -nv : nv;
    const float f = sqrtf(vnorm * (vnorm - v0));
    bool nontrivial;
    if (f == 0.0f) {
        for (int k = 0; k < m - i; ++k) u[k] = 0.0f;
        nontrivial = false;
    } else {
        u[0] = (v0 - vnorm) / f;
        for (int k = 1; k < m - i; ++k) u[k] = r(i + k, i) / f;
        nontrivial = true;
    }
    r(i, i) = vnorm;
    for (int k = i + 1; k < m; ++k) r(k, i) = 0.0f;
    if (hh.n1 == n)
        for (int k = i; k < m; ++k) hh(k, i) = u[k - i];
    if (nontrivial) {
        for (int k = i + 1; k < n; ++k) {
            float d = 0.0f;
            for (int l = i; l < m; ++l) d = d + r(l, k) * u[l - i];
            for (int l = i; l < m; ++l) r(l, k) = r(l, k) - d * u[l - i];
        }
        for (int k = 0; k < rhs.n1; ++k) {
            float d = 0.0f;
            for (int l = i; l < m; ++l) d = d + rhs(l, k) * u[l - i];
            for (int l = i; l < m; ++l) rhs(l, k) = rhs(l, k) - d * u[l - i];
        }
    }
}

// Returns the numerical rank.  perm == nullptr: no pivoting.
__device__ inline int to_triangular(View r, View rhs, View hh, int* perm) {
    const int m = r.n0, n = r.n1;
    const int max_rank = m < n ? m : n;
    if (n == 0) return 0;
    bool pivoting = perm != nullptr;
    float col_sq[3];
    if (pivoting) {
        for (int k = 0; k < n; ++k) {
            float s = 0.0f;
            for (int l = 0; l < m; ++l) {
                float e = r(l, k);
                s = s + e * e;
            }
            col_sq[k] = s;
        }
        int pivot = -1;
        float cur = -FLT_MAX;
        for (int l = 0; l < n; ++l)
            if (col_sq[l] > cur) { cur = col_sq[l]; pivot = l; }
        if (pivot != 0 && pivot >= 0) {
            for (int l = 0; l < m; ++l) { float t = r(l, 0); r(l, 0) = r(l, pivot); r(l, pivot) = t; }
            float t = col_sq[0]; col_sq[0] = col_sq[pivot]; col_sq[pivot] = t;
            int ti = perm[0]; perm[0] = perm[pivot]; perm[pivot] = ti;
        }
    }
    householder_step(0, r, rhs, hh);
    int rank = 1;
    float max_sv = fabsf(r(0, 0)), min_sv = max_sv;
    double tol = (double)((float)m * max_sv * FLT_EPSILON);
    if ((double)min_sv <= tol) {
        rank = 0;
        pivoting = false;
    }
    for (int k = 1; k < max_rank; ++k) {
        if (pivoting) {
            for (int l = k; l < n; ++l) {
                float e = r(k, l);
                col_sq[l] = col_sq[l] - e * e;
            }
            int best = -1;
            float cur = -FLT_MAX;
            for (int l = k; l < n; ++l)
                if (col_sq[l] > cur) { cur = col_sq[l]; best = l; }
            if (best != k && best >= 0) {
                for (int l = 0; l < m; ++l) { float t = r(l, k); r(l, k) = r(l, best); r(l, best) = t; }
                float t = col_sq[k]; col_sq[k] = col_sq[best]; col_sq[best] = t;
                int ti = perm[k]; perm[k] = perm[best]; perm[best] = ti;
            }
        }
        householder_step(k, r, rhs, hh);
        const float nv = fabsf(r(k, k));
        max_sv = fmaxf(nv, max_sv);
        min_sv = fminf(nv, min_sv);
        tol = (double)((float)m * max_sv * FLT_EPSILON);
        if ((double)min_sv > tol)
            ++rank;
        else
            pivoting = false;
    }
    return rank;
}

__device__ inline bool solve_upper(View r, View b, View x) {
    const int m = r.n0;
    for (int k = 0; k < b.n1; ++k)
        for (int i = m - 1; i >= 0; --i) {
            if (r(i, i) == 0.0f) return false;
            float sum = b(i, k);
            for (int j = i + 1; j < m; ++j) sum = sum - r(i, j) * x(j, k);
            x(i, k) = sum / r(i, i);
        }
    return true;
}

__device__ inline bool solve_lower(View l, View b, View x) {
    const int m = l.n1;
    for (int k = 0; k < b.n1; ++k)
        for (int i = 0; i < m; ++i) {
            if (l(i, i) == 0.0f) return false;
            float sum = b(i, k);
            for (int j = 0; j < i; ++j) sum = sum - l(i, j) * x(j, k);
            x(i, k) = sum / l(i, i);
        }
    return true;
}

// linalg::inverse(a) for 3x3, row-major in/out.  False unless full rank.
__device__ inline bool inverse3(const float* a, float* out) {
    float r[9], q[9];
    for (int i = 0; i < 9; ++i) { r[i] = a[i]; q[i] = 0.0f; }
    q[0] = q[4] = q[8] = 1.0f;
    View R = make_view(r, 3, 3), TQ = make_view(q, 3, 3).T();
    if (to_triangular(R, TQ, empty_view(), nullptr) != 3) return false;
    solve_upper(R, TQ, make_view(out, 3, 3));
    return true;
}

// linalg::linearSolve(A, b, res, "QR") for 3x3 A, 3x1 b.  Returns rank == 3; res written either way
// (minimum-norm least squares when rank deficient).
__device__ inline bool solve3(const float* a_in, const float* b_in, float* res) {
    float a[9], b[3];
    for (int i = 0; i < 9; ++i) a[i] = a_in[i];
    for (int i = 0; i < 3; ++i) b[i] = b_in[i];
    int perm[3] = {0, 1, 2};
    View A = make_view(a, 3, 3), B = make_view(b, 3, 1);
    const int rank = to_triangular(A, B, empty_view(), perm);
    float ps[3] = {0.0f, 0.0f, 0.0f};
    View PS = make_view(ps, 3, 1);
    if (rank < 3) {
        float hh[9];
        for (int i = 0; i < 9; ++i) hh[i] = 0.0f;
        View H = View{hh, 3, 1, 3, rank};  // n x rank inside 3x3 row-major storage
        View Asub = A.sub(0, 0, rank, 3);
        to_triangular(Asub.T(), empty_view(), H, nullptr);
        solve_lower(Asub.sub(0, 0, rank, rank), B.sub(0, 0, rank, 1), PS.sub(0, 0, rank, 1));
        // applyHouseholderColumnReflections
        for (int k = rank - 1; k >= 0; --k) {
            float d = 0.0f;
            for (int i = k; i < 3; ++i) d = d + ps[i] * H(i, k);
            for (int i = k; i < 3; ++i) ps[i] = ps[i] - d * H(i, k);
        }
    } else {
        solve_upper(A, B, PS);
    }
    for (int k = 0; k < 3; ++k) res[perm[k]] = ps[k];
    return rank == 3;
}

}  // namespace qr
}  // namespace siftgpu
