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

// ---- register-resident 3x3 specialisations for the per-candidate hot path (sift.cpp:306, :311) ----------
// Same operation order as the generic routines above (every loop bound is a compile-time constant, so the
// arrays live in registers); only the full-rank outcome is produced, which is all _eliminateEdgeResponses uses:
// a rank-deficient inverse or solve rejects the candidate and its minimum-norm solution is never read.

// One Householder step on column I of r (3x3) applied to NB right-hand-side columns of t.
template <int I, int NB>
__device__ __forceinline__ void hh_step3(float (&r)[3][3], float (&t)[3][NB]) {
    float ss = 0.0f;
#pragma unroll
    for (int k = I; k < 3; ++k) ss = ss + r[k][I] * r[k][I];
    const float nv = sqrtf(ss);
    const float v0 = r[I][I];
    const float vnorm = (v0 > 0.0f) ? -nv : nv;
    const float f = sqrtf(vnorm * (vnorm - v0));
    float u[3 - I];
    const bool nontrivial = !(f == 0.0f);
    if (nontrivial) {
        u[0] = (v0 - vnorm) / f;
#pragma unroll
        for (int k = 1; k < 3 - I; ++k) u[k] = r[I + k][I] / f;
    } else {
#pragma unroll
        for (int k = 0; k < 3 - I; ++k) u[k] = 0.0f;
    }
    r[I][I] = vnorm;
#pragma unroll
    for (int k = I + 1; k < 3; ++k) r[k][I] = 0.0f;
    if (nontrivial) {
#pragma unroll
        for (int k = I + 1; k < 3; ++k) {
            float d = 0.0f;
#pragma unroll
            for (int l = I; l < 3; ++l) d = d + r[l][k] * u[l - I];
#pragma unroll
            for (int l = I; l < 3; ++l) r[l][k] = r[l][k] - d * u[l - I];
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            float d = 0.0f;
#pragma unroll
            for (int l = I; l < 3; ++l) d = d + t[l][k] * u[l - I];
#pragma unroll
            for (int l = I; l < 3; ++l) t[l][k] = t[l][k] - d * u[l - I];
        }
    }
}

// Vigra's rank estimate for n < 4 from the diagonal of R as it is produced.
struct RankTracker {
    float max_sv, min_sv;
    int rank;
    bool full;  // stays true while every step increments the rank
    __device__ __forceinline__ void first(float r00) {
        max_sv = fabsf(r00);
        min_sv = max_sv;
        const double tol = (double)(3.0f * max_sv * FLT_EPSILON);
        rank = ((double)min_sv <= tol) ? 0 : 1;
        full = rank == 1;
    }
    __device__ __forceinline__ void next(float rkk) {
        const float nv = fabsf(rkk);
        max_sv = fmaxf(nv, max_sv);
        min_sv = fminf(nv, min_sv);
        const double tol = (double)(3.0f * max_sv * FLT_EPSILON);
        if ((double)min_sv > tol) ++rank; else full = false;
    }
};

// linalg::inverse for 3x3 (row-major a): false unless full rank; out = R^-1 Q^T.
__device__ __forceinline__ bool inverse3_fast(const float* a, float* out) {
    float r[3][3], t[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { r[i][j] = a[i * 3 + j]; t[i][j] = i == j ? 1.0f : 0.0f; }
    RankTracker rk;
    hh_step3<0, 3>(r, t);
    rk.first(r[0][0]);
    hh_step3<1, 3>(r, t);
    rk.next(r[1][1]);
    hh_step3<2, 3>(r, t);
    rk.next(r[2][2]);
    if (rk.rank != 3) return false;
    // linearSolveUpperTriangular(R, Q^T, out)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float x[3];
#pragma unroll
        for (int i = 2; i >= 0; --i) {
            if (r[i][i] == 0.0f) return true;  // solve_upper gives up; inverse() ignores its return value
            float sum = t[i][k];
#pragma unroll
            for (int j = i + 1; j < 3; ++j) sum = sum - r[i][j] * x[j];
            x[i] = sum / r[i][i];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) out[i * 3 + k] = x[i];
    }
    return true;
}

__device__ __forceinline__ void swap_cols3(float (&r)[3][3], float (&csq)[3], int (&perm)[3], int a, int b) {
    // a, b are compile-time at every call site below
#pragma unroll
    for (int l = 0; l < 3; ++l) { const float tmp = r[l][a]; r[l][a] = r[l][b]; r[l][b] = tmp; }
    const float ts = csq[a]; csq[a] = csq[b]; csq[b] = ts;
    const int ti = perm[a]; perm[a] = perm[b]; perm[b] = ti;
}

// linalg::linearSolve(A, b, res, "QR") for 3x3 A and 3x1 b.  Returns rank == 3; res is only written then.
__device__ __forceinline__ bool solve3_fullrank(const float* a_in, const float* b_in, float* res) {
    float r[3][3], t[3][1];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        t[i][0] = b_in[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) r[i][j] = a_in[i * 3 + j];
    }
    int perm[3] = {0, 1, 2};
    float csq[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float sq = 0.0f;
#pragma unroll
        for (int l = 0; l < 3; ++l) sq = sq + r[l][k] * r[l][k];
        csq[k] = sq;
    }
    {   // first strict maximum, starting from -FLT_MAX
        int pivot = -1;
        float cur = -FLT_MAX;
#pragma unroll
        for (int l = 0; l < 3; ++l)
            if (csq[l] > cur) { cur = csq[l]; pivot = l; }
        if (pivot == 1) swap_cols3(r, csq, perm, 0, 1);
        else if (pivot == 2) swap_cols3(r, csq, perm, 0, 2);
    }
    RankTracker rk;
    hh_step3<0, 1>(r, t);
    rk.first(r[0][0]);
    if (rk.rank == 1) {  // pivoting continues only while the matrix still looks full rank
#pragma unroll
        for (int l = 1; l < 3; ++l) csq[l] = csq[l] - r[1][l] * r[1][l];
        int best = -1;
        float cur = -FLT_MAX;
#pragma unroll
        for (int l = 1; l < 3; ++l)
            if (csq[l] > cur) { cur = csq[l]; best = l; }
        if (best == 2) swap_cols3(r, csq, perm, 1, 2);
    }
    hh_step3<1, 1>(r, t);
    rk.next(r[1][1]);
    // k = 2: the downdate and argmax run over the single remaining column; no swap is possible
    hh_step3<2, 1>(r, t);
    rk.next(r[2][2]);
    if (rk.rank != 3) return false;
    float x[3];
#pragma unroll
    for (int i = 2; i >= 0; --i) {
        if (r[i][i] == 0.0f) break;  // cannot happen at full rank; mirrors linearSolveUpperTriangular's early return
        float sum = t[i][0];
#pragma unroll
        for (int j = i + 1; j < 3; ++j) sum = sum - r[i][j] * x[j];
        x[i] = sum / r[i][i];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (perm[k] == 0) res[0] = x[k];
        else if (perm[k] == 1) res[1] = x[k];
        else res[2] = x[k];
    }
    return true;
}

// Minimum-norm least-squares tail of linearSolveQRReplace for a rank-deficient 3x3 system (rank RANK < 3), in
// registers: QR of (R[0:RANK, 0:3])^T without pivoting, Householder vectors kept; lower-triangular solve;
// Householder reflections applied back; inverse permutation.  r, t, perm come from the pivoted first stage.
template <int RANK>
__device__ __forceinline__ void min_norm_tail3(const float (&r)[3][3], const float (&t)[3][1], const int (&perm)[3], float* res) {
    float ps[3] = {0.0f, 0.0f, 0.0f};
    if (RANK > 0) {
        constexpr int NR = RANK > 0 ? RANK : 1;
        float at[3][NR], hh[3][NR];  // at = transpose of the first RANK rows of R
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < NR; ++j) { at[i][j] = r[j][i]; hh[i][j] = 0.0f; }
#pragma unroll
        for (int I = 0; I < NR; ++I) {
            float ss = 0.0f;
#pragma unroll
            for (int k = I; k < 3; ++k) ss = ss + at[k][I] * at[k][I];
            const float nv = sqrtf(ss);
            const float v0 = at[I][I];
            const float vnorm = (v0 > 0.0f) ? -nv : nv;
            const float f = sqrtf(vnorm * (vnorm - v0));
            float u[3];
            const bool nontrivial = !(f == 0.0f);
#pragma unroll
            for (int k = 0; k < 3; ++k) u[k] = 0.0f;
            if (nontrivial) {
                u[0] = (v0 - vnorm) / f;
#pragma unroll
                for (int k = 1; k < 3 - I; ++k) u[k] = at[I + k][I] / f;
            }
            at[I][I] = vnorm;
#pragma unroll
            for (int k = I + 1; k < 3; ++k) at[k][I] = 0.0f;
#pragma unroll
            for (int k = I; k < 3; ++k) hh[k][I] = u[k - I];
            if (nontrivial) {
#pragma unroll
                for (int k = I + 1; k < NR; ++k) {
                    float d = 0.0f;
#pragma unroll
                    for (int l = I; l < 3; ++l) d = d + at[l][k] * u[l - I];
#pragma unroll
                    for (int l = I; l < 3; ++l) at[l][k] = at[l][k] - d * u[l - I];
                }
            }
        }
        // linearSolveLowerTriangular(Asub[0:RANK, 0:RANK], b[0:RANK]): l(i, j) = at[j][i]
        bool alive = true;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (alive && at[i][i] == 0.0f) alive = false;
            if (alive) {
                float sum = t[i][0];
#pragma unroll
                for (int j = 0; j < i; ++j) sum = sum - at[j][i] * ps[j];
                ps[i] = sum / at[i][i];
            }
        }
        // applyHouseholderColumnReflections
#pragma unroll
        for (int k = NR - 1; k >= 0; --k) {
            float d = 0.0f;
#pragma unroll
            for (int i = k; i < 3; ++i) d = d + ps[i] * hh[i][k];
#pragma unroll
            for (int i = k; i < 3; ++i) ps[i] = ps[i] - d * hh[i][k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (perm[k] == 0) res[0] = ps[k];
        else if (perm[k] == 1) res[1] = ps[k];
        else res[2] = ps[k];
    }
}

// linalg::linearSolve(A, b, res, "QR") for 3x3 A and 3x1 b, any rank, in registers.  Returns rank == 3; res holds the
// solution (full rank) or the minimum-norm least-squares solution (rank deficient), like Vigra.
__device__ __forceinline__ bool solve3_fast(const float* a_in, const float* b_in, float* res) {
    float r[3][3], t[3][1];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        t[i][0] = b_in[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) r[i][j] = a_in[i * 3 + j];
    }
    int perm[3] = {0, 1, 2};
    float csq[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float sq = 0.0f;
#pragma unroll
        for (int l = 0; l < 3; ++l) sq = sq + r[l][k] * r[l][k];
        csq[k] = sq;
    }
    {
        int pivot = -1;
        float cur = -FLT_MAX;
#pragma unroll
        for (int l = 0; l < 3; ++l)
            if (csq[l] > cur) { cur = csq[l]; pivot = l; }
        if (pivot == 1) swap_cols3(r, csq, perm, 0, 1);
        else if (pivot == 2) swap_cols3(r, csq, perm, 0, 2);
    }
    RankTracker rk;
    hh_step3<0, 1>(r, t);
    rk.first(r[0][0]);
    if (rk.rank == 1) {
#pragma unroll
        for (int l = 1; l < 3; ++l) csq[l] = csq[l] - r[1][l] * r[1][l];
        int best = -1;
        float cur = -FLT_MAX;
#pragma unroll
        for (int l = 1; l < 3; ++l)
            if (csq[l] > cur) { cur = csq[l]; best = l; }
        if (best == 2) swap_cols3(r, csq, perm, 1, 2);
    }
    hh_step3<1, 1>(r, t);
    rk.next(r[1][1]);
    hh_step3<2, 1>(r, t);
    rk.next(r[2][2]);
    if (rk.rank == 3) {
        float x[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 2; i >= 0; --i) {
            if (r[i][i] == 0.0f) break;
            float sum = t[i][0];
#pragma unroll
            for (int j = i + 1; j < 3; ++j) sum = sum - r[i][j] * x[j];
            x[i] = sum / r[i][i];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (perm[k] == 0) res[0] = x[k];
            else if (perm[k] == 1) res[1] = x[k];
            else res[2] = x[k];
        }
        return true;
    }
    if (rk.rank == 2) min_norm_tail3<2>(r, t, perm, res);
    else if (rk.rank == 1) min_norm_tail3<1>(r, t, perm, res);
    else min_norm_tail3<0>(r, t, perm, res);
    return false;
}

}  // namespace qr
}  // namespace siftgpu
