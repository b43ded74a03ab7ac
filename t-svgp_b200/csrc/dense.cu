#include "dense.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace tsvgp {

#define TRY(x) do { int e_ = (x); if (e_) return e_; } while (0)

constexpr int NB = 128;
static int g_chol_outer = 512;   // outer panel width of the two-level blocking (multiple of 128)

int chol_lower(double* A, long ld, int n, double* dinv, int* info, cudaStream_t s) {
    const int OB = g_chol_outer;
    for (int P0 = 0; P0 < n; P0 += OB) {
        const int Pend = P0 + OB < n ? P0 + OB : n;
        for (int q = P0; q < Pend; q += NB) {
            double* Aqq = A + (long)q * ld + q;
            double* Dq = dinv + (long)(q / NB) * NB * NB;
            TRY(diag_potrf_inv_launch(Aqq, ld, Dq, q / NB, info, s));
            const int below = n - (q + NB);
            if (below <= 0) continue;
            double* panel = A + (long)(q + NB) * ld + q;
            {   // panel <- panel * L_qq^-T   (in place; B(k,j) = Linv[j][k])
                GemmP p;
                p.A = panel; p.lda = ld; p.a_kc = 1;
                p.B = Dq; p.ldb = NB; p.b_kc = 1;
                p.C = panel; p.ldc = ld;
                p.m = below; p.n = NB; p.k = NB;
                TRY(gemm_launch(p, s));
            }
            const int ncols = Pend - (q + NB);
            if (ncols > 0) {   // update the rest of the outer panel: A[q+NB:, q+NB:Pend] -= panel * panel[0:ncols]^T
                GemmP p;
                p.A = panel; p.lda = ld; p.a_kc = 1;
                p.B = panel; p.ldb = ld; p.b_kc = 1;
                p.C = A + (long)(q + NB) * ld + (q + NB); p.ldc = ld;
                p.m = below; p.n = ncols; p.k = NB;
                p.alpha = -1.0; p.beta = 1.0; p.lower_out = 1;
                TRY(gemm_launch(p, s));
            }
        }
        if (Pend < n) {   // trailing matrix -= L[Pend:, P0:Pend] L[Pend:, P0:Pend]^T
            GemmP p;
            const double* Lp = A + (long)Pend * ld + P0;
            p.A = Lp; p.lda = ld; p.a_kc = 1;
            p.B = Lp; p.ldb = ld; p.b_kc = 1;
            p.C = A + (long)Pend * ld + Pend; p.ldc = ld;
            p.m = n - Pend; p.n = n - Pend; p.k = Pend - P0;
            p.alpha = -1.0; p.beta = 1.0; p.lower_out = 1;
            TRY(gemm_launch(p, s));
        }
    }
    return zero_upper_launch(A, ld, n, s);
}

static int trtri_rec(const double* L, long ld, int lo, int hi, double* Linv, double* tmp, cudaStream_t s) {
    if (hi - lo <= 1) return 0;
    const int mid = (lo + hi) / 2;
    TRY(trtri_rec(L, ld, lo, mid, Linv, tmp, s));
    TRY(trtri_rec(L, ld, mid, hi, Linv, tmp, s));
    const int mrows = (hi - mid) * NB, ncols = (mid - lo) * NB;
    {   // tmp = C * Ainv,  C = L[mid:hi, lo:mid],  Ainv = Linv[lo:mid, lo:mid] (lower: B(k,j) != 0 only for k >= j)
        GemmP p;
        p.A = L + (long)mid * NB * ld + (long)lo * NB; p.lda = ld; p.a_kc = 1;
        p.B = Linv + (long)lo * NB * ld + (long)lo * NB; p.ldb = ld; p.b_kc = 0; p.b_tri = 2;
        p.C = tmp; p.ldc = ld;
        p.m = mrows; p.n = ncols; p.k = ncols;
        TRY(gemm_launch(p, s));
    }
    {   // Linv[mid:hi, lo:mid] = -Binv * tmp,  Binv = Linv[mid:hi, mid:hi] (lower: A(i,k) != 0 only for k <= i)
        GemmP p;
        p.A = Linv + (long)mid * NB * ld + (long)mid * NB; p.lda = ld; p.a_kc = 1; p.a_tri = 1;
        p.B = tmp; p.ldb = ld; p.b_kc = 0;
        p.C = Linv + (long)mid * NB * ld + (long)lo * NB; p.ldc = ld;
        p.m = mrows; p.n = ncols; p.k = mrows;
        p.alpha = -1.0;
        TRY(gemm_launch(p, s));
    }
    return 0;
}

int trtri_lower(const double* L, long ld, int n, const double* dinv, double* Linv, double* tmp, cudaStream_t s) {
    const int nblk = n / NB;
    if (cudaMemsetAsync(Linv, 0, sizeof(double) * (size_t)n * ld, s) != cudaSuccess) return (int)cudaGetLastError();
    for (int b = 0; b < nblk; ++b)
        TRY(place_block_launch(dinv + (long)b * NB * NB, NB, Linv + (long)b * NB * ld + (long)b * NB, ld, NB, NB, s));
    return trtri_rec(L, ld, 0, nblk, Linv, tmp, s);
}

}  // namespace tsvgp
