#include "dense.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace tsvgp {

#define TRY(x) do { int e_ = (x); if (e_) return e_; } while (0)

constexpr int NB = 128;
static int g_chol_outer = 512;   // outer panel width of the two-level blocking (multiple of 128)

int chol_lower(double* A, long ld, int n, double* dinv, int* info, cudaStream_t s, double* ws, size_t ws_doubles) {
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
                TRY(gemm_launch_auto(p, s, ws, ws_doubles));
            }
            const int ncols = Pend - (q + NB);
            if (ncols > 0) {   // update the rest of the outer panel: A[q+NB:, q+NB:Pend] -= panel * panel[0:ncols]^T
                GemmP p;
                p.A = panel; p.lda = ld; p.a_kc = 1;
                p.B = panel; p.ldb = ld; p.b_kc = 1;
                p.C = A + (long)(q + NB) * ld + (q + NB); p.ldc = ld;
                p.m = below; p.n = ncols; p.k = NB;
                p.alpha = -1.0; p.beta = 1.0; p.lower_out = 1;
                TRY(gemm_launch_auto(p, s, ws, ws_doubles));
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
            TRY(gemm_launch_auto(p, s, ws, ws_doubles));
        }
    }
    return zero_upper_launch(A, ld, n, s);
}

// Bottom-up triangular inverse.  At level h (blocks) every node [lo, lo + 2h) combines its finished halves:
//   Linv[right, left] = -Linv[right, right] * L[right, left] * Linv[left, left]
// All full nodes of a level have the same shape and a constant stride, so they go out as ONE batched launch per product
// (2 launches per level instead of 2 per node); a ragged last node (n not a power of two) is launched on its own.
static int trtri_level(const double* L, long ld, int lo, int hl, int hr, int batch, long stride, double* Linv, double* tmp,
                       long tmp_stride, cudaStream_t s, double* ws, size_t ws_doubles) {
    const int mid = lo + hl;
    const int mrows = hr * NB, ncols = hl * NB;
    {   // tmp = C * Ainv,  C = L[right, left],  Ainv = Linv[left, left] (lower: B(k,j) != 0 only for k >= j)
        GemmP p;
        p.A = L + (long)mid * NB * ld + (long)lo * NB; p.lda = ld; p.a_kc = 1;
        p.B = Linv + (long)lo * NB * ld + (long)lo * NB; p.ldb = ld; p.b_kc = 0; p.b_tri = 2;
        p.C = tmp; p.ldc = ld;
        p.m = mrows; p.n = ncols; p.k = ncols;
        p.batch = batch; p.sA = stride; p.sB = stride; p.sC = tmp_stride;
        TRY(gemm_launch_auto(p, s, ws, ws_doubles));
    }
    {   // Linv[right, left] = -Binv * tmp,  Binv = Linv[right, right] (lower: A(i,k) != 0 only for k <= i)
        GemmP p;
        p.A = Linv + (long)mid * NB * ld + (long)mid * NB; p.lda = ld; p.a_kc = 1; p.a_tri = 1;
        p.B = tmp; p.ldb = ld; p.b_kc = 0;
        p.C = Linv + (long)mid * NB * ld + (long)lo * NB; p.ldc = ld;
        p.m = mrows; p.n = ncols; p.k = mrows;
        p.alpha = -1.0;
        p.batch = batch; p.sA = stride; p.sB = tmp_stride; p.sC = stride;
        TRY(gemm_launch_auto(p, s, ws, ws_doubles));
    }
    return 0;
}

int trtri_lower(const double* L, long ld, int n, const double* dinv, double* Linv, double* tmp, cudaStream_t s, double* ws, size_t ws_doubles) {
    const int nblk = n / NB;
    TRY(trtri_seed_launch(dinv, Linv, ld, n, s));
    for (int h = 1; h < nblk; h *= 2) {
        const int full = nblk / (2 * h);                 // nodes with both halves of size h
        const long stride = (long)2 * h * NB * (ld + 1);
        const long tmp_stride = (long)h * NB * ld;       // node b's scratch: rows [b*h*NB, (b+1)*h*NB) of tmp
        if (full > 0) TRY(trtri_level(L, ld, 0, h, h, full, stride, Linv, tmp, tmp_stride, s, ws, ws_doubles));
        const int lo = full * 2 * h, rem = nblk - lo;    // ragged node: left half complete, right half shorter
        if (rem > h) TRY(trtri_level(L, ld, lo, h, rem - h, 1, 0, Linv, tmp, 0, s, ws, ws_doubles));
    }
    return 0;
}

}  // namespace tsvgp
