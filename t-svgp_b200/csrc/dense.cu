#include "dense.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include <stdlib.h>

namespace tsvgp {

#define TRY(x) do { int e_ = (x); if (e_) return e_; } while (0)

constexpr int NB = 128;
static int g_chol_outer = 512;   // outer panel width of the two-level blocking (multiple of 128)

// Right-looking with look-ahead over two streams.  Per block column q, on `s` (the critical path):
//   diag(q) -> panel(q) = A[q+1:, q] L_qq^-T -> [wait: trailing(q-1) done] -> A[q+1:, q+1] -= panel(q) panel(q)[0]^T  -> diag(q+1) ...
// and on aux->s2, once panel(q) exists:  A[q+2:, q+2:] -= panel(q)[1:] panel(q)[1:]^T  (lower tiles), hidden under diag(q+1).
// The two streams always write disjoint block columns (q+1 on `s`, >= q+2 on s2); `s` joins s2 before it returns.
static int chol_lower_lookahead(double* A, long ld, int n, double* dinv, int* info, cudaStream_t s, double* ws, size_t ws_doubles,
                                const CholAux& aux) {
#define CUQ(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
    bool f_pending = false;
    for (int q = 0; q < n; q += NB) {
        double* Aqq = A + (long)q * ld + q;
        double* Dq = dinv + (long)(q / NB) * NB * NB;
        TRY(diag_potrf_inv_launch(Aqq, ld, Dq, q / NB, info, s));
        const int below = n - (q + NB);
        if (below <= 0) break;
        double* panel = A + (long)(q + NB) * ld + q;
        {   // panel <- panel * L_qq^-T
            GemmP p;
            p.A = panel; p.lda = ld; p.a_kc = 1;
            p.B = Dq; p.ldb = NB; p.b_kc = 1;
            p.C = panel; p.ldc = ld;
            p.m = below; p.n = NB; p.k = NB;
            TRY(gemm_launch_auto(p, s, ws, ws_doubles));
        }
        const int rest = below - NB;   // rows / columns from q + 2 blocks on
        if (rest > 0) {   // trailing update of the columns >= q+2 on the helper stream
            CUQ(cudaEventRecord(aux.e, s));
            CUQ(cudaStreamWaitEvent(aux.s2, aux.e, 0));
            GemmP p;
            const double* Lp = panel + (long)NB * ld;
            p.A = Lp; p.lda = ld; p.a_kc = 1;
            p.B = Lp; p.ldb = ld; p.b_kc = 1;
            p.C = A + (long)(q + 2 * NB) * ld + (q + 2 * NB); p.ldc = ld;
            p.m = rest; p.n = rest; p.k = NB;
            p.alpha = -1.0; p.beta = 1.0; p.lower_out = 1;
            // trailing(q-1) precedes it in s2's own order; block column q+1 is written on `s` meanwhile (disjoint)
            TRY(gemm_launch_auto(p, aux.s2, nullptr, 0));
        }
        // block column q+1 on the critical stream; trailing(q-1), which also wrote this column, must be complete
        if (f_pending) CUQ(cudaStreamWaitEvent(s, aux.f, 0));
        {
            GemmP p;
            p.A = panel; p.lda = ld; p.a_kc = 1;
            p.B = panel; p.ldb = ld; p.b_kc = 1;
            p.C = A + (long)(q + NB) * ld + (q + NB); p.ldc = ld;
            p.m = below; p.n = NB; p.k = NB;
            p.alpha = -1.0; p.beta = 1.0;
            TRY(gemm_launch_auto(p, s, ws, ws_doubles));
        }
        if (rest > 0) { CUQ(cudaEventRecord(aux.f, aux.s2)); f_pending = true; }
        else f_pending = false;
    }
    if (f_pending) CUQ(cudaStreamWaitEvent(s, aux.f, 0));
#undef CUQ
    return zero_upper_launch(A, ld, n, s);
}

int chol_lower(double* A, long ld, int n, double* dinv, int* info, cudaStream_t s, double* ws, size_t ws_doubles, const CholAux* aux) {
    static const bool la_off = getenv("TSVGP_CHOL_LOOKAHEAD") && atoi(getenv("TSVGP_CHOL_LOOKAHEAD")) == 0;
    // measured (tools/diag_bench, profiles/diag_r02.txt): 1.12 -> 1.04 ms at n = 2048, 2.98 -> 2.90 ms at n = 4096; at n = 8192 the
    // single-level trailing updates (k = 128 per pass over the trailing matrix) lose to the two-level blocking below (12.4 vs 10.8 ms)
    if (aux && aux->s2 && n >= 4 * NB && n <= 4096 && !la_off) return chol_lower_lookahead(A, ld, n, dinv, info, s, ws, ws_doubles, *aux);
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
