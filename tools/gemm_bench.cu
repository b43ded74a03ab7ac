// Tooling: times the DMMA GEMM engine on the two streaming-pass shapes (variance product, weighted SYRK) and on dense
// M x M shapes, and checks each against a naive FP64 kernel.  Build: make -C tools gemm_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../t-svgp_b200/csrc/gemm.cuh"
#include "../t-svgp_b200/csrc/common.cuh"
namespace tsvgp { thread_local long g_launches = 0; int g_debug_sync = 0; int g_pdl = 1; thread_local int g_pdl_suspended = 0; }
using namespace tsvgp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void fill(double* p, long n, unsigned seed, int tri_ld, int tri) {   // tri: 1 keep lower (col<=row), 2 keep upper
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned x = seed ^ (unsigned)(i * 2654435761u); x ^= x >> 13; x *= 0x5bd1e995; x ^= x >> 15;
    double v = ((x & 0xffff) / 65536.0) - 0.5;
    if (tri) { long r = i / tri_ld, c = i % tri_ld; if ((tri == 1 && c > r) || (tri == 2 && c < r)) v = 0.0; }
    p[i] = v;
}
// naive reference: C = alpha * sum_k A(i,k) s(k) B(k,j) + beta*C0 for a sampled set of entries
__global__ void ref_entries(GemmP p, const double* C0, const int* ii, const int* jj, int ne, double* out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int i = ii[e], j = jj[e];
    double s = 0;
    for (int k = 0; k < p.k; ++k) {
        double a = p.a_kc ? p.A[(long)i * p.lda + k] : p.A[(long)k * p.lda + i];
        double b = p.b_kc ? p.B[(long)j * p.ldb + k] : p.B[(long)k * p.ldb + j];
        if (p.kscale) b *= p.kscale[k];
        s = fma(a, b, s);
    }
    out[e] = p.alpha * s + (p.beta != 0.0 ? p.beta * C0[(long)i * p.ldc + j] : 0.0);
}

static double time_launch(const GemmP& p, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) gemm_launch(p, 0);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) gemm_launch(p, 0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char** argv) {
    int M = argc > 1 ? atoi(argv[1]) : 2048, NC = argc > 2 ? atoi(argv[2]) : 2048;
    CK(gemm_init() ? cudaErrorUnknown : cudaSuccess);
    double *T, *K, *C, *C0, *h, *q;
    CK(cudaMalloc(&T, sizeof(double) * M * M)); CK(cudaMalloc(&K, sizeof(double) * (size_t)M * NC));
    CK(cudaMalloc(&C, sizeof(double) * M * M)); CK(cudaMalloc(&C0, sizeof(double) * M * M));
    CK(cudaMalloc(&h, sizeof(double) * NC)); CK(cudaMalloc(&q, sizeof(double) * (M / 128) * NC));
    fill<<<(M * M + 255) / 256, 256>>>(T, (long)M * M, 1, M, 1);
    fill<<<((long)M * NC + 255) / 256, 256>>>(K, (long)M * NC, 2, 1, 0);
    fill<<<(NC + 255) / 256, 256>>>(h, NC, 3, 1, 0);
    fill<<<(M * M + 255) / 256, 256>>>(C0, (long)M * M, 4, 1, 0);
    CK(cudaDeviceSynchronize());
    const int ne = 4096;
    std::vector<int> ii(ne), jj(ne);
    int *dii, *djj; double* dout; CK(cudaMalloc(&dii, ne * 4)); CK(cudaMalloc(&djj, ne * 4)); CK(cudaMalloc(&dout, ne * 8));
    std::vector<double> ref(ne), got(ne);

    {   // variance product: q[ti][n] = sum_{i in tile} (sum_k T[k][i] K[k][n])^2
        GemmP p; p.A = T; p.lda = M; p.a_kc = 0; p.a_tri = 2; p.B = K; p.ldb = NC; p.b_kc = 0; p.m = M; p.n = NC; p.k = M;
        p.epilogue = EPI_COLNORM; p.norm_out = q; p.ldn = NC;
        double ms = time_launch(p, 20);
        double alg = (double)M * M * NC;   // algorithmic: triangular, M^2 flops per point
        printf("variance  M=%d nc=%d : %.3f ms  %.2f TFLOP/s algorithmic (%.2f executed)\n", M, NC, ms, alg / ms / 1e9,
               alg * (1.0 + 128.0 / M) / ms / 1e9);
        // check through the STORE epilogue of the same instantiation family
        GemmP s = p; s.epilogue = EPI_STORE; s.C = C; s.ldc = NC > M ? M : NC; 
    }
    {   // weighted SYRK: C = C0*1 + K diag(h) K^T, lower tiles
        GemmP p; p.A = K; p.lda = NC; p.a_kc = 1; p.B = K; p.ldb = NC; p.b_kc = 1; p.C = C; p.ldc = M; p.m = p.n = M; p.k = NC;
        p.beta = 1.0; p.lower_out = 1; p.kscale = h;
        double ms = time_launch(p, 20);
        double alg = (double)M * M * NC;
        printf("syrk      M=%d nc=%d : %.3f ms  %.2f TFLOP/s algorithmic (%.2f executed)\n", M, NC, ms, alg / ms / 1e9,
               alg * (1.0 + 128.0 / M) / ms / 1e9);
        CK(cudaMemcpy(C, C0, sizeof(double) * M * M, cudaMemcpyDeviceToDevice));
        gemm_launch(p, 0);
        for (int e = 0; e < ne; ++e) { int i = rand() % M, j = rand() % M; if (j > i) { int t = i; i = j; j = t; } ii[e] = i; jj[e] = j; }
        CK(cudaMemcpy(dii, ii.data(), ne * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(djj, jj.data(), ne * 4, cudaMemcpyHostToDevice));
        ref_entries<<<(ne + 127) / 128, 128>>>(p, C0, dii, djj, ne, dout);
        CK(cudaMemcpy(ref.data(), dout, ne * 8, cudaMemcpyDeviceToHost));
        double err = 0, mx = 0;
        std::vector<double> Ch((size_t)M * M); CK(cudaMemcpy(Ch.data(), C, sizeof(double) * M * M, cudaMemcpyDeviceToHost));
        for (int e = 0; e < ne; ++e) { err = fmax(err, fabs(Ch[(size_t)ii[e] * M + jj[e]] - ref[e])); mx = fmax(mx, fabs(ref[e])); }
        printf("          syrk check: max abs err %.3e (max |ref| %.3e)\n", err, mx);
    }
    for (int v = 0; v < 4; ++v) {   // SYRK variants: scale on/off, balanced split on/off
        GemmP p; p.A = K; p.lda = NC; p.a_kc = 1; p.B = K; p.ldb = NC; p.b_kc = 1; p.C = C; p.ldc = M; p.m = p.n = M; p.k = NC;
        p.beta = 1.0; p.lower_out = 1; p.kscale = (v & 1) ? h : nullptr;
        if (v & 2) { int nt = M / 128; int ksp = balanced_ksplit(nt * (nt + 1) / 2, NC); if (ksp < NC) { p.ksp = ksp; p.C2 = C0; } }
        double ms = time_launch(p, 20);
        double alg = (double)M * M * NC;
        printf("syrk scale=%d split=%d (ksp %d): %.3f ms  %.2f TFLOP/s algorithmic (%.2f executed)\n", v & 1, (v >> 1) & 1, p.ksp, ms,
               alg / ms / 1e9, alg * (1.0 + 128.0 / M) / ms / 1e9);
    }
    {   // warp -> sub-tile maps (GemmP::wmap / wmap_diag): which pairing of warps on an SM sub-partition is fastest?
        const unsigned gmaps[] = {0x73625140u, 0x37265140u, 0x76543210u, 0x76325410u, 0x62735140u, 0x51407362u};
        const unsigned dmaps[] = {0x73625140u, 0x17326054u, 0x16703524u, 0x10765432u};
        for (unsigned gm : gmaps) {
            GemmP p; p.A = T; p.lda = M; p.a_kc = 0; p.a_tri = 2; p.B = K; p.ldb = NC; p.b_kc = 0; p.m = M; p.n = NC; p.k = M;
            p.epilogue = EPI_COLNORM; p.norm_out = q; p.ldn = NC; p.wmap = gm;
            double ms = time_launch(p, 20);
            printf("wmap %08x variance: %.4f ms  %.2f TFLOP/s algorithmic\n", gm, ms, (double)M * M * NC / ms / 1e9);
            GemmP f; f.A = C0; f.lda = M; f.a_kc = 1; f.B = C0; f.ldb = M; f.b_kc = 1; f.C = C; f.ldc = M; f.m = f.n = f.k = M; f.wmap = gm;
            ms = time_launch(f, 10);
            printf("wmap %08x dense full x full^T: %.4f ms  %.2f TFLOP/s\n", gm, ms, 2.0 * M * M * M / ms / 1e9);
        }
        for (unsigned dm : dmaps)
            for (int sc = 0; sc < 2; ++sc) {
                GemmP p; p.A = K; p.lda = NC; p.a_kc = 1; p.B = K; p.ldb = NC; p.b_kc = 1; p.C = C; p.ldc = M; p.m = p.n = M; p.k = NC;
                p.beta = 1.0; p.lower_out = 1; p.kscale = sc ? h : nullptr; p.wmap_diag = dm;
                int nt = M / 128; int ksp = balanced_ksplit(nt * (nt + 1) / 2, NC); if (ksp < NC) { p.ksp = ksp; p.C2 = C0; }
                double ms = time_launch(p, 20);
                GemmP d1 = p; d1.ksp = 0; d1.C2 = nullptr;
                double ms1 = time_launch(d1, 20);
                printf("wmap_diag %08x syrk scale=%d: split(ksp %d) %.4f ms %.2f TFLOP/s alg | unsplit %.4f ms\n", dm, sc, p.ksp, ms,
                       (double)M * M * NC / ms / 1e9, ms1);
            }
        {   // diagonal tiles alone: a 128-row SYRK (one diagonal tile) against one full off-diagonal tile of the same k
            for (unsigned dm : dmaps) {
                GemmP p; p.A = K; p.lda = NC; p.a_kc = 1; p.B = K; p.ldb = NC; p.b_kc = 1; p.C = C; p.ldc = M; p.m = p.n = 128; p.k = NC;
                p.beta = 1.0; p.lower_out = 1; p.wmap_diag = dm;
                double ms = time_launch(p, 20);
                GemmP f = p; f.lower_out = 0;
                double msf = time_launch(f, 20);
                printf("wmap_diag %08x single diagonal tile %.4f ms vs full tile %.4f ms (ratio %.3f)\n", dm, ms, msf, ms / msf);
            }
        }
    }
    for (int variant = 0; variant < 3; ++variant) {   // dense M x M x M products as used by the update phase
        GemmP p; p.C = C; p.ldc = M; p.m = p.n = p.k = M;
        const char* name;
        if (variant == 0) { p.A = C0; p.lda = M; p.a_kc = 1; p.B = T; p.ldb = M; p.b_kc = 0; p.b_tri = 2; name = "full x lower (kc,mc)"; }
        else if (variant == 1) { p.A = T; p.lda = M; p.a_kc = 0; p.a_tri = 2; p.B = C0; p.ldb = M; p.b_kc = 0; name = "lower^T x full (mc,mc)"; }
        else { p.A = C0; p.lda = M; p.a_kc = 1; p.B = C0; p.ldb = M; p.b_kc = 1; name = "full x full^T (kc,kc)"; }
        double ms = time_launch(p, 10);
        double ex = variant == 2 ? 2.0 * M * M * M : 1.0 * M * M * M * (1.0 + 128.0 / M);
        printf("dense %-24s M=%d : %.3f ms  %.2f TFLOP/s executed\n", name, M, ms, ex / ms / 1e9);
        gemm_launch(p, 0);
        for (int e = 0; e < ne; ++e) { ii[e] = rand() % M; jj[e] = rand() % M; }
        CK(cudaMemcpy(dii, ii.data(), ne * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(djj, jj.data(), ne * 4, cudaMemcpyHostToDevice));
        ref_entries<<<(ne + 127) / 128, 128>>>(p, C0, dii, djj, ne, dout);
        CK(cudaMemcpy(ref.data(), dout, ne * 8, cudaMemcpyDeviceToHost));
        std::vector<double> Ch((size_t)M * M); CK(cudaMemcpy(Ch.data(), C, sizeof(double) * M * M, cudaMemcpyDeviceToHost));
        double err = 0, mx = 0;
        for (int e = 0; e < ne; ++e) { err = fmax(err, fabs(Ch[(size_t)ii[e] * M + jj[e]] - ref[e])); mx = fmax(mx, fabs(ref[e])); }
        printf("          check: max abs err %.3e (max |ref| %.3e)\n", err, mx);
    }
    return 0;
}
