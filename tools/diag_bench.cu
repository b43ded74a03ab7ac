// Tooling: times the 128x128 diagonal-block kernels (Cholesky + inverse, inverse alone) and the blocked chol_lower / trtri_lower.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../t-svgp_b200/csrc/kernels.cuh"
#include "../t-svgp_b200/csrc/dense.cuh"
#include "../t-svgp_b200/csrc/gemm.cuh"
#include "../t-svgp_b200/csrc/common.cuh"
namespace tsvgp { thread_local long g_launches = 0; int g_debug_sync = 0; }
using namespace tsvgp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 2048;
    gemm_init(); diag_init();
    std::vector<double> A((size_t)n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = exp(-0.5 * (i - j) * (i - j) / 9.0) + (i == j ? 0.5 : 0.0);
    double *dA, *dW, *dinv, *dLinv, *tmp; int* info;
    CK(cudaMalloc(&dA, sizeof(double) * n * n)); CK(cudaMalloc(&dW, sizeof(double) * n * n)); CK(cudaMalloc(&dLinv, sizeof(double) * n * n));
    CK(cudaMalloc(&tmp, sizeof(double) * n * n)); CK(cudaMalloc(&dinv, sizeof(double) * (n / 128) * 128 * 128)); CK(cudaMalloc(&info, 16)); CK(cudaMemset(info, 0, 16));
    CK(cudaMemcpy(dA, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemcpy(dW, dA, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
        cudaEventRecord(e0); diag_potrf_inv_launch(dW, n, dinv, 0, info, 0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("diag_potrf_inv (1 block): %.1f us\n", ms * 1e3);
    }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); diag_trtri_launch(dW, n, dinv, 1, 0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("diag_trtri (1 block): %.1f us\n", ms * 1e3);
    }
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemcpy(dW, dA, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
        cudaEventRecord(e0); chol_lower(dW, n, n, dinv, info, 0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("chol_lower n=%d: %.3f ms  (%.2f TFLOP/s of n^3/3)\n", n, ms, (double)n * n * n / 3 / ms / 1e9);
        cudaEventRecord(e0); trtri_lower(dW, n, n, dinv, dLinv, tmp, 0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("trtri_lower n=%d: %.3f ms\n", n, ms);
    }
    int h; CK(cudaMemcpy(&h, info, 4, cudaMemcpyDeviceToHost)); printf("info %d\n", h);
    std::vector<double> L((size_t)n * n), Li((size_t)n * n);
    CK(cudaMemcpy(L.data(), dW, sizeof(double) * n * n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(Li.data(), dLinv, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    double e1m = 0, e2m = 0;
    for (int t = 0; t < 2000; ++t) {
        int i = rand() % n, j = rand() % (i + 1); double s = 0, u = 0;
        for (int k = 0; k <= j; ++k) s += L[(size_t)i * n + k] * L[(size_t)j * n + k];
        for (int k = j; k <= i; ++k) u += Li[(size_t)i * n + k] * L[(size_t)k * n + j];
        e1m = fmax(e1m, fabs(s - A[(size_t)i * n + j])); e2m = fmax(e2m, fabs(u - (i == j ? 1.0 : 0.0)));
    }
    printf("max |LL^T - A| %.2e   max |Linv L - I| %.2e\n", e1m, e2m);
    return 0;
}
