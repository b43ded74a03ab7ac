// FP64 peak probe for B200 (sm_100a): DMMA m8n8k4 issue rate, DFMA rate, mixed, cuBLAS DGEMM, cuSOLVER potrf.
// Tooling only (not part of the product library). Build: see tools/Makefile.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: NACC DMMAs + NF DFMAs per iteration
template <int NACC, int NF>
__global__ void k_mixed(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
    double f[NF];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory: warp tile 64x32 (8 A frags + 4 B frags per k4 step, 32 DMMAs), data re-read from smem
__global__ void __launch_bounds__(256, 1) k_dmma_smem(double* out, int iters) {
    extern __shared__ double smem_dyn[];
    double* sA = smem_dyn; double* sB = smem_dyn + 128 * 36;
    for (int i = threadIdx.x; i < 128 * 36; i += blockDim.x) { sA[i] = 1e-3 * (i % 7); sB[i] = 1e-3 * (i % 5); }
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    double c[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sA[(wm + i * 8 + g) * 36 + k + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[(wn + j * 8 + g) * 36 + k + t];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sms %d cc %d.%d clockRate(attr) %d kHz\n", p.name, sms, p.major, p.minor, clk_khz);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 4 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    // DMMA rate for various warps/SM
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_dmma<16><<<sms, threads>>>(out, iters, 1.0, 1.0);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms = time_ms(e0, e1);
            double fma = (double)sms * (threads / 32) * iters * 16.0 * 256.0;
            if (rep) printf("DMMA m8n8k4 NACC=16 threads/SM=%4d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/ns/SM)\n", threads, ms, 2 * fma / ms * 1e-9, fma / ms * 1e-6 / sms);
        }
    }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_dmma<4><<<sms, 256>>>(out, iters, 1.0, 1.0);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1);
        double fma = (double)sms * 8 * iters * 4.0 * 256.0;
        if (rep) printf("DMMA m8n8k4 NACC=4  threads/SM= 256: %.3f ms  %.2f TFLOP/s (latency-bound probe: %.1f ns per dependent DMMA)\n", ms, 2 * fma / ms * 1e-9, ms * 1e6 / iters / 1.0 / 4 * 1.0);
    }
    for (int threads : {128, 256, 512, 1024}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_dfma<16><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms = time_ms(e0, e1);
            double fma = (double)sms * threads * iters * 16.0;
            if (rep) printf("DFMA NACC=16 threads/SM=%4d: %.3f ms  %.2f TFLOP/s (%.1f FMA/ns/SM)\n", threads, ms, 2 * fma / ms * 1e-9, fma / ms * 1e-6 / sms);
        }
    }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_mixed<16, 16><<<sms, 256>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1);
        double fma_mma = (double)sms * 8 * iters * 16.0 * 256.0, fma_f = (double)sms * 256 * iters * 16.0;
        if (rep) printf("MIXED 16 DMMA + 16 DFMA /iter threads/SM=256: %.3f ms  DMMA %.2f TF + DFMA %.2f TF = %.2f TF\n", ms, 2 * fma_mma / ms * 1e-9, 2 * fma_f / ms * 1e-9, 2 * (fma_mma + fma_f) / ms * 1e-9);
    }
    CK(cudaFuncSetAttribute(k_dmma_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 36 * 8));
    for (int rep = 0; rep < 2; ++rep) {
        int it2 = 2000;
        cudaEventRecord(e0);
        k_dmma_smem<<<sms, 256, 2 * 128 * 36 * 8>>>(out, it2);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1);
        double fma = (double)sms * 8 * it2 * 8.0 * 32.0 * 256.0;
        if (rep) printf("DMMA smem-fed 128x128 tile (8 warps 64x32): %.3f ms  %.2f TFLOP/s\n", ms, 2 * fma / ms * 1e-9);
    }
    // sustained DMMA for ~2 s to see power-capped clocks
    {
        cudaEventRecord(e0);
        for (int r = 0; r < 40; ++r) k_dmma<16><<<sms, 512>>>(out, iters * 4, 1.0, 1.0);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1);
        double fma = 40.0 * sms * 16 * iters * 4 * 16.0 * 256.0;
        printf("DMMA sustained (%.0f ms): %.2f TFLOP/s\n", ms, 2 * fma / ms * 1e-9);
    }
    // cuBLAS DGEMM
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {2048, 4096, 8192}) {
        double *A, *B, *C; size_t bytes = (size_t)n * n * 8;
        CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        std::vector<double> hA((size_t)n * n); for (size_t i = 0; i < hA.size(); ++i) hA[i] = (double)(i % 1013) * 1e-3;
        CK(cudaMemcpy(A, hA.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hA.data(), bytes, cudaMemcpyHostToDevice));
        double one = 1, zero = 0;
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms = time_ms(e0, e1); if (ms < best) best = ms;
        }
        printf("cuBLAS DGEMM NT n=%d: best %.3f ms  %.2f TFLOP/s\n", n, best, 2.0 * n * n * n / best * 1e-9);
        if (n == 8192) {
            cudaEventRecord(e0);
            int reps = 40;
            for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms = time_ms(e0, e1);
            printf("cuBLAS DGEMM NT n=8192 sustained x%d (%.0f ms): %.2f TFLOP/s\n", reps, ms, reps * 2.0 * n * n * n / ms * 1e-9);
            // DSYRK
            cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &one, A, n, &zero, C, n);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &one, A, n, &zero, C, n);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            ms = time_ms(e0, e1);
            printf("cuBLAS DSYRK n=k=8192: %.3f ms  %.2f TFLOP/s (n^3 flops)\n", ms, 1.0 * n * n * n / ms * 1e-9);
        }
        // cuSOLVER potrf on SPD matrix
        {
            cusolverDnHandle_t sh; cusolverDnCreate(&sh);
            for (size_t i = 0; i < (size_t)n; ++i) for (size_t j = 0; j < (size_t)n; ++j) hA[i * n + j] = (i == j) ? n : 1.0 / (1.0 + (i > j ? i - j : j - i));
            CK(cudaMemcpy(A, hA.data(), bytes, cudaMemcpyHostToDevice));
            int lwork = 0; cusolverDnDpotrf_bufferSize(sh, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork);
            double* work; CK(cudaMalloc(&work, sizeof(double) * lwork)); int* info; CK(cudaMalloc(&info, 4));
            float bestp = 1e30f;
            for (int r = 0; r < 3; ++r) {
                CK(cudaMemcpy(C, A, bytes, cudaMemcpyDeviceToDevice));
                cudaEventRecord(e0);
                cusolverDnDpotrf(sh, CUBLAS_FILL_MODE_LOWER, n, C, n, work, lwork, info);
                cudaEventRecord(e1); CK(cudaDeviceSynchronize());
                float ms = time_ms(e0, e1); if (ms < bestp) bestp = ms;
            }
            printf("cuSOLVER DPOTRF n=%d: best %.3f ms  %.2f TFLOP/s\n", n, bestp, (1.0 / 3.0) * n * n * n / bestp * 1e-9);
            cudaFree(work); cudaFree(info); cusolverDnDestroy(sh);
        }
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    return 0;
}
