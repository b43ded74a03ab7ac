// Tooling: a prototype of the ON-CHIP variant of the streaming pass that DESIGN §3 argues against, so that the
// argument rests on a measurement.  M = 2048 inducing points = one 16-CTA thread-block cluster; CTA r of a cluster builds the
// covariance block  K_r = k(Z[128 r : 128 r + 128], X[panel])  [128 x 128 points] in ITS shared memory and the 16 blocks of a panel
// never leave the chip: the weighted SYRK  B += K diag(h) K^T  of the panel reads them through distributed shared memory (DSMEM).
// The 136 lower output tiles do not fit on chip (64 accumulator registers per thread = ONE 128 x 128 tile per CTA), so every tile
// of every 128-point panel ends in 16 384 `red.global.add.f64` into the L2-resident B (17.8 MB lower half).
//
//   phase 1  each CTA: K_r for the panel (Matern-5/2 or SE, lengthscale-scaled inputs)        -> own shared memory, [128][132]
//            cluster.sync
//   phase 2  each CTA: 8.5 of the 136 lower tiles (rows r and 15 - r of a CTA pair have 17 tiles; the middle one is split in k),
//            operands pulled from the owners' shared memory 16 points at a time (register-staged prefetch, double-buffered local
//            staging in the DMMA engine's [128][20] layout, the weights h applied on the way), 64 x 32 warp tiles of m8n8k4 DMMA,
//            then red.add of the tile into B
//            cluster.sync                                                                      (peers are done with K_r)
//
// Prints per-128-point time over the whole GPU next to the product's materialised-slab numbers (kuf_kernel + SYRK per 16384-point
// slab from profiles/ncu_summary.json), the co-resident cluster count, and the worst relative error against a plain FP64 reference.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o onchip_bench onchip_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int CL = 16;            // CTAs per cluster
constexpr int MB = 128;           // inducing rows per CTA
constexpr int M = CL * MB;        // 2048
constexpr int P = 128;            // points per panel
constexpr int LDK = P + 4;        // resident block [row][point], rows 1056 B apart: conflict-free fragment reads either way
constexpr int KC = 16;            // points per staged chunk
constexpr int LDS = KC + 4;       // staged chunk [row][16 points] (the DMMA engine's k-contiguous layout)
constexpr int STAGE = MB * LDS;   // doubles per operand per buffer
constexpr int NT = 256;
constexpr int DMAX = 16;
constexpr size_t SMEM_BYTES = sizeof(double) * ((size_t)MB * LDK + 4 * STAGE + P);

struct Args {
    const double* ZsT;   // [D][M] scaled inducing inputs, feature-major
    const double* z2;    // [M]
    const double* XsT;   // [D][n] scaled inputs, feature-major
    const double* x2;    // [n]
    const double* h;     // [n] weights
    double* B;           // [M][M], lower tiles accumulated
    int D, n, kind;      // kind 0 = SE, 1 = Matern-5/2
    double var;
    int phase_mask;      // 1 = build K_r, 2 = SYRK, 4 = red.add epilogue
    int split;           // 1 = the middle tile of a CTA pair is split between the two (8.5 tiles each)
    long long* stamps;   // optional [grid][4] clock64 stamps of CTA phases
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double kern(int kind, double var, double r2) {
    r2 = fmax(r2, 0.0);
    if (kind == 0) return var * exp(-0.5 * r2);
    const double r = sqrt(5.0 * r2);
    return var * (1.0 + r + r * r * (1.0 / 3.0)) * exp(-r);
}

// tile list of a CTA pair (r, 15 - r), r < 8: the 16 - r tiles of row 15 - r, then the r + 1 tiles of row r = 17 tiles = 136 chunk steps
// of 16 points.  The upper CTA takes steps [0, 68) (its own row: A operand local), the lower one [68, 136): the tile in the middle is
// split 4 + 4 chunks between the two and both add their half into B, so every CTA of the cluster does 8.5 tiles per panel.
// (split = 0: whole tiles, 8 / 9 alternating with the panel parity — what the first measurement used.)
__device__ __forceinline__ void my_steps(int rank, int parity, int split, int& s_begin, int& s_end, int& lo) {
    lo = rank < 8 ? rank : 15 - rank;
    const int cut = split ? 68 : 8 * (8 + ((parity + lo) & 1));
    if (rank >= 8) { s_begin = 0; s_end = cut; } else { s_begin = cut; s_end = 136; }
}
__device__ __forceinline__ void tile_of(int lo, int idx, int& ti, int& tj) {
    const int hi = 15 - lo;
    if (idx < 16 - lo) { ti = hi; tj = idx; } else { ti = lo; tj = idx - (16 - lo); }
}

__global__ void __launch_bounds__(NT, 1) onchip_pass(Args a) {
    extern __shared__ __align__(16) double smem[];
    double* Kblk = smem;                       // [MB][LDK]
    double* stage = smem + MB * LDK;           // [2 buffers][A, B][MB][LDS]
    double* hs = stage + 4 * STAGE;            // [P]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const int npanels = a.n / P;
    long long st[4] = {0, 0, 0, 0};

    for (int panel = cid, it = 0; panel < npanels; panel += ncl, ++it) {
        const long n0 = (long)panel * P;
        long long c0 = clock64();
        // ---- phase 1: K_r for this panel.  Thread -> point p = tid % 128, rows tid / 128 + 2 k.  Z rows wait in the (idle) staging area.
        if (tid < P) hs[tid] = a.h[n0 + tid];
        if (!(a.phase_mask & 1) && it == 0)    // timing variant without the build: any finite block
            for (int e = tid; e < MB * LDK; e += NT) Kblk[e] = 1.0 / (1 + e % 7);
        if (a.phase_mask & 1) {
            double* zs = stage;                // [D][MB] + z2 at [D * MB ..]
            for (int e = tid; e < a.D * MB; e += NT) zs[e] = a.ZsT[(long)(e / MB) * M + rank * MB + e % MB];
            for (int e = tid; e < MB; e += NT) zs[a.D * MB + e] = a.z2[rank * MB + e];
            const int p = tid & (P - 1);
            double x[DMAX];
#pragma unroll
            for (int d = 0; d < DMAX; ++d) x[d] = d < a.D ? a.XsT[(long)d * a.n + n0 + p] : 0.0;
            const double xx = a.x2[n0 + p];
            __syncthreads();
            for (int row = tid >> 7; row < MB; row += 2) {
                double dot = 0.0;
#pragma unroll
                for (int d = 0; d < DMAX; ++d)
                    if (d < a.D) dot = fma(zs[d * MB + row], x[d], dot);
                Kblk[row * LDK + p] = kern(a.kind, a.var, zs[a.D * MB + row] + xx - 2.0 * dot);
            }
        }
        cluster.sync();
        long long c1 = clock64();
        // ---- phase 2: this CTA's tiles of the panel
        if (a.phase_mask & 2) {
            int s_begin, s_end, lo;
            my_steps(rank, it, a.split, s_begin, s_end, lo);
            double acc[8][4][2];
            double2 pa[4], pb[4];
            auto fetch = [&](int s) {          // operands of step s (tile s / 8, chunk s % 8) from their owners' shared memory into registers
                int ti, tj;
                tile_of(lo, s / (P / KC), ti, tj);
                const int kc = (s % (P / KC)) * KC;
                const double* srcA = cluster.map_shared_rank(Kblk, ti);
                const double* srcB = cluster.map_shared_rank(Kblk, tj);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int idx = tid + c * NT, r = idx >> 3, cc = (idx & 7) * 2;
                    pa[c] = *reinterpret_cast<const double2*>(srcA + r * LDK + kc + cc);
                    pb[c] = *reinterpret_cast<const double2*>(srcB + r * LDK + kc + cc);
                }
            };
            fetch(s_begin);
            for (int s = s_begin; s < s_end; ++s) {
                const int kc = (s % (P / KC)) * KC;
                double* sa = stage + (s & 1) * 2 * STAGE;
                double* sb = sa + STAGE;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int idx = tid + c * NT, r = idx >> 3, cc = (idx & 7) * 2;
                    *reinterpret_cast<double2*>(sa + r * LDS + cc) = pa[c];
                    double2 w = pb[c];
                    w.x *= hs[kc + cc]; w.y *= hs[kc + cc + 1];
                    *reinterpret_cast<double2*>(sb + r * LDS + cc) = w;
                }
                __syncthreads();
                if (s + 1 < s_end) fetch(s + 1);
                if (kc == 0 || s == s_begin) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
                }
#pragma unroll
                for (int kk = 0; kk < KC; kk += 4) {
                    double fa[8], fb[4];
#pragma unroll
                    for (int i = 0; i < 8; ++i) fa[i] = sa[(wm + 8 * i + g) * LDS + kk + t];
#pragma unroll
                    for (int j = 0; j < 4; ++j) fb[j] = sb[(wn + 8 * j + g) * LDS + kk + t];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
                }
                if ((kc == P - KC || s == s_end - 1) && (a.phase_mask & 4)) {   // tile (or this CTA's half of it) complete: add it into B (fire-and-forget reductions at L2)
                    int ti, tj;
                    tile_of(lo, s / (P / KC), ti, tj);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        double* row = a.B + (long)(ti * MB + wm + 8 * i + g) * M + tj * MB + wn + 2 * t;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            atomicAdd(row + 8 * j, acc[i][j][0]);
                            atomicAdd(row + 8 * j + 1, acc[i][j][1]);
                        }
                    }
                }
            }
            if (!(a.phase_mask & 4)) {   // keep the accumulators alive
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) v += acc[i][j][0] + acc[i][j][1];
                if (v == 1.2345e-300) a.B[0] = v;
            }
        }
        long long c2 = clock64();
        cluster.sync();                        // nobody reads K_r any more
        long long c3 = clock64();
        st[0] += c1 - c0; st[1] += c2 - c1; st[2] += c3 - c2; st[3] += 1;
    }
    if (a.stamps && tid == 0)
        for (int k = 0; k < 4; ++k) a.stamps[(long)blockIdx.x * 4 + k] = st[k];
}

// plain FP64 reference of the lower triangle: B[i][j] = sum_n h_n K[i][n] K[j][n]
__global__ void ref_syrk(Args a, double* Bref) {
    const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
    if (j > i) return;
    double s = 0.0;
    for (int n = 0; n < a.n; ++n) {
        double di = 0.0, dj = 0.0;
        for (int d = 0; d < a.D; ++d) {
            const double x = a.XsT[(long)d * a.n + n];
            di = fma(a.ZsT[(long)d * M + i], x, di);
            dj = fma(a.ZsT[(long)d * M + j], x, dj);
        }
        const double ki = kern(a.kind, a.var, a.z2[i] + a.x2[n] - 2.0 * di), kj = kern(a.kind, a.var, a.z2[j] + a.x2[n] - 2.0 * dj);
        s = fma(a.h[n] * ki, kj, s);
    }
    Bref[(long)i * M + j] = s;
}

int main(int argc, char** argv) {
    const int n_time = argc > 1 ? atoi(argv[1]) : 147456;   // 1152 panels: a multiple of 8 and 9 clusters
    const int D = 16, kind = 1;
    const int n_check = 9 * P * 2;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, smem/CTA opt-in %zu B; prototype needs %zu B\n", prop.name, prop.multiProcessorCount, (size_t)prop.sharedMemPerBlockOptin, SMEM_BYTES);
    srand(1);
    auto rnd = [] { return (double)rand() / RAND_MAX * 2.0 - 1.0; };
    const int n_max = n_time > n_check ? n_time : n_check;
    std::vector<double> ZsT((size_t)D * M), z2(M, 0.0), XsT((size_t)D * n_max), x2(n_max, 0.0), h(n_max);
    for (auto& v : ZsT) v = rnd() * 0.6;
    for (int i = 0; i < M; ++i) for (int d = 0; d < D; ++d) z2[i] += ZsT[(size_t)d * M + i] * ZsT[(size_t)d * M + i];
    for (auto& v : h) v = -0.5 - 0.4 * rnd();
    double *dZ, *dz2, *dX, *dx2, *dh, *dB, *dBref;
    long long* dst;
    CK(cudaMalloc(&dZ, sizeof(double) * D * M)); CK(cudaMalloc(&dz2, sizeof(double) * M)); CK(cudaMalloc(&dX, sizeof(double) * D * n_max));
    CK(cudaMalloc(&dx2, sizeof(double) * n_max)); CK(cudaMalloc(&dh, sizeof(double) * n_max)); CK(cudaMalloc(&dB, sizeof(double) * M * M));
    CK(cudaMalloc(&dBref, sizeof(double) * M * M)); CK(cudaMalloc(&dst, sizeof(long long) * 4 * 1024));
    CK(cudaMemcpy(dZ, ZsT.data(), sizeof(double) * D * M, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dz2, z2.data(), sizeof(double) * M, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dh, h.data(), sizeof(double) * n_max, cudaMemcpyHostToDevice));

    CK(cudaFuncSetAttribute(onchip_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    CK(cudaFuncSetAttribute(onchip_pass, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.attrs = attr; cfg.numAttrs = 1; cfg.stream = 0;
    cfg.gridDim = dim3(CL * 9);
    int ncl_max = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl_max, onchip_pass, &cfg));
    printf("co-resident 16-CTA clusters (cudaOccupancyMaxActiveClusters): %d  -> %d of %d SMs busy\n", ncl_max, ncl_max * CL, prop.multiProcessorCount);
    if (ncl_max < 1) { printf("the cluster shape does not fit this device\n"); return 1; }
    cfg.gridDim = dim3(CL * ncl_max);

    auto upload = [&](int n) {   // feature-major [D][n]
        std::vector<double> xt((size_t)D * n), xx(n, 0.0);
        srand(7);
        for (auto& v : xt) v = rnd() * 0.6;
        for (int p = 0; p < n; ++p) for (int d = 0; d < D; ++d) xx[p] += xt[(size_t)d * n + p] * xt[(size_t)d * n + p];
        CK(cudaMemcpy(dX, xt.data(), sizeof(double) * D * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dx2, xx.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    };
    Args a{dZ, dz2, dX, dx2, dh, dB, D, n_check, kind, 1.3, 7, 1, nullptr};

    for (int split = 0; split < 2; ++split) {   // ---- correctness on 18 panels
        upload(n_check);
        a.split = split;
        CK(cudaMemset(dB, 0, sizeof(double) * M * M)); CK(cudaMemset(dBref, 0, sizeof(double) * M * M));
        CK(cudaLaunchKernelEx(&cfg, onchip_pass, a));
        CK(cudaDeviceSynchronize());
        ref_syrk<<<dim3(M / 16, M / 16), dim3(16, 16)>>>(a, dBref);
        CK(cudaDeviceSynchronize());
        std::vector<double> Bh((size_t)M * M), Br((size_t)M * M);
        CK(cudaMemcpy(Bh.data(), dB, sizeof(double) * M * M, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(Br.data(), dBref, sizeof(double) * M * M, cudaMemcpyDeviceToHost));
        double worst = 0.0, scale = 0.0;
        for (int i = 0; i < M; ++i) for (int j = 0; j <= i; ++j) scale = fmax(scale, fabs(Br[(size_t)i * M + j]));
        for (int i = 0; i < M; ++i) for (int j = 0; j <= i; ++j) worst = fmax(worst, fabs(Bh[(size_t)i * M + j] - Br[(size_t)i * M + j]));
        printf("check (%d points, split %d): max |B - B_ref| / max |B_ref| over the lower triangle = %.3e  (max |B_ref| %.4g)\n", n_check, split, worst / scale, scale);
    }

    // ---- timing
    upload(n_time);
    a.n = n_time;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double flop_panel = 2.0 * 136 * 128.0 * 128.0 * P;   // executed DMMA flops of one panel (full diagonal tiles, like the product)
    const double alg_panel = (double)M * M * P;                // algorithmic: M^2 flop per point
    struct Variant { const char* name; int mask, split; } variants[] = {
        {"K_r build + SYRK + red.add (the full on-chip SYRK)", 7, 1}, {"  same, whole tiles (8 / 9 per CTA)", 7, 0}, {"K_r build + SYRK, no red.add", 3, 1},
        {"SYRK + red.add, K_r stale", 6, 1}, {"K_r build only (+ 2 cluster.sync per panel)", 1, 1}};
    const int only = argc > 2 ? atoi(argv[2]) : -1;   // ncu: time one variant only
    for (int vi = 0; vi < 5; ++vi) {
        if (only >= 0 && vi != only) continue;
        const Variant& v = variants[vi];
        a.phase_mask = v.mask; a.split = v.split; a.stamps = dst;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemsetAsync(dB, 0, sizeof(double) * M * M));
            CK(cudaEventRecord(e0));
            CK(cudaLaunchKernelEx(&cfg, onchip_pass, a));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        const int npanels = n_time / P;
        std::vector<long long> sth(4 * CL * ncl_max);
        CK(cudaMemcpy(sth.data(), dst, sizeof(long long) * sth.size(), cudaMemcpyDeviceToHost));
        double ph[3] = {0, 0, 0}, cnt = 0;
        for (int b = 0; b < CL * ncl_max; ++b) { for (int k = 0; k < 3; ++k) ph[k] += (double)sth[4 * b + k]; cnt += (double)sth[4 * b + 3]; }
        printf("%-52s %8.3f ms / %d points = %6.2f us per 128 points (whole GPU)", v.name, best, n_time, best * 1e3 / npanels);
        if (v.mask & 2) printf("  %5.2f TFLOP/s executed, %5.2f algorithmic", flop_panel * npanels / best * 1e-9, alg_panel * npanels / best * 1e-9);
        printf("\n    per CTA and panel (cycles): build + sync %.0f, tiles %.0f, wait at the closing sync %.0f\n", ph[0] / cnt, ph[1] / cnt, ph[2] / cnt);
    }
    printf("product (materialised slab, profiles/ncu_summary.json kernels_r02_16384): kuf 174.7 us + SYRK 2185 us per 16384 points = %.2f us per 128 points\n",
           (174.7 + 2185.0) / 128.0);
    return 0;
}
