// FP64 peak microbenchmarks for the roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
//   1. DFMA: register-resident fused multiply-add chains (the CUDA-core FP64 pipe)
//   2. DMMA: mma.sync.aligned.m8n8k4.f64 chains (the FP64 tensor path reachable on sm_100a;
//            tcgen05 has no FP64 type)
//   3. cuBLAS DGEMM 8192^3 (burst: best of 10)
//   4. shared-memory-fed 48x48x48 DMMA product loop (what the expm kernel actually does)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/fp64_peak.cu -lcublas -o tools/fp64_peak
// Output: one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NT>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
    double c0[NT], c1[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
    double a = 1e-3 * (threadIdx.x & 31), b = 1e-3 * (threadIdx.x >> 5);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Shared-memory-fed product: 4 warps compute C(48x48) = A(48x48) * B(48x48), each warp a 24x24
// block (3x3 DMMA tiles), operands row-major with a padded stride, repeated `reps` times.
constexpr int LD = 52;
__global__ void __launch_bounds__(128) smem_dmma_kernel(double* out, int reps) {
    __shared__ double sA[48 * LD];
    __shared__ double sB[48 * LD];
    for (int i = threadIdx.x; i < 48 * LD; i += blockDim.x) { sA[i] = 1e-3 * (i % 7); sB[i] = 1e-3 * (i % 5); }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = (warp >> 1) * 24, c0 = (warp & 1) * 24;
    double acc[3][3][2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 4
        for (int k = 0; k < 48; k += 4) {
            double a[3], b[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) a[i] = sA[(r0 + 8 * i + g) * LD + k + t];
#pragma unroll
            for (int j = 0; j < 3; ++j) b[j] = sB[(k + t) * LD + c0 + 8 * j + g];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double best_ms(F launch, int tries = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int i = 0; i < tries; ++i) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, (double)ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 256));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);

    {   // DFMA: several occupancies
        const int iters = 1 << 14;
        double bestT = 0;
        for (int ctas = 2; ctas <= 8; ctas *= 2) {
            double ms = best_ms([&] { dfma_kernel<16><<<sms * ctas, 256>>>(out, iters, 1.0000001, 1e-9); });
            double tf = 2.0 * 16 * iters * 256.0 * sms * ctas / (ms * 1e-3) / 1e12;
            bestT = std::max(bestT, tf);
            printf(", \"dfma_tflops_%dcta\": %.3f", ctas, tf);
        }
        printf(", \"dfma_tflops\": %.3f", bestT);
    }
    {   // DMMA
        const int iters = 1 << 13;
        double bestT = 0;
        for (int ctas = 1; ctas <= 8; ctas *= 2) {
            double ms = best_ms([&] { dmma_kernel<8><<<sms * ctas, 256>>>(out, iters); });
            double tf = 2.0 * 256.0 * 8 * iters * 8 /*warps*/ * sms * ctas / (ms * 1e-3) / 1e12;
            bestT = std::max(bestT, tf);
            printf(", \"dmma_tflops_%dcta\": %.3f", ctas, tf);
        }
        printf(", \"dmma_tflops\": %.3f", bestT);
    }
    {   // smem-fed 48^3 products
        const int reps = 2000;
        for (int ctas = 1; ctas <= 16; ctas *= 2) {
            double ms = best_ms([&] { smem_dmma_kernel<<<sms * ctas, 128>>>(out, reps); });
            double tf = 2.0 * 48 * 48 * 48 * (double)reps * sms * ctas / (ms * 1e-3) / 1e12;
            printf(", \"smem_dmma48_tflops_%dcta\": %.3f", ctas, tf);
        }
    }
    {   // cuBLAS DGEMM
        const int n = 8192;
        double *A, *B, *C;
        CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
        CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
        cublasHandle_t h; cublasCreate(&h);
        const double one = 1.0, zero = 0.0;
        double ms = best_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 10);
        printf(", \"cublas_dgemm_tflops\": %.3f", 2.0 * n * n * (double)n / (ms * 1e-3) / 1e12);
        // sustained: back to back for ~3 s
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int cnt = std::max(1, (int)(3000.0 / ms));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < cnt; ++i) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float tot; CK(cudaEventElapsedTime(&tot, e0, e1));
        printf(", \"cublas_dgemm_tflops_sustained\": %.3f", 2.0 * n * n * (double)n * cnt / (tot * 1e-3) / 1e12);
        cublasDestroy(h);
    }
    {   // sustained DFMA for ~2 s
        const int iters = 1 << 16;
        double ms1 = best_ms([&] { dfma_kernel<16><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9); }, 2);
        int cnt = std::max(1, (int)(2000.0 / ms1));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < cnt; ++i) dfma_kernel<16><<<sms * 8, 256>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float tot; CK(cudaEventElapsedTime(&tot, e0, e1));
        printf(", \"dfma_tflops_sustained\": %.3f", 2.0 * 16 * iters * 256.0 * sms * 8 * cnt / (tot * 1e-3) / 1e12);
    }
    printf("}\n");
    return 0;
}
