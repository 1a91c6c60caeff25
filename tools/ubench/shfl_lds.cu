// Micro-benchmark: do warp shuffles and shared-memory loads compete for the same pipe on sm_100a?
// Each warp runs ITER iterations of: NL x LDS.64 (conflict-free, lane-dependent) and NS x SHFL.32 (lane-dependent source).
#include <cstdio>
#include <cuda_runtime.h>

template <int NL, int NS>
__global__ void __launch_bounds__(128) k(double* out, int iters) {
    __shared__ double sm[4][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    sm[w][lane] = lane; sm[w][lane + 32] = lane * 0.5;
    __syncwarp();
    double acc = 0.0;
    int a = lane;
    unsigned lo = lane, hi = lane * 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NL; ++i) acc += sm[w][(a + i * 5) & 63];
#pragma unroll
        for (int i = 0; i < NS; ++i) lo += __shfl_sync(0xffffffffu, hi, (lane + i + it) & 31);
        a = (a + 7) & 63;
        hi += lo;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + lo + hi;
}

template <int NL, int NS>
float run(double* d, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NL, NS><<<148 * 4, 128>>>(d, iters);
    cudaEventRecord(e0);
    k<NL, NS><<<148 * 4, 128>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double* d;
    cudaMalloc(&d, 148 * 4 * 128 * sizeof(double));
    const int it = 20000;
    printf("LDS8        %.3f ms\n", run<8, 0>(d, it));
    printf("SHFL16      %.3f ms\n", run<0, 16>(d, it));
    printf("LDS8+SHFL16 %.3f ms\n", run<8, 16>(d, it));
    printf("LDS4+SHFL8  %.3f ms\n", run<4, 8>(d, it));
    printf("LDS4        %.3f ms\n", run<4, 0>(d, it));
    printf("SHFL8       %.3f ms\n", run<0, 8>(d, it));
    return 0;
}
