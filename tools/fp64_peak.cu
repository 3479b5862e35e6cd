// FP64 peak probe for B200 (sm_100a): DFMA, DMMA m8n8k4, a DFMA+DMMA mix, and cuBLAS DGEMM.
// Establishes the roofline denominator that MEASURED_PEAKS.json lacks (it has only HBM and bf16).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/fp64_peak.cu -lcublas -o tools/fp64_peak
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a, double b) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 1e-9 + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// even warps DFMA, odd warps DMMA: do the two share one pipe?
template <int CH>
__global__ void __launch_bounds__(256) k_mix(double *out, int iters, double a, double b) {
    double c0[CH], c1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 1e-9 + i; c1[i] = i; }
    if ((threadIdx.x >> 5) & 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < CH; ++i) dmma884(c0[i], c1[i], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < CH; ++i) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.major, p.minor, p.clockRate);
    const int sms = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256 * 2));
    const int iters = 20000;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps;
        {   constexpr int CH = 8;
            float ms = time_ms([&] { k_dfma<CH><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * CH * (double)iters * 256.0 * grid;
            printf("{\"probe\": \"dfma\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9);
        }
        {   constexpr int CH = 8;
            float ms = time_ms([&] { k_dmma<CH><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 512.0 * CH * (double)iters * 8.0 * grid;
            printf("{\"probe\": \"dmma884\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9);
        }
        {   constexpr int CH = 8;
            float ms = time_ms([&] { k_mix<CH><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
            double fl = (512.0 * CH * 4.0 + 2.0 * 2 * CH * 32 * 4.0) * (double)iters * grid;
            printf("{\"probe\": \"mix_dfma_dmma\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9);
        }
    }
    // DMMA with a single warp per SMSP (issue-limited?)
    {   constexpr int CH = 16;
        float ms = time_ms([&] { k_dmma<CH><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); });
        double fl = 512.0 * CH * (double)iters * 4.0 * sms;
        printf("{\"probe\": \"dmma884_1warp_per_smsp\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
    }
    // cuBLAS DGEMM
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {4096, 8192}) {
        double *A, *B, *C; size_t bytes = sizeof(double) * (size_t)n * n;
        CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
        double al = 1.0, be = 0.0;
        float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 5);
        printf("{\"probe\": \"cublas_dgemm_tn\", \"n\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", n, ms, 2.0 * n * (double)n * n / ms * 1e-9);
        ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 5);
        printf("{\"probe\": \"cublas_dgemm_nn\", \"n\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", n, ms, 2.0 * n * (double)n * n / ms * 1e-9);
        // sustained: back to back for ~3 s
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        int reps = std::max(3, (int)(3000.0f / ms));
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1); float tot; cudaEventElapsedTime(&tot, e0, e1);
        printf("{\"probe\": \"cublas_dgemm_nn_sustained\", \"n\": %d, \"reps\": %d, \"tflops\": %.2f}\n", n, reps, 2.0 * n * (double)n * n * reps / tot * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    return 0;
}
