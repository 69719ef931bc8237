// fp32_issue.cu -- how many sequentially-rounded fp32 adds can one B200 SM retire per clock?
// K1's box sum needs 24 dependent-order float adds per pixel (no reassociation allowed), so the
// kernel is bound by FADD issue, not by HBM. This microbenchmark measures the candidates:
//   add.rn.f32 (FADD), add.rn.f32x2 (packed, sm_100+), max.f32 (FMNMX), 3-input max.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_issue fp32_issue.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) issue_kernel(float* out, const float* in, int iters, long long* cycles) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 7]; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            if (MODE == 0) {             // 8 independent FADD chains
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            } else if (MODE == 1) {      // 4 independent packed chains = 8 adds
#pragma unroll
                for (int i = 0; i < 8; i += 2)
                    asm volatile("{ .reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x; }"
                                 : "+f"(a[i]), "+f"(a[i + 1]) : "f"(b[i]), "f"(b[i + 1]));
            } else if (MODE == 2) {      // 8 FMNMX
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            } else if (MODE == 3) {      // 8 three-input max
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) & 7]));
            } else if (MODE == 4) {      // 6 FADD + 2 FMNMX interleaved (fma pipe + alu pipe)
#pragma unroll
                for (int i = 0; i < 6; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[6]) : "f"(b[6]));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[7]) : "f"(b[7]));
            } else if (MODE == 5) {      // 8 FFMA (a = a * 1 + b) for comparison
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[(i + 3) & 7]), "f"(b[i]));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int ops_per_inner) {
    const int sms = 148, iters = 4096;
    const int grid = sms * ctas_per_sm;
    float *out, *in;
    long long* cyc;
    CHECK(cudaMalloc(&out, sizeof(float) * grid * 256));
    CHECK(cudaMalloc(&in, sizeof(float) * 1024));
    CHECK(cudaMemset(in, 0, sizeof(float) * 1024));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * grid));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    issue_kernel<MODE><<<grid, 256>>>(out, in, 16, cyc);
    CHECK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    issue_kernel<MODE><<<grid, 256>>>(out, in, iters, cyc);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)malloc(sizeof(long long) * grid);
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    const double lane_ops = (double)grid * 256 * iters * 8.0 * ops_per_inner;      // scalar ops
    const double per_clk_sm = (double)ctas_per_sm * 256 * iters * 8.0 * ops_per_inner / mean;
    printf("%-34s ctas/SM %d  %8.3f ms  %8.2f Tops/s  %7.1f lane-ops/clk/SM  (mean %.0f clk => %.0f MHz)\n", name, ctas_per_sm, ms,
           lane_ops / ms / 1e9, per_clk_sm, mean, mean / ms / 1e3);
    free(h);
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
    for (int c = 2; c <= 8; c *= 2) {
        run<0>("add.rn.f32 (FADD)", c, 8);
        run<1>("add.rn.f32x2 (packed)", c, 8);
        run<2>("max.f32 (FMNMX)", c, 8);
        run<3>("max.f32 3-input", c, 8);
        run<4>("6 FADD + 2 FMNMX", c, 8);
        run<5>("fma.rn.f32 (FFMA)", c, 8);
    }
    return 0;
}
