// ubench_fp32.cu — what is the FP32-pipe roofline of the L1 inner loop on this B200?
// Measures issue throughput (lane-ops / clk / SM) of the instruction patterns pairwise.cu can use:
//   fadd      : acc += x                      (scalar FADD)
//   l1_scalar : d = a - b; acc += |d|         (2 scalar FADD per element)
//   l1_packed : FADD2 d = a.bcast - b2; FADD2 acc2 += |d|   (2 packed per 2 elements)
//   l2_packed : FADD2 + FFMA2
//   ffma / ffma2 : plain FMA throughput
//   l1_minmix : 3/4 of the elements via FADD2 pair, 1/4 via FMNMX (ALU pipe) + scalar FADD
//   l1_minpacked_* : 1/2, 3/4 or all of the elements via FMNMX + ONE packed FADD2 per two elements
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp32 ubench_fp32.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define ITERS 2048

__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub_bc(float a, uint64_t b) { uint64_t aa = pk(a, a), d; asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(aa), "l"(b)); return d; }
__device__ __forceinline__ uint64_t add_abs(uint64_t acc, uint64_t d) { float lo, hi; unpk(d, lo, hi); uint64_t ad = pk(fabsf(lo), fabsf(hi)); asm("add.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(ad)); return acc; }
__device__ __forceinline__ uint64_t fma_sq(uint64_t acc, uint64_t d) { asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(d)); return acc; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b)); return c; }

template <int KIND>
__global__ void __launch_bounds__(256) bench(const float *__restrict__ in, float *out, long long *cycles)
{
    // 8 a-values x 8 b-values -> 64 element updates per iteration, like one k-step of the tile kernel
    float a[8], b[8];
    for (int i = 0; i < 8; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 256]; }
    float acc[8][8];
    uint64_t acc2[8][4];
    for (int m = 0; m < 8; m++) for (int q = 0; q < 8; q++) acc[m][q] = 0.f;
    for (int m = 0; m < 8; m++) for (int q = 0; q < 4; q++) acc2[m][q] = 0ull;
    uint64_t b2[4] = {pk(b[0], b[1]), pk(b[2], b[3]), pk(b[4], b[5]), pk(b[6], b[7])};
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (KIND == 0) {
#pragma unroll
            for (int m = 0; m < 8; m++)
#pragma unroll
                for (int q = 0; q < 8; q++) acc[m][q] += a[m];
        } else if (KIND == 1) {
#pragma unroll
            for (int m = 0; m < 8; m++)
#pragma unroll
                for (int q = 0; q < 8; q++) acc[m][q] += fabsf(a[m] - b[q]);
        } else if (KIND == 2) {
#pragma unroll
            for (int m = 0; m < 8; m++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc2[m][q] = add_abs(acc2[m][q], sub_bc(a[m], b2[q]));
        } else if (KIND == 3) {
#pragma unroll
            for (int m = 0; m < 8; m++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc2[m][q] = fma_sq(acc2[m][q], sub_bc(a[m], b2[q]));
        } else if (KIND == 4) {
#pragma unroll
            for (int m = 0; m < 8; m++)
#pragma unroll
                for (int q = 0; q < 8; q++) acc[m][q] = fmaf(a[m], b[q], acc[m][q]);
        } else if (KIND == 5) {
#pragma unroll
            for (int m = 0; m < 8; m++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc2[m][q] = fma2(pk(a[m], a[m]), b2[q], acc2[m][q]);
        } else if (KIND == 6) {
            // columns q = 0..2 (6 of 8 b-values... use 3 packed pairs) via FADD2, the last pair via min on the ALU pipe:
            // sum|a-b| = sum a + sum b - 2 sum min(a,b): accumulate sum of min (FMNMX + FADD)
#pragma unroll
            for (int m = 0; m < 8; m++) {
#pragma unroll
                for (int q = 0; q < 3; q++) acc2[m][q] = add_abs(acc2[m][q], sub_bc(a[m], b2[q]));
                acc[m][6] += fminf(a[m], b[6]);
                acc[m][7] += fminf(a[m], b[7]);
            }
        } else if (KIND == 7 || KIND == 8 || KIND == 9) {
            // packed min form: the FMNMX results of two neighbouring columns land in one register pair and are accumulated
            // by ONE FADD2 — 1 FMA-pipe lane-op + 1 ALU-pipe lane-op per element instead of 2 FMA-pipe lane-ops.
            // KIND 7: 2 of 4 column pairs via min (f = 1/2), KIND 8: 3 of 4 (f = 3/4), KIND 9: all four (f = 1)
            constexpr int NMIN = KIND == 7 ? 2 : (KIND == 8 ? 3 : 4);
#pragma unroll
            for (int m = 0; m < 8; m++) {
#pragma unroll
                for (int q = 0; q < 4 - NMIN; q++) acc2[m][q] = add_abs(acc2[m][q], sub_bc(a[m], b2[q]));
#pragma unroll
                for (int q = 4 - NMIN; q < 4; q++) {
                    const uint64_t mn = pk(fminf(a[m], b[2 * q]), fminf(a[m], b[2 * q + 1]));
                    asm("add.f32x2 %0, %0, %1;" : "+l"(acc2[m][q]) : "l"(mn));
                }
            }
        }
        // perturb the operands a little so nothing is loop-invariant (1 FADD per 64 updates for each of 16 regs is noise)
        a[it & 7] += 1.0f;
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int m = 0; m < 8; m++) for (int q = 0; q < 8; q++) s += acc[m][q];
    for (int m = 0; m < 8; m++) for (int q = 0; q < 4; q++) { float lo, hi; unpk(acc2[m][q], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char *name, int blocks_per_sm, int sms, const float *in, float *out, long long *cyc)
{
    int grid = blocks_per_sm * sms;
    bench<KIND><<<grid, 256>>>(in, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<KIND><<<grid, 256>>>(in, out, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long *h = (long long *)malloc(grid * sizeof(long long));
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; i++) mean += h[i]; mean /= grid;
    double elems = (double)ITERS * 64 * 256 * blocks_per_sm;          // element updates per SM
    printf("{\"kernel\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"mean_cycles\": %.0f, "
           "\"elem_updates_per_clk_per_sm\": %.2f, \"elem_updates_per_s\": %.4e}\n",
           name, blocks_per_sm, ms, mean, elems / mean, elems * sms / (ms * 1e-3));
    free(h);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    float *in, *out; long long *cyc;
    cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4); cudaMalloc(&cyc, sms * 8 * sizeof(long long));
    for (int bps : {1, 2, 4}) {
        run<0>("fadd", bps, sms, in, out, cyc);
        run<1>("l1_scalar", bps, sms, in, out, cyc);
        run<2>("l1_packed", bps, sms, in, out, cyc);
        run<3>("l2_packed", bps, sms, in, out, cyc);
        run<4>("ffma", bps, sms, in, out, cyc);
        run<5>("ffma2", bps, sms, in, out, cyc);
        run<6>("l1_minmix", bps, sms, in, out, cyc);
        run<7>("l1_minpacked_half", bps, sms, in, out, cyc);
        run<8>("l1_minpacked_3of4", bps, sms, in, out, cyc);
        run<9>("l1_minpacked_all", bps, sms, in, out, cyc);
    }
    return 0;
}
