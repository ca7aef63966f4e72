// Microbenchmark: issue rate of legacy mma.sync (m16n8k8 tf32, m16n8k16 bf16) on sm_100a, alone and interleaved with FFMA work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int NACC, int FFMA>
__global__ void __launch_bounds__(256) k_rate(int iters, float* out, uint32_t seed) {
    float d[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; i++) a[i] = __float_as_uint(1.0f + (threadIdx.x & 3) * 0.25f + seed);
    for (int i = 0; i < 2; i++) b[i] = __float_as_uint(0.5f + (threadIdx.x & 7) * 0.125f);
    for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) d[j][i] = 0.f;
    float f[8];
    for (int i = 0; i < 8; i++) f[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < NACC; j++) {
            if (KIND == 0) mma_tf32(d[j], a, b); else if (KIND == 1) mma_bf16(d[j], a, b);
#pragma unroll
            for (int k = 0; k < FFMA; k++) f[k & 7] = fmaf(f[k & 7], 1.0001f, 0.5f);
        }
    }
    float s = 0.f;
    for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += d[j][i];
    for (int i = 0; i < 8; i++) s += f[i];
    if (s == 123.456f) out[0] = s;
}

template <int KIND, int NACC, int FFMA>
void run(const char* name, int warps_per_sm) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int iters = 20000;
    float* out; cudaMalloc(&out, 4);
    const int ctas = p.multiProcessorCount * (warps_per_sm / 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_rate<KIND, NACC, FFMA><<<ctas, 256>>>(100, out, 0);
    cudaEventRecord(e0);
    k_rate<KIND, NACC, FFMA><<<ctas, 256>>>(iters, out, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * khz * 1e3;
    const double warp_mma_per_smsp = (double)iters * NACC * (KIND < 2 ? 1 : 0) * warps_per_sm / 4.0;
    const double slots = (double)iters * NACC * (1 + FFMA) * warps_per_sm / 4.0;
    printf("%-34s warps/SM %2d  %.3f ms  cycles/MMA/SMSP %.2f  cycles/instr/SMSP %.2f  (clock attr %d kHz)\n", name, warps_per_sm, ms,
           warp_mma_per_smsp > 0 ? cycles / warp_mma_per_smsp : 0.0, cycles / slots, khz);
    cudaFree(out);
}

int main() {
    run<0, 4, 0>("tf32 m16n8k8, 4 acc", 8);
    run<0, 4, 0>("tf32 m16n8k8, 4 acc", 16);
    run<0, 4, 0>("tf32 m16n8k8, 4 acc", 32);
    run<0, 1, 0>("tf32 m16n8k8, 1 acc (dependent)", 8);
    run<0, 2, 8>("tf32 + 8 FFMA per MMA", 16);
    run<0, 2, 16>("tf32 + 16 FFMA per MMA", 16);
    run<0, 2, 32>("tf32 + 32 FFMA per MMA", 16);
    run<1, 4, 0>("bf16 m16n8k16, 4 acc", 16);
    run<2, 2, 16>("FFMA only (16 per slot)", 16);
    return 0;
}
