// Probe: does mma.sync m16n8k16 f16 (fp32 accumulate) consume fp16 DENORMAL A-operands exactly?
// A[m][k] = denormal half with bits (c << s), B[k][n] = arbitrary fp16; compare against fp64.
#include <cstdio>
#include <cuda_fp16.h>
#include <cstdint>
#include <cmath>
__global__ void k(const uint16_t* A, const uint16_t* B, float* C) {
    // A row-major 16x16, B col-major (k x n) stored as B[n][k], C row-major 16x8
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    auto ld2 = [&](const uint16_t* p, int i0, int i1) { return (uint32_t)p[i0] | ((uint32_t)p[i1] << 16); };
    uint32_t a0 = ld2(A, g * 16 + 2 * t, g * 16 + 2 * t + 1);
    uint32_t a1 = ld2(A, (g + 8) * 16 + 2 * t, (g + 8) * 16 + 2 * t + 1);
    uint32_t a2 = ld2(A, g * 16 + 2 * t + 8, g * 16 + 2 * t + 9);
    uint32_t a3 = ld2(A, (g + 8) * 16 + 2 * t + 8, (g + 8) * 16 + 2 * t + 9);
    uint32_t b0 = ld2(B, g * 16 + 2 * t, g * 16 + 2 * t + 1);
    uint32_t b1 = ld2(B, g * 16 + 2 * t + 8, g * 16 + 2 * t + 9);
    float c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1; C[(g + 8) * 8 + 2 * t] = c2; C[(g + 8) * 8 + 2 * t + 1] = c3;
}
int run(bool denorm) {
    uint16_t hA[256], hB[128]; float hC[128];
    double worst = 0; int bad = 0;
    for (int shift = 0; shift <= 8; shift += 2) {
        srand(shift + 1);
        for (int i = 0; i < 256; ++i) { int c = rand() & 3; if (denorm) hA[i] = (uint16_t)(c << shift); else { __half h = __float2half((float)(c << shift)); hA[i] = *(uint16_t*)&h; } }
        for (int i = 0; i < 128; ++i) { __half h = __float2half((rand() / (float)RAND_MAX - 0.5f) * 8.f); hB[i] = *(uint16_t*)&h; }
        uint16_t *dA, *dB; float* dC;
        cudaMalloc(&dA, 512); cudaMalloc(&dB, 256); cudaMalloc(&dC, 512);
        cudaMemcpy(dA, hA, 512, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
        k<<<1, 32>>>(dA, dB, dC); cudaMemcpy(hC, dC, 512, cudaMemcpyDeviceToHost);
        for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) {
            double ref = 0, l1 = 0;
            for (int kk = 0; kk < 16; ++kk) {
                double av = denorm ? (double)hA[m * 16 + kk] * ldexp(1.0, -24) : (double)__half2float(*(__half*)&hA[m * 16 + kk]);
                double t = av * (double)__half2float(*(__half*)&hB[n * 16 + kk]);
                ref += t; l1 += fabs(t);
            }
            double err = fabs(hC[m * 8 + n] - ref), rel = err / (l1 + 1e-300);      // error relative to the L1 mass
            if (rel > 2e-7) ++bad;
            if (rel > worst) worst = rel;
        }
        printf("shift %d: sample C[3][2]=%g\n", shift, hC[3 * 8 + 2]);
    }
    printf("mma probe denorm=%d: bad=%d worst_rel_to_L1=%.3e\n", (int)denorm, bad, worst);
    return 0;
}
int main() { run(true); run(false); return 0; }
