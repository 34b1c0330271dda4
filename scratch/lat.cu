// scratch/lat.cu -- dependent-issue latency of FP64 instructions on one warp (clock64 around chains).
#include <cstdio>
__global__ void k(double* out, long long* cyc, double a, double b) {
    double x = a + threadIdx.x * 1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = fma(x, b, a);
    }
    long long t1 = clock64();
    double y = x;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) y = y + b;
    }
    long long t2 = clock64();
    double z = y;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) z = __drcp_rn(z) + a;
    }
    long long t3 = clock64();
    // two independent chains
    double p = z, q = z + 1;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { p = fma(p, b, a); q = fma(q, b, a); }
    }
    long long t4 = clock64();
    double r0 = p, r1 = q, r2 = p + 1, r3 = q + 2, r4 = p + 3, r5 = q + 4, r6 = p + 5, r7 = q + 6;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            r0 = fma(r0, b, a); r1 = fma(r1, b, a); r2 = fma(r2, b, a); r3 = fma(r3, b, a);
            r4 = fma(r4, b, a); r5 = fma(r5, b, a); r6 = fma(r6, b, a); r7 = fma(r7, b, a);
        }
    }
    long long t5 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
    }
    out[threadIdx.x + blockIdx.x * blockDim.x] = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 8 * 4096); cudaMalloc(&c, 64);
    for (int warps = 1; warps <= 8; warps *= 2) {
        k<<<1, 32 * warps * 4>>>(o, c, 1e-9, 0.999999);   // warps per SMSP = `warps`
        long long h[5]; cudaMemcpy(h, c, 40, cudaMemcpyDeviceToHost);
        printf("warps/SMSP %d: DFMA dep %.1f cyc, DADD dep %.1f, drcp+DADD %.1f, 2 chains %.1f per DFMA-pair, 8 chains %.1f per 8 DFMA\n", warps,
               h[0] / 1024.0, h[1] / 1024.0, h[2] / 256.0, h[3] / 1024.0, h[4] / 1024.0);
    }
    return 0;
}
