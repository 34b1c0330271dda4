// scratch/thr.cu -- issue interval of FP64 instruction kinds (8 independent chains per thread).
#include <cstdio>
template <int KIND>
__global__ void k(double* out, long long* cyc, double a, double b) {
    double r[8];
    for (int i = 0; i < 8; ++i) r[i] = a + threadIdx.x * 1e-3 + i;
    int cnt = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (KIND == 0) r[u] = fma(r[u], b, a);
            if (KIND == 1) r[u] = r[u] * b;
            if (KIND == 2) r[u] = r[u] + a;
            if (KIND == 3) { r[u] = fma(r[u], b, a); r[u] = r[u] * b; }                 // DFMA + DMUL
            if (KIND == 4) { cnt += (r[u] > a) ? 1 : 0; r[u] = fma(r[u], b, a); }       // DSETP + DFMA
            if (KIND == 5) { r[u] = fma(r[u], b, a); cnt += __double2hiint(r[u]) & 1; } // DFMA + int
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = cnt;
    for (int i = 0; i < 8; ++i) s += r[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 8 * 4096); cudaMalloc(&c, 64);
    const char* names[] = {"DFMA", "DMUL", "DADD", "DFMA+DMUL", "DSETP+DFMA", "DFMA+int"};
    for (int kind = 0; kind < 6; ++kind) {
        for (int warps = 1; warps <= 4; warps *= 4) {
            if (kind == 0) k<0><<<1, 128 * warps>>>(o, c, 1e-9, 0.999999);
            if (kind == 1) k<1><<<1, 128 * warps>>>(o, c, 1e-9, 0.999999);
            if (kind == 2) k<2><<<1, 128 * warps>>>(o, c, 1e-9, 0.999999);
            if (kind == 3) k<3><<<1, 128 * warps>>>(o, c, 1e-9, 0.999999);
            if (kind == 4) k<4><<<1, 128 * warps>>>(o, c, 1e-9, 0.999999);
            if (kind == 5) k<5><<<1, 128 * warps>>>(o, c, 1e-9, 0.999999);
            long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            const int per = (kind >= 3) ? 2 : 1;
            printf("%-12s warps/SMSP %d: %.2f cycles per instruction per SMSP (%.2f per loop body item)\n", names[kind], warps,
                   (double)h / (256.0 * 8 * per * warps), (double)h / (256.0 * 8 * warps));
        }
    }
    return 0;
}
