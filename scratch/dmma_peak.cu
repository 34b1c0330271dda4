// micro-benchmark: FP64 tensor-core (DMMA) throughput via mma.sync on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) dmma884(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dmma16816(double* out, int iters) {
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
    double c[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> double timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best * 1e-3;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, sizeof(double) * sms * 8 * 256);
    const int iters = 20000;
    for (int wpb : {4, 8}) {
        int blocks = sms * (16 / wpb) ;
        double t = timeit([&] { dmma884<<<blocks, wpb * 32>>>(d, iters); });
        double fl = 2.0 * 8 * 8 * 4 * 8 * (double)iters * blocks * wpb;
        printf("m8n8k4   warps/SM %d: %.2f TFLOP/s  (%s)\n", blocks * wpb / sms, fl / t / 1e12, cudaGetErrorString(cudaGetLastError()));
        t = timeit([&] { dmma16816<<<blocks, wpb * 32>>>(d, iters); });
        fl = 2.0 * 16 * 8 * 16 * 4 * (double)iters * blocks * wpb;
        printf("m16n8k16 warps/SM %d: %.2f TFLOP/s  (%s)\n", blocks * wpb / sms, fl / t / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
