// scratch/panel_bench.cu -- phase timing of panel_sym_kernel (clock64 at phase boundaries of CTA 0).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DEMME_PS_CLOCKS -o scratch/panel_bench scratch/panel_bench.cu emme_b200/csrc/peer.cu
#include <cstdio>
#include <vector>
#include "../emme_b200/csrc/dense.cu"

int main(int argc, char** argv) {
    using namespace emme;
    const int dim = argc > 1 ? atoi(argv[1]) : 1024;
    std::vector<double2> h((size_t)dim * dim);
    for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) {
            const int a = i < j ? i : j, b = i < j ? j : i;
            h[(size_t)i * dim + j] = i == j ? make_double2(2.0, 0.1) : make_double2(0.3 * sin(0.37 * a + 0.11 * b) / sqrt((double)dim), 0.2 * cos(0.23 * a - 0.07 * b) / sqrt((double)dim));
        }
    double2 *W, *Y, *dvec;
    int *flag, *info;
    cudaMalloc(&dvec, sizeof(double2) * dim);
    cudaMalloc(&W, sizeof(double2) * dim * dim);
    cudaMalloc(&Y, sizeof(double2) * dim * dim);
    cudaMalloc(&flag, 4);
    cudaMalloc(&info, 4);
    cudaMemcpy(W, h.data(), sizeof(double2) * dim * dim, cudaMemcpyHostToDevice);
    cudaMemset(Y, 0, sizeof(double2) * dim * dim);
    cudaMemset(flag, 0, 4);
    cudaMemset(info, 0, 4);
    gemm_setup();
    const int k0 = 0, jb = 32, ke = 32;
    const int tile = dim <= 4736 ? 16 : (dim <= 9472 ? 32 : 64);
    const int n_row = (dim - ke + tile - 1) / tile, n_col = (ke + tile - 1) / tile;
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(W, h.data(), sizeof(double2) * dim * dim, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        panel_sym_kernel<<<n_row + n_col, 128, PS_SMEM_BYTES>>>(W, Y, dim, dim, k0, jb, 0, ke, n_row, tile, 4.0, flag, dvec);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long c[8];
        cudaMemcpyFromSymbol(c, g_ps_clocks, sizeof c);
        printf("rep %d: %.1f us; cycles: load %lld, factor %lld (%.0f/pivot), wait %lld, products+stores %lld; err=%s\n", rep, ms * 1e3,
               c[1] - c[0], c[2] - c[1], (double)(c[2] - c[1]) / jb, c[3] - c[2], c[4] - c[3], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
