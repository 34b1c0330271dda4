// peer.cu -- stream-ordered signalling kernels over peer-mapped flag pages (see peer.h).
#include "peer.h"

namespace emme {

static unsigned long long g_timeout_ns = 20ull * 1000000000ull;
unsigned long long peer_timeout_ns() { return g_timeout_ns; }
void peer_set_timeout(double seconds) {
    g_timeout_ns = seconds <= 0 ? 1000000ull : (unsigned long long)(seconds * 1e9);
}

__global__ void __launch_bounds__(32)
peer_barrier_kernel(const PeerFlags f, unsigned long long epoch, unsigned long long timeout_ns) {
    const int r = threadIdx.x;
    if (r >= f.n) return;
    // everything this stream did before the barrier (including stores into peers' buffers by the
    // preceding kernels) is ordered before the announcement
    __threadfence_system();
    st_release_sys(f.p[r] + PEER_W_BARRIER + f.me, epoch);
    if (!spin_until(f.p[f.me] + PEER_W_BARRIER + r, epoch, timeout_ns))
        atomicMax(f.p[f.me] + PEER_W_ERROR, (unsigned long long)(PEER_W_BARRIER + r + 1));
}

__global__ void __launch_bounds__(32)
peer_signal_kernel(const PeerFlags f, int word, unsigned long long value, unsigned mask) {
    const int r = threadIdx.x;
    if (r >= f.n || r == f.me || !((mask >> r) & 1u)) return;
    __threadfence_system();
    st_release_sys(f.p[r] + word, value);
}

__global__ void __launch_bounds__(32)
peer_wait_kernel(const PeerFlags f, int word, unsigned long long value, unsigned long long timeout_ns) {
    if (threadIdx.x != 0) return;
    if (!spin_until(f.p[f.me] + word, value, timeout_ns))
        atomicMax(f.p[f.me] + PEER_W_ERROR, (unsigned long long)(word + 1));
}

__global__ void __launch_bounds__(32)
peer_post_kernel(const PeerFlags f, unsigned long long serial, const double2* __restrict__ v,
                 const int* __restrict__ flag, const int* __restrict__ info) {
    const int r = threadIdx.x;
    if (r >= f.n) return;
    unsigned long long* slot = f.p[r] + PEER_W_MAILBOX + 4 * f.me;
    const double2 t = *v;
    slot[1] = (unsigned long long)__double_as_longlong(t.x);
    slot[2] = (unsigned long long)__double_as_longlong(t.y);
    slot[3] = ((unsigned long long)(unsigned)*info << 32) | (unsigned)*flag;
    __threadfence_system();
    st_release_sys(slot, serial);
}

__global__ void __launch_bounds__(32)
peer_collect_kernel(const PeerFlags f, unsigned long long serial, double2* __restrict__ v, int* __restrict__ flag,
                    int* __restrict__ info, unsigned long long timeout_ns) {
    if (threadIdx.x != 0) return;
    double x = 0., y = 0.;
    unsigned fl = 0, inf = 0;
    for (int r = 0; r < f.n; ++r) {
        const unsigned long long* slot = f.p[f.me] + PEER_W_MAILBOX + 4 * r;
        if (!spin_until(slot, serial, timeout_ns)) {
            atomicMax(f.p[f.me] + PEER_W_ERROR, (unsigned long long)(PEER_W_MAILBOX + 4 * r + 1));
            fl |= 1u;
            continue;
        }
        x += __longlong_as_double((long long)slot[1]);     // fixed rank order: every rank forms the same sum
        y += __longlong_as_double((long long)slot[2]);
        fl |= (unsigned)(slot[3] & 0xffffffffull);
        if (inf == 0) inf = (unsigned)(slot[3] >> 32);
    }
    *v = make_double2(x, y);
    *flag = (int)fl;
    *info = (int)inf;
}

cudaError_t launch_peer_barrier(const PeerFlags& f, unsigned long long epoch, cudaStream_t stream) {
    peer_barrier_kernel<<<1, 32, 0, stream>>>(f, epoch, g_timeout_ns);
    return cudaGetLastError();
}
cudaError_t launch_peer_signal(const PeerFlags& f, int word, unsigned long long value, cudaStream_t stream,
                               unsigned mask) {
    peer_signal_kernel<<<1, 32, 0, stream>>>(f, word, value, mask);
    return cudaGetLastError();
}
cudaError_t launch_peer_wait(const PeerFlags& f, int word, unsigned long long value, cudaStream_t stream) {
    peer_wait_kernel<<<1, 32, 0, stream>>>(f, word, value, g_timeout_ns);
    return cudaGetLastError();
}
cudaError_t launch_peer_post(const PeerFlags& f, unsigned long long serial, const double2* d_value,
                             const int* d_flag, const int* d_info, cudaStream_t stream) {
    peer_post_kernel<<<1, 32, 0, stream>>>(f, serial, d_value, d_flag, d_info);
    return cudaGetLastError();
}
cudaError_t launch_peer_collect(const PeerFlags& f, unsigned long long serial, double2* d_value, int* d_flag,
                                int* d_info, cudaStream_t stream) {
    peer_collect_kernel<<<1, 32, 0, stream>>>(f, serial, d_value, d_flag, d_info, g_timeout_ns);
    return cudaGetLastError();
}

cudaError_t peer_preload() {
    cudaFuncAttributes a;
    const void* ks[] = {(const void*)peer_barrier_kernel, (const void*)peer_signal_kernel, (const void*)peer_wait_kernel,
                        (const void*)peer_post_kernel, (const void*)peer_collect_kernel};
    for (const void* k : ks) {
        cudaError_t e = cudaFuncGetAttributes(&a, k);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace emme
