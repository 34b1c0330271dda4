// peer.h -- device-side signalling between the GPUs of one NVSwitch box (peer.cu).
//
// Every rank owns one small "flag page" in device memory (cudaMalloc, exported with CUDA IPC or --
// virtual ranks inside one process -- passed as a plain pointer) and maps the pages of all peers.
// A rank signals by storing a monotonically increasing 64-bit value into a word of a PEER's page
// (st.release.sys over NVLink, ordered after the data it has stored into that peer's buffers) and
// waits by polling its OWN page (ld.acquire.sys, local memory).  No host round trip, no NCCL: the
// barriers and panel-ready flags are ordinary kernels in the handle's stream.  Values never
// decrease, so a page is zeroed once, at creation, and never reset.
#pragma once
#include <cuda_runtime.h>

#include "assembly.h"   // EMME_MAX_PEERS

namespace emme {

// word offsets inside a flag page (unsigned long long units)
constexpr int PEER_W_BARRIER = 0;      // [8]  word r: last barrier epoch announced by rank r
constexpr int PEER_W_ERROR = 8;        //      local only: a wait timed out (value = word waited for + 1)
constexpr int PEER_W_MAILBOX = 16;     // [8][4] rank r -> {serial, payload0, payload1, payload2}
constexpr int PEER_W_PANEL = 64;       // [PEER_MAX_PANELS] panel K of dense step `serial` has arrived
constexpr int PEER_MAX_PANELS = 2048;
constexpr int PEER_W_USER = PEER_W_PANEL + PEER_MAX_PANELS;   // [..] free for other protocols (PIC)
constexpr int PEER_PAGE_WORDS = 8192;  // 64 KB

struct PeerFlags {
    unsigned long long* p[EMME_MAX_PEERS];   // p[r]: flag page of rank r (p[me] is local memory)
    int n, me;
};

// all ranks arrive at `epoch` (every rank passes the same, increasing epoch): 1 CTA, 32 threads
cudaError_t launch_peer_barrier(const PeerFlags& f, unsigned long long epoch, cudaStream_t stream);
// store `value` into word `word` of the page of every OTHER rank in `mask` (bit r = rank r; ordered
// after this stream's earlier kernels)
cudaError_t launch_peer_signal(const PeerFlags& f, int word, unsigned long long value, cudaStream_t stream,
                               unsigned mask = 0xffffffffu);
// block the stream until word `word` of the local page is >= value (bounded: sets PEER_W_ERROR on timeout)
cudaError_t launch_peer_wait(const PeerFlags& f, int word, unsigned long long value, cudaStream_t stream);
// mailbox: {serial, a, b, c} into slot `me` of EVERY rank's page (own included)
cudaError_t launch_peer_post(const PeerFlags& f, unsigned long long serial, const double2* d_value,
                             const int* d_flag, const int* d_info, cudaStream_t stream);
// after a barrier: sum the posted values in rank order (identical on every rank), OR the flags, first info
cudaError_t launch_peer_collect(const PeerFlags& f, unsigned long long serial, double2* d_value, int* d_flag,
                                int* d_info, cudaStream_t stream);
// load the kernels above now (see dense_preload)
cudaError_t peer_preload();
// seconds a wait may spin before it gives up (default 20; tests shorten it)
void peer_set_timeout(double seconds);
unsigned long long peer_timeout_ns();

// device helpers shared with kernels that signal from their own epilogue
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// poll *p until it is >= value; false on timeout
__device__ __forceinline__ bool spin_until(const unsigned long long* p, unsigned long long value,
                                           unsigned long long timeout_ns) {
    if (ld_acquire_sys(p) >= value) return true;
    const unsigned long long t0 = global_ns();
    unsigned ns = 32;
    for (;;) {
        if (ld_acquire_sys(p) >= value) return true;
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
        if (global_ns() - t0 > timeout_ns) return false;
    }
}

}  // namespace emme
