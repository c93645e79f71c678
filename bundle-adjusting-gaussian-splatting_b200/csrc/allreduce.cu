// In-place all-reduce(SUM) of the flat gradient buffer over NVLink peer memory: the one exchange
// step of view-parallel training, one process per GPU of one NVSwitch node.
//
// The reference is single-GPU and has no collective; view-parallel training sums 59 floats per Gaussian over the ranks
// once per step (236 MB at 1 M Gaussians).  NCCL's ring does that in 0.64 ms on 8 B200s (645 GB/s bus bandwidth; NVLS
// does not initialise on this pool).  With every GPU one NVSwitch hop from every other, the exchange needs no ring and
// no second phase:
//
//   rank r owns elements [r n / G, (r + 1) n / G).  For each element it owns it LOADS the G contributions straight from
//   the G gradient buffers (its own + G-1 peers, mapped through CUDA IPC), adds them in rank order, and STORES the sum
//   into all G buffers.  Nobody else reads or writes that element, so no synchronisation is needed between the load and
//   the store; inbound (loads) and outbound (stores) NVLink traffic run concurrently in the same loop, each
//   (G-1)/G x n x 4 bytes per GPU -- what a ring moves, without its G-1 dependent steps.
//
// Two cross-GPU barriers bracket the loop, built from 32-bit epoch flags in peer memory: "every rank's gradients are
// final" before the first load, "every rank has finished storing" before the kernel retires.  All ranks add in the same
// order, so the buffers end up bit-identical on every rank.  The sums differ from NCCL's by fp32 summation order only.
#include "../../include/bags_b200.h"
#include "common.cuh"

#include <string.h>

namespace bags {

constexpr int AR_MAX_RANKS = 8;
constexpr int AR_THREADS = 512;
// flag words of a rank (u32): [0..7] arrival epoch written by rank p, [8..15] completion epoch written by rank p,
// [16] epoch of the last finished all-reduce, [17] CTAs of the current one that are done
constexpr int AR_ARRIVE = 0, AR_DONE = 8, AR_EPOCH = 16, AR_CTAS = 17;

struct P2PArgs {
    float* buf[AR_MAX_RANKS];
    u32* flag[AR_MAX_RANKS];
    int world, rank;
};

__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u32* p, u32 v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float4* p, const float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int G>
__global__ void __launch_bounds__(AR_THREADS)
allreduce_p2p_kernel(const P2PArgs a, const long long n4) {
    u32* my = a.flag[a.rank];
    __shared__ u32 s_epoch;
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_epoch = ld_acquire_sys(my + AR_EPOCH) + 1u;     // advanced by the last CTA of the previous call
    __syncthreads();
    const u32 epoch = s_epoch;
    // ---- barrier 1: every rank's buffer is final (this kernel runs behind the rank's backward in stream order) --------
    if (blockIdx.x == 0 && threadIdx.x < G) st_release_sys(a.flag[threadIdx.x] + AR_ARRIVE + a.rank, epoch);
    if (threadIdx.x < G)
        while ((int)(ld_acquire_sys(my + AR_ARRIVE + threadIdx.x) - epoch) < 0) {}
    __syncthreads();

    // ---- the owned slice: load G contributions, add in rank order, store to all G buffers ----------------------------
    const long long per = (n4 + G - 1) / G;
    const long long lo = per * a.rank;
    const long long hi = lo + per < n4 ? lo + per : n4;
    const long long stride = (long long)gridDim.x * AR_THREADS;
    for (long long i = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i < hi; i += stride) {
        float4 v[G];
#pragma unroll
        for (int p = 0; p < G; p++) v[p] = ld_peer(reinterpret_cast<const float4*>(a.buf[p]) + i);
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < G; p++) {
            s.x = __fadd_rn(s.x, v[p].x); s.y = __fadd_rn(s.y, v[p].y);
            s.z = __fadd_rn(s.z, v[p].z); s.w = __fadd_rn(s.w, v[p].w);
        }
#pragma unroll
        for (int p = 0; p < G; p++) st_peer(reinterpret_cast<float4*>(a.buf[p]) + i, s);
    }

    // ---- barrier 2: every rank has finished storing into every buffer ---------------------------------------------------
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(my + AR_CTAS, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x < G) st_release_sys(a.flag[threadIdx.x] + AR_DONE + a.rank, epoch);
    if (threadIdx.x < G)
        while ((int)(ld_acquire_sys(my + AR_DONE + threadIdx.x) - epoch) < 0) {}
    __syncthreads();
    if (threadIdx.x == 0) {
        my[AR_CTAS] = 0u;
        st_release_sys(my + AR_EPOCH, epoch);
    }
}

}  // namespace bags

using namespace bags;

extern "C" size_t bags_allreduce_p2p_flag_bytes(void) { return 256; }

extern "C" int bags_enable_peer_access(int peer_device) {
    int dev = 0, can = 0;
    BAGS_CUDA_OK(cudaGetDevice(&dev));
    if (peer_device == dev) return 0;
    BAGS_CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) {
        set_error("bags_enable_peer_access: device %d cannot access the memory of device %d", dev, peer_device);
        return BAGS_E_INVALID;
    }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        return 0;
    }
    BAGS_CUDA_OK(e);
    return 0;
}

// Maps a peer process's device allocation (its 64-byte cudaIpcMemHandle_t) into THIS process under the CURRENT device,
// with peer access set up by the driver (cudaIpcMemLazyEnablePeerAccess): the returned pointer is valid in kernels of the
// current device.  A handle can be opened once per process and device.
extern "C" int bags_ipc_open(const void* handle64, void** out_ptr) {
    if (!handle64 || !out_ptr) {
        set_error("bags_ipc_open: null argument");
        return BAGS_E_INVALID;
    }
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    BAGS_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out_ptr = p;
    return 0;
}
extern "C" int bags_ipc_close(void* ptr) {
    if (ptr) BAGS_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int bags_allreduce_p2p(int world, int rank, float* const* buffers, uint32_t* const* flags, long long n,
                                  void* stream) {
    if (world < 1 || world > AR_MAX_RANKS || rank < 0 || rank >= world || !buffers || !flags || n < 0) {
        set_error("bags_allreduce_p2p: invalid argument (1..%d ranks)", AR_MAX_RANKS);
        return BAGS_E_INVALID;
    }
    if (n % 4 != 0) {
        set_error("bags_allreduce_p2p: the element count must be a multiple of 4 (got %lld)", n);
        return BAGS_E_INVALID;
    }
    P2PArgs a;
    a.world = world;
    a.rank = rank;
    for (int p = 0; p < AR_MAX_RANKS; p++) {
        a.buf[p] = p < world ? buffers[p] : nullptr;
        a.flag[p] = p < world ? flags[p] : nullptr;
        if (p < world && (!a.buf[p] || !a.flag[p] || ((size_t)a.buf[p] & 15))) {
            set_error("bags_allreduce_p2p: rank %d: null or misaligned buffer", p);
            return BAGS_E_INVALID;
        }
    }
    if (world == 1 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    BAGS_CUDA_OK(cudaGetDevice(&dev));
    BAGS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long n4 = n / 4;
    const long long per = (n4 + world - 1) / world;
    // every CTA must be resident while it waits in the barriers: at most one per SM
    int grid = (int)((per + AR_THREADS - 1) / AR_THREADS);
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    switch (world) {
        case 2: allreduce_p2p_kernel<2><<<grid, AR_THREADS, 0, st>>>(a, n4); break;
        case 3: allreduce_p2p_kernel<3><<<grid, AR_THREADS, 0, st>>>(a, n4); break;
        case 4: allreduce_p2p_kernel<4><<<grid, AR_THREADS, 0, st>>>(a, n4); break;
        case 5: allreduce_p2p_kernel<5><<<grid, AR_THREADS, 0, st>>>(a, n4); break;
        case 6: allreduce_p2p_kernel<6><<<grid, AR_THREADS, 0, st>>>(a, n4); break;
        case 7: allreduce_p2p_kernel<7><<<grid, AR_THREADS, 0, st>>>(a, n4); break;
        default: allreduce_p2p_kernel<8><<<grid, AR_THREADS, 0, st>>>(a, n4); break;
    }
    BAGS_LAUNCH_OK("allreduce_p2p_kernel");
    return 0;
}
