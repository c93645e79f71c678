// Tile binning for sm_100a: hand-written onesweep radix sort, chained-scan tile-instance emission, tile ranges.
//
// The reference bins with duplicateWithKeys -> cub::DeviceRadixSort::SortPairs on 64-bit (tile << 32 | depth) keys
// over 32+bit key bits -> identifyTileRanges (3dgs-pose/cuda_rasterizer/rasterizer_impl.cu:70-138,287-327).
// A stable LSD sort of that key over instances emitted in (Gaussian id, tile) order yields the order
// (tile, depth bits, Gaussian id).  The same order is produced here with far less traffic by a two-level sort:
//
//   1. depth sort   : stable sort of the P Gaussians by their 32 depth bits (culled ones carry 0xFFFFFFFF),
//                     4 x 8-bit onesweep passes over P x 8 B  -> order (depth, id);
//   2. emit         : in that order, append (tile, id) for every tile of the stored rectangle (y outer, x inner
//                     like rasterizer_impl.cu:98-108): a two-launch scan of the tile counts, then a
//                     slot-parallel fill (one thread per output instance);
//   3. tile sort    : stable sort of the R instances by tile id only, ceil(bit/8) (= 2) passes over R x 8 B
//                     -> order (tile, depth, id), identical to the reference's point_list;
//   4. ranges       : boundaries of the sorted tile ids; empty tiles keep (0,0) like the reference's cudaMemset.
//
// Per tile instance this moves 4 (histogram) + 2 x 16 (passes) + 8 (emit) bytes instead of the reference's
// 12 (emit) + 8 + 6 x 24 (sort) + 8 (ranges).
#include "../../include/bags_b200.h"
#include "common.cuh"

#include <atomic>

namespace bags {

// =================================================================================================================
// Onesweep radix sort (u32 keys, u32 values), 8-bit digits, one CTA per 4096-key tile (4 CTAs per SM), decoupled
// look-back with several predecessors in flight per round.
// =================================================================================================================
constexpr size_t sort_smem_bytes(int items) {
    return (size_t)(2 * items * SORT_THREADS + SORT_WARPS * RADIX + 2 * RADIX + 8) * sizeof(u32);
}
constexpr int SORT_LOOKBACK_FIRST = 8;   // predecessor status words fetched in the first look-back round
static int g_sort_small_tiles = 1;       // 2048-key tiles for passes that do not fill the GPU twice (see launch_sort_pairs)
static int g_sort_hwmatch = 0;           // MATCH.ANY ranking in those small-tile passes
static int g_sort_lookback = 4;          // ... and in every later round (4 or 8; see bags_set_tuning)
constexpr u32 FLAG_AGG = 1u << 30;   // tile aggregate available
constexpr u32 FLAG_INCL = 2u << 30;  // inclusive prefix available
constexpr u32 VAL_MASK = (1u << 30) - 1;

__device__ __forceinline__ u32 ld_volatile_u32(const u32* p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(u32* p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Element count of a launch.  Synchronous callers know it on the host (n_dev == NULL: n_cap is the count).  The
// asynchronous forward (bags_rasterize_forward_async) sizes its grids by the caller's capacity and every kernel reads the
// actual count -- written by the preprocess kernel -- from device memory; a count above the capacity is clamped (the
// caller sees the overflow in the count it gets back and must re-run with a larger capacity).
__device__ __forceinline__ int resolve_n(const int n_cap, const unsigned long long* __restrict__ n_dev) {
    if (n_dev == nullptr) return n_cap;
    const unsigned long long v = *n_dev;
    return v < (unsigned long long)n_cap ? (int)v : n_cap;
}

// Histograms of every digit position in one read of the keys.
__global__ void __launch_bounds__(512)
sort_hist_kernel(const u32* __restrict__ keys, int n_cap, const unsigned long long* __restrict__ n_dev, int begin_bit,
                 int end_bit, u32* __restrict__ ghist, const size_t data_stride, const size_t count_stride) {
    // batched launch: blockIdx.y = view (strides are 0 for a single view)
    keys = at_view(keys, data_stride, blockIdx.y);
    ghist = at_view(ghist, data_stride, blockIdx.y);
    if (n_dev != nullptr) n_dev = at_view(n_dev, count_stride, blockIdx.y);
    const int n = resolve_n(n_cap, n_dev);
    __shared__ u32 sh[SORT_MAX_PASSES][RADIX];
    for (int i = threadIdx.x; i < SORT_MAX_PASSES * RADIX; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int npass = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u32 k = keys[i];
        for (int p = 0; p < npass; p++) {
            int sft = begin_bit + p * RADIX_BITS;
            int bits = min(RADIX_BITS, end_bit - sft);
            u32 d = (k >> sft) & ((1u << bits) - 1u);
            atomicAdd(&sh[p][d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * RADIX; i += blockDim.x) {
        u32 v = (&sh[0][0])[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

// Lanes of the warp holding the same digit as the caller: one ballot per digit bit (the match.any instruction has a
// far lower throughput than `bits` votes).  Four instructions per bit: bit test into a predicate, vote, select of the
// polarity mask, one three-input logic op.
template <int B>
__device__ __forceinline__ void match_bit(u32& peers, const u32 d) {
    // peers &= set ? vote : ~vote, spelled out so that ptxas keeps it at four instructions per bit
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, v, f;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
        "selp.b32 f, 0, 0xffffffff, p;\n\t"
        "lop3.b32 %0, %0, v, f, 0x60;\n\t"
        "}"
        : "+r"(peers) : "r"(d), "n"(1u << B));
}
template <int BITS>
__device__ __forceinline__ u32 match_digit_fixed(const u32 d) {
    u32 peers = 0xffffffffu;
    match_bit<0>(peers, d);
    if (BITS > 1) match_bit<1>(peers, d);
    if (BITS > 2) match_bit<2>(peers, d);
    if (BITS > 3) match_bit<3>(peers, d);
    if (BITS > 4) match_bit<4>(peers, d);
    if (BITS > 5) match_bit<5>(peers, d);
    if (BITS > 6) match_bit<6>(peers, d);
    if (BITS > 7) match_bit<7>(peers, d);
    return peers;
}

// Exclusive scan of 256 values held one per thread by the 256 threads of the CTA (8 warps).
__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32* warp_tot /*[8] smem*/, int tid) {
    const int lane = tid & 31, w = tid >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    asm volatile("bar.sync 1, 256;");
    u32 base = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (k < w) base += warp_tot[k];
    asm volatile("bar.sync 1, 256;");
    return base + inc - v;
}

// One look-back round for digit `tid`: K predecessor status words are loaded at once (independent loads), then
// consumed nearest first; a word that is not published yet is polled.  Returns true once an inclusive prefix (or
// the start of the array) has been reached.
template <int K>
__device__ __forceinline__ bool lookback_round(const u32* __restrict__ status, const int tid, int& look, u32& excl) {
    u32 sw[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int t = look - k;
        sw[k] = (t >= 0) ? ld_volatile_u32(status + (size_t)t * RADIX + tid) : FLAG_INCL;
    }
    bool done = false;
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (done) continue;
        u32 v = sw[k];
        while ((v & ~VAL_MASK) == 0) v = ld_volatile_u32(status + (size_t)(look - k) * RADIX + tid);
        excl += v & VAL_MASK;
        if ((v & ~VAL_MASK) == FLAG_INCL) done = true;
    }
    look -= K;
    return done;
}

// Digit of a key.  BYTE: the digit is one whole byte of the key (8 bits, shift a multiple of 8): one PRMT.
template <bool BYTE>
__device__ __forceinline__ u32 digit_of(const u32 key, const int shift, const u32 mask, const u32 sel) {
    if (BYTE) return __byte_perm(key, 0u, sel);
    return (key >> shift) & mask;
}

// Rank of every key of the thread inside its warp's digit groups, in memory order.
// Per item: the lanes holding the same digit (one ballot per digit bit), the number of them below this lane, and one
// shared-memory atomic by the leader lane that reserves the group's slots in the warp's digit counter.  The atomics
// of one warp reach the shared-memory unit in program order, so equal digits of later items see the counts of earlier
// ones without a warp barrier per item, and groups of items pipeline instead of forming a chain of 16
// load-add-store-shuffle round trips.  Padding keys (0xFFFFFFFF, only past the end of the last tile) fall into the
// highest digit after every real key, so they need no special case.
template <int BITS, bool BYTE, int ITEMS, bool HWMATCH>
__device__ __forceinline__ void rank_items(const u32 (&key)[ITEMS], u32 (&rank)[ITEMS], u32* wc,
                                           const int shift, const u32 mask, const u32 sel, const int lane) {
    const u32 lt_mask = (1u << lane) - 1u;
    const u32 wc_addr = (u32)__cvta_generic_to_shared(wc);
    constexpr int RANK_GROUP = 4;   // items in flight between the atomics and the shuffles that distribute them
#pragma unroll
    for (int i0 = 0; i0 < ITEMS; i0 += RANK_GROUP) {
        u32 old[RANK_GROUP];
#pragma unroll
        for (int k = 0; k < RANK_GROUP; k++) {
            const int i = i0 + k;
            const u32 d = digit_of<BYTE>(key[i], shift, mask, sel);
            // HWMATCH: one MATCH.ANY instead of BITS votes -- far lower throughput, but a single instruction with no
            // predicate registers, so the items of a group overlap; only offered to the latency-bound small tiles
            const u32 peers = HWMATCH ? __match_any_sync(0xffffffffu, d) : match_digit_fixed<BITS>(d);
            const u32 lower = __popc(peers & lt_mask);
            const u32 cnt = __popc(peers);
            old[k] = 0;
            // lowest lane of the group leads (predicated atomic: no divergent region)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t@p atom.shared.add.u32 %0, [%1], %3;\n\t}"
                         : "+r"(old[k]) : "r"(wc_addr + d * 4u), "r"(lower), "r"(cnt) : "memory");
            rank[i] = lower | ((u32)(__ffs(peers) - 1) << 16);
        }
#pragma unroll
        for (int k = 0; k < RANK_GROUP; k++) {
            const int i = i0 + k;
            rank[i] = __shfl_sync(0xffffffffu, old[k], (int)(rank[i] >> 16)) + (rank[i] & 0xffffu);
        }
    }
}

template <bool IOTA, int LOOKBACK, bool BYTE, int ITEMS, bool HWMATCH>
__global__ void __launch_bounds__(SORT_THREADS, 4)
onesweep_pass_kernel(const u32* __restrict__ keys_in, const u32* __restrict__ vals_in, u32* __restrict__ keys_out,
                     u32* __restrict__ vals_out, int n_cap, const unsigned long long* __restrict__ n_dev, int shift,
                     int bits, const u32* __restrict__ ghist, u32* __restrict__ tile_counter,
                     u32* __restrict__ status, const size_t data_stride, const size_t count_stride) {
    constexpr int TILE = ITEMS * SORT_THREADS;
    if (data_stride != 0) {   // batched launch: blockIdx.y = view
        keys_in = at_view(keys_in, data_stride, blockIdx.y);
        vals_in = at_view(vals_in, data_stride, blockIdx.y);
        keys_out = at_view(keys_out, data_stride, blockIdx.y);
        vals_out = at_view(vals_out, data_stride, blockIdx.y);
        ghist = at_view(ghist, data_stride, blockIdx.y);
        tile_counter = at_view(tile_counter, data_stride, blockIdx.y);
        status = at_view(status, data_stride, blockIdx.y);
        if (n_dev != nullptr) n_dev = at_view(n_dev, count_stride, blockIdx.y);
    }
    extern __shared__ u32 sort_smem[];
    u32* s_keys = sort_smem;                                             // [TILE] keys in locally sorted order
    u32* s_vals = s_keys + TILE;                                    // [TILE] their values
    u32 (*s_wcnt)[RADIX] = reinterpret_cast<u32 (*)[RADIX]>(s_vals + TILE);  // [SORT_WARPS][RADIX]
    u32* s_gbase = s_vals + TILE + SORT_WARPS * RADIX;              // [RADIX] global index of local position 0
    u32* s_wtot = s_gbase + 2 * RADIX;                                   // [8]
    __shared__ u32 s_tile;
    static_assert(SORT_THREADS == RADIX, "one thread per digit in the look-back");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s_wcnt[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const long long tile_base = (long long)tile * TILE;
    const int n = resolve_n(n_cap, n_dev);
    // grids sized by a capacity: the CTAs that drew a tile past the end leave (tile ids are handed out in order, so
    // no live tile ever looks back at one of them)
    if (tile_base >= (long long)n) return;
    const int nvalid = (int)min((long long)TILE, (long long)n - tile_base);
    const bool full = nvalid == TILE;                               // CTA-uniform
    const u32 mask = (1u << bits) - 1u;
    const u32 sel = 0x4440u | (u32)(shift >> 3);

    // ---- load keys, warp-striped: item i of lane l of warp w sits at w*512 + i*32 + l ----
    u32 key[ITEMS];
    u32 rank[ITEMS];
    const int wbase = warp * (32 * ITEMS) + lane;   // item i of lane l of warp w sits at w*32*ITEMS + i*32 + l
    const u32* kin = keys_in + tile_base + wbase;
    if (full) {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) key[i] = kin[i * 32];
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) key[i] = (wbase + i * 32 < nvalid) ? kin[i * 32] : 0xFFFFFFFFu;
    }

    // ---- rank inside the warp, in memory order ----
    u32* wc = s_wcnt[warp];
    if (BYTE) {
        rank_items<8, true, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane);
    } else {
        switch (bits) {   // CTA-uniform
            case 8: rank_items<8, false, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane); break;
            case 7: rank_items<7, false, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane); break;
            case 6: rank_items<6, false, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane); break;
            case 5: rank_items<5, false, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane); break;
            case 4: rank_items<4, false, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane); break;
            case 3: rank_items<3, false, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane); break;
            case 2: rank_items<2, false, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane); break;
            default: rank_items<1, false, ITEMS, HWMATCH>(key, rank, wc, shift, mask, sel, lane); break;
        }
    }
    __syncthreads();

    // ---- per digit (one thread each): exclusive offsets across warps, tile count, look-back, global base ----
    {
        u32 sum = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            u32 c = s_wcnt[w][tid];
            s_wcnt[w][tid] = sum;
            sum += c;
        }
        u32* st = status + (size_t)tile * RADIX + tid;
        u32 excl = 0;
        if (tile > 0) {
            st_volatile_u32(st, sum | FLAG_AGG);
            // Decoupled look-back, several predecessors per round: the loads of a round are independent, so a walk
            // over k tiles costs about k / LOOKBACK memory round trips instead of k.  Measured on B200 (C3, per-view
            // depth sort / tile sort, us): 8 words -> 83 / 118, 16 -> 90 / 122, 32 -> 97 / 131, 64 -> 122 / 164:
            // the wider rounds spill (64-register cap) and the passes are bound by their instruction stream, not by
            // the chain; 4 words measured another 2-5% faster than 8 and is the default.
            int look = (int)tile - 1;
            bool done = lookback_round<SORT_LOOKBACK_FIRST>(status, tid, look, excl);
            while (!done) done = lookback_round<LOOKBACK>(status, tid, look, excl);
        }
        st_volatile_u32(st, ((excl + sum) & VAL_MASK) | FLAG_INCL);

        u32 local_excl = block_excl_scan_256(sum, s_wtot, tid);
        u32 gh = ghist[tid];
        u32 gexcl = block_excl_scan_256(gh, s_wtot, tid);
        s_gbase[tid] = gexcl + excl - local_excl;
        // fold the digit's offset inside the tile into the per-warp offsets: one table lookup per key below
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) s_wcnt[w][tid] += local_excl;
    }
    __syncthreads();

    // ---- scatter keys and values to their locally sorted position, then write coalesced runs ----
    {
        const u32* vin = vals_in + tile_base + wbase;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const u32 d = digit_of<BYTE>(key[i], shift, mask, sel);
            const u32 pos = wc[d] + rank[i];
            u32 v;
            if (IOTA) v = (u32)(tile_base + wbase + i * 32);
            else if (full) v = vin[i * 32];
            else v = (wbase + i * 32 < nvalid) ? vin[i * 32] : 0u;
            s_keys[pos] = key[i];
            s_vals[pos] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int p = i * SORT_THREADS + tid;
        if (full || p < nvalid) {
            const u32 k = s_keys[p];
            const u32 dst = s_gbase[digit_of<BYTE>(k, shift, mask, sel)] + (u32)p;
            keys_out[dst] = k;
            vals_out[dst] = s_vals[p];
        }
    }
}

// Zeroes the bookkeeping of one sort (digit histograms, dynamic tile counters, look-back words).  Separate from
// launch_sort_pairs so that a producer kernel can fill the digit histograms on the way (emit_fill does, for the tile
// sort) between the two.
int sort_prepare(SortScratch& sc, cudaStream_t st) {
    BAGS_CUDA_OK(cudaMemsetAsync(sc.ghist, 0, sc.zero_bytes, st));
    return 0;
}

// Sorts bits [begin_bit, end_bit).  Ping-pongs between (keys_in, vals_in) and (keys_out, vals_out); when the number
// of passes is even the result lands in the *_in buffers, so the final location is reported through final_*.
// n_dev: see resolve_n (NULL = n is exact).  hist_mode: SORT_FRESH / SORT_PREPARED / SORT_HIST_DONE (common.cuh).
int launch_sort_pairs(int n, const unsigned long long* n_dev, u32* keys_in, u32* vals_in, u32* keys_out,
                      u32* vals_out, int begin_bit, int end_bit, SortScratch& sc, bool iota_vals, int hist_mode,
                      cudaStream_t st, u32** final_keys, u32** final_vals, int nviews, size_t data_stride,
                      size_t count_stride) {
    if (nviews <= 1) { nviews = 1; data_stride = 0; count_stride = 0; }
    const int npass = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
    if (npass < 1 || npass > SORT_MAX_PASSES) {
        set_error("sort: unsupported bit range [%d,%d)", begin_bit, end_bit);
        return BAGS_E_INVALID;
    }
    if (n <= 0) {
        *final_keys = keys_in;
        *final_vals = vals_in;
        return 0;
    }
    if (hist_mode == SORT_FRESH) {
        if (nviews > 1) {
            set_error("sort: a batched sort needs its bookkeeping zeroed by the caller");
            return BAGS_E_INVALID;
        }
        int rc = sort_prepare(sc, st);
        if (rc) return rc;
    }
    if (hist_mode != SORT_HIST_DONE) {
        int hist_blocks = min((n + 512 * 8 - 1) / (512 * 8), 148 * 4);
        sort_hist_kernel<<<dim3(hist_blocks, nviews), 512, 0, st>>>(keys_in, n, n_dev, begin_bit, end_bit, sc.ghist,
                                                                     data_stride, count_stride);
        BAGS_LAUNCH_OK("sort_hist_kernel");
    }
    // Keys per CTA: 4096 when the pass runs in several waves (throughput bound), 2048 when even that leaves the GPU
    // under-filled -- a pass over 1 M keys is bounded by the latency of one CTA's instruction stream, which halves.
    // (the views of a batch are one launch: `nviews` times the CTAs)
    const bool small = g_sort_small_tiles && (long long)nviews * ((n + SORT_TILE - 1) / SORT_TILE) < 2 * 148;
    const int tile_keys = small ? SORT_TILE_SMALL : SORT_TILE;
    const int ntiles = (n + tile_keys - 1) / tile_keys;
    u32 *ki = keys_in, *vi = vals_in, *ko = keys_out, *vo = vals_out;
    int cur_dev = 0;
    BAGS_CUDA_OK(cudaGetDevice(&cur_dev));
    for (int p = 0; p < npass; p++) {
        int sft = begin_bit + p * RADIX_BITS;
        int bits = min(RADIX_BITS, end_bit - sft);
        u32* status = sc.status + (size_t)p * sc.ntiles * RADIX;
        const bool iota = p == 0 && iota_vals;
        const bool byte_digit = bits == 8 && (sft & 7) == 0;
        // the shared-memory attribute is per device and per kernel: set once for each (the flag is only a cache of
        // an idempotent call, so concurrent first calls from two host threads are harmless)
#define BAGS_PASS(IOTA, LB, BYTE, ITEMS, HWM)                                                                      \
    do {                                                                                                           \
        static std::atomic<unsigned char> attr_set[64];   /* per kernel variant (expansion site) and device */     \
        if (!attr_set[cur_dev & 63].load(std::memory_order_acquire)) {                                             \
            BAGS_CUDA_OK(cudaFuncSetAttribute(onesweep_pass_kernel<IOTA, LB, BYTE, ITEMS, HWM>,                    \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,                         \
                                              (int)sort_smem_bytes(ITEMS)));                                       \
            attr_set[cur_dev & 63].store(1, std::memory_order_release);                                            \
        }                                                                                                          \
        onesweep_pass_kernel<IOTA, LB, BYTE, ITEMS, HWM>                                                           \
            <<<dim3(ntiles, nviews), SORT_THREADS, sort_smem_bytes(ITEMS), st>>>(                                  \
                ki, vi, ko, vo, n, n_dev, sft, bits, sc.ghist + p * RADIX, sc.tile_counter + p, status,            \
                data_stride, count_stride);                                                                        \
    } while (0)
#define BAGS_PASS_I(IOTA, LB, BYTE)                                                                                \
    do {                                                                                                           \
        if (small && g_sort_hwmatch) BAGS_PASS(IOTA, LB, BYTE, SORT_ITEMS_SMALL, true);                            \
        else if (small) BAGS_PASS(IOTA, LB, BYTE, SORT_ITEMS_SMALL, false);                                        \
        else BAGS_PASS(IOTA, LB, BYTE, SORT_ITEMS, false);                                                         \
    } while (0)
#define BAGS_PASS_LB(LB)                                                                                           \
    do {                                                                                                           \
        if (iota && byte_digit) BAGS_PASS_I(true, LB, true);                                                       \
        else if (iota) BAGS_PASS_I(true, LB, false);                                                               \
        else if (byte_digit) BAGS_PASS_I(false, LB, true);                                                         \
        else BAGS_PASS_I(false, LB, false);                                                                        \
    } while (0)
        if (g_sort_lookback == 8) BAGS_PASS_LB(8);
        else BAGS_PASS_LB(4);
#undef BAGS_PASS_LB
#undef BAGS_PASS_I
#undef BAGS_PASS
        BAGS_LAUNCH_OK("onesweep_pass_kernel");
        u32* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    *final_keys = ki;
    *final_vals = vi;
    return 0;
}

// =================================================================================================================
// Emit (tile, id) instances in depth order, in two load-balanced steps:
//   scan : exclusive prefix sum of the tile counts in depth order (per-CTA totals, then every CTA adds up the totals
//          before it) -> offsets[i] = first output slot of the i-th Gaussian in depth order;
//   fill : one thread per OUTPUT slot.  A CTA owns EMIT_FILL_SLOTS consecutive slots, finds the first Gaussian of its
//          range with a warp-wide 32-ary search over `offsets`, stages that window of offsets in shared memory and
//          binary-searches it per slot.  Near Gaussians cover hundreds of tiles and far ones a single tile; filling
//          by slot instead of by Gaussian removes that skew (the per-Gaussian version spent most of its time in the
//          first CTAs' tails) and makes every store fully coalesced.
// =================================================================================================================
// count: per-CTA totals (no inter-CTA dependency).  scan: every CTA sums the totals of the CTAs before it (a block
// reduction over at most a few thousand words) and scans its own counts -- two short launches instead of a chained
// look-back, whose CTAs spent most of their life spinning on their predecessors' flags.
__device__ __forceinline__ u32 emit_count_of(int i, int P, const u32* __restrict__ sorted_idx,
                                             const ushort4* __restrict__ rect) {
    if (i >= P) return 0u;
    const ushort4 r = rect[sorted_idx[i]];
    return (u32)r.z * (u32)r.w;
}

__global__ void __launch_bounds__(EMIT_THREADS)
emit_count_kernel(int P, const u32* __restrict__ sorted_idx, const ushort4* __restrict__ rect,
                  u32* __restrict__ tile_total, const size_t geom_stride) {
    sorted_idx = at_view(sorted_idx, geom_stride, blockIdx.y);
    rect = at_view(rect, geom_stride, blockIdx.y);
    tile_total = at_view(tile_total, geom_stride, blockIdx.y);
    __shared__ u32 s_wsum[EMIT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u32 v = emit_count_of(blockIdx.x * EMIT_THREADS + tid, P, sorted_idx, rect);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
        u32 t = s_wsum[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) tile_total[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(EMIT_THREADS)
emit_scan_kernel(int P, const u32* __restrict__ sorted_idx, const ushort4* __restrict__ rect,
                 const u32* __restrict__ tile_total, u32* __restrict__ offsets, const size_t geom_stride) {
    sorted_idx = at_view(sorted_idx, geom_stride, blockIdx.y);
    rect = at_view(rect, geom_stride, blockIdx.y);
    tile_total = at_view(tile_total, geom_stride, blockIdx.y);
    offsets = at_view(offsets, geom_stride, blockIdx.y);
    __shared__ u32 s_wsum[EMIT_THREADS / 32];
    __shared__ u32 s_wbase[EMIT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * EMIT_THREADS + tid;
    const u32 cnt = emit_count_of(i, P, sorted_idx, rect);

    // sum of the totals of all CTAs before this one
    u32 before = 0;
    for (int j = tid; j < (int)blockIdx.x; j += EMIT_THREADS) before += tile_total[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) s_wbase[warp] = before;

    // warp inclusive scan of counts
    u32 inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    // lane k of every warp holds warp k's total / partial base: one more shuffle reduction gives this warp's offset
    u32 wsum = s_wsum[lane], base = s_wbase[lane];
    u32 wbase = (lane < warp) ? wsum : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wbase += __shfl_xor_sync(0xffffffffu, wbase, o);
        base += __shfl_xor_sync(0xffffffffu, base, o);
    }
    if (i < P) offsets[i] = base + wbase + inc - cnt;
}

constexpr int EMIT_FILL_THREADS = 256;
constexpr int EMIT_FILL_ITEMS = 8;
constexpr int EMIT_FILL_SLOTS = EMIT_FILL_THREADS * EMIT_FILL_ITEMS;   // output slots per CTA
constexpr int EMIT_FILL_WINDOW = EMIT_FILL_SLOTS + 1;                  // staged offsets (owners of 2048 slots when
                                                                       // every Gaussian has >= 1 tile)
constexpr int EMIT_FILL_PAD = 2 * EMIT_FILL_SLOTS;                     // power-of-two search array (sentinel padded)

// `ghist` (may be NULL): the digit histograms of the tile sort ([pass][RADIX], 8-bit digits of the tile id from bit 0,
// `hist_passes` of them) are accumulated here, where the tile ids are produced, instead of by a separate read of the
// R keys (sort_hist_kernel).
__global__ void __launch_bounds__(EMIT_FILL_THREADS)
emit_fill_kernel(int P, int R_cap, const unsigned long long* __restrict__ R_dev, const u32* __restrict__ offsets,
                 const u32* __restrict__ sorted_idx, const ushort4* __restrict__ rect, int grid_x,
                 u32* __restrict__ tile_out, u32* __restrict__ id_out, u32* __restrict__ ghist, int hist_passes,
                 const ViewStride vs) {
    if (vs.geom != 0) {   // batched launch: blockIdx.y = view
        R_dev = at_view(R_dev, vs.geom, blockIdx.y);
        offsets = at_view(offsets, vs.geom, blockIdx.y);
        sorted_idx = at_view(sorted_idx, vs.geom, blockIdx.y);
        rect = at_view(rect, vs.geom, blockIdx.y);
        tile_out = at_view(tile_out, vs.bin, blockIdx.y);
        id_out = at_view(id_out, vs.bin, blockIdx.y);
        if (ghist != nullptr) ghist = at_view(ghist, vs.bin, blockIdx.y);
    }
    __shared__ u32 s_off[EMIT_FILL_PAD];
    __shared__ u32 s_hist[2][RADIX];
    __shared__ int s_lo;
    const int tid = threadIdx.x, lane = tid & 31;
    const u32 slot0 = blockIdx.x * (u32)EMIT_FILL_SLOTS;
    const u32 R = (u32)resolve_n(R_cap, R_dev);
    if (slot0 >= R) return;   // CTA-uniform (grid sized by the capacity)
    if (ghist != nullptr)
        for (int k = tid; k < 2 * RADIX; k += EMIT_FILL_THREADS) (&s_hist[0][0])[k] = 0;

    // first Gaussian (in depth order) of this CTA's range: the last i with offsets[i] <= slot0
    if (tid < 32) {
        int lo = 0, hi = P;   // offsets[lo] <= slot0 (offsets[0] == 0); the answer lies in [lo, hi)
        while (hi - lo > 1) {
            const int step = (hi - lo + 31) >> 5;
            const int p = lo + lane * step;
            const bool le = (p < hi) && (offsets[p] <= slot0);
            const int c = __popc(__ballot_sync(0xffffffffu, le));   // offsets are non-decreasing: a prefix of lanes
            lo += (c - 1) * step;
            hi = min(lo + step, hi);
        }
        if (lane == 0) s_lo = lo;
    }
    __syncthreads();
    const int i_lo = s_lo;
    const int nwin = min(EMIT_FILL_WINDOW, P - i_lo);
    for (int k = tid; k < EMIT_FILL_PAD; k += EMIT_FILL_THREADS) s_off[k] = (k < nwin) ? offsets[i_lo + k] : 0xFFFFFFFFu;
    __syncthreads();

    // All EMIT_FILL_ITEMS slots of a thread advance together (branch-free binary search, then the two dependent
    // gathers), so their shared-memory and global latencies overlap.
    u32 slot[EMIT_FILL_ITEMS];
    int own[EMIT_FILL_ITEMS];
#pragma unroll
    for (int it = 0; it < EMIT_FILL_ITEMS; it++) {
        slot[it] = slot0 + (u32)(it * EMIT_FILL_THREADS + tid);
        own[it] = 0;
    }
#pragma unroll
    for (int step = EMIT_FILL_PAD / 2; step > 0; step >>= 1) {
#pragma unroll
        for (int it = 0; it < EMIT_FILL_ITEMS; it++)
            if (s_off[own[it] + step] <= slot[it]) own[it] += step;   // last j with s_off[j] <= slot (s_off[0] <= slot0)
    }
    u32 start[EMIT_FILL_ITEMS], g[EMIT_FILL_ITEMS];
#pragma unroll
    for (int it = 0; it < EMIT_FILL_ITEMS; it++) {
        int i = i_lo + own[it];
        start[it] = s_off[own[it]];
        if (slot[it] < R && own[it] == nwin - 1 && i + 1 < P && offsets[i + 1] <= slot[it]) {
            // more than EMIT_FILL_SLOTS Gaussians share this CTA's range (zero-count entries inside the list):
            // finish the search in global memory
            int glo = i + 1, ghi = P;
            while (ghi - glo > 1) {
                const int mid = (glo + ghi) >> 1;
                if (offsets[mid] <= slot[it]) glo = mid; else ghi = mid;
            }
            i = glo;
            start[it] = offsets[glo];
        }
        g[it] = (slot[it] < R) ? sorted_idx[i] : 0u;
    }
    ushort4 r[EMIT_FILL_ITEMS];
#pragma unroll
    for (int it = 0; it < EMIT_FILL_ITEMS; it++) r[it] = rect[g[it]];
#pragma unroll
    for (int it = 0; it < EMIT_FILL_ITEMS; it++) {
        if (slot[it] >= R) continue;
        const u32 local = slot[it] - start[it];
        const u32 w = max((u32)r[it].z, 1u);
        const u32 ty = (u32)r[it].y + local / w;
        const u32 tx = (u32)r[it].x + local % w;
        const u32 tile = ty * (u32)grid_x + tx;
        tile_out[slot[it]] = tile;
        id_out[slot[it]] = g[it];
        if (ghist != nullptr) {
            atomicAdd(&s_hist[0][tile & (RADIX - 1)], 1u);
            if (hist_passes > 1) atomicAdd(&s_hist[1][(tile >> RADIX_BITS) & (RADIX - 1)], 1u);
        }
    }
    if (ghist != nullptr) {
        __syncthreads();
        for (int k = tid; k < hist_passes * RADIX; k += EMIT_FILL_THREADS) {
            const u32 v = (&s_hist[0][0])[k];
            if (v) atomicAdd(&ghist[k], v);
        }
    }
}

int launch_emit_scan(int P, const GeomState& g, const u32* sorted_idx, u32* offsets, cudaStream_t st, int nviews,
                     const ViewStride& vs) {
    const int blocks = (P + EMIT_THREADS - 1) / EMIT_THREADS;
    const dim3 grid(blocks, nviews > 1 ? nviews : 1);
    const size_t gs = nviews > 1 ? vs.geom : 0;
    emit_count_kernel<<<grid, EMIT_THREADS, 0, st>>>(P, sorted_idx, g.rect, g.emit_status, gs);
    BAGS_LAUNCH_OK("emit_count_kernel");
    emit_scan_kernel<<<grid, EMIT_THREADS, 0, st>>>(P, sorted_idx, g.rect, g.emit_status, offsets, gs);
    BAGS_LAUNCH_OK("emit_scan_kernel");
    return 0;
}

// hist_passes: 1 or 2 -> the tile sort's digit histograms are accumulated into b.sort.ghist (sort_prepare() must have
// run); 0 -> not (more than two 8-bit digits: the sort makes its own histogram pass).
int launch_emit_fill(int P, int R, const unsigned long long* R_dev, const GeomState& g, const u32* sorted_idx,
                     const u32* offsets, BinningState& b, int grid_x, int hist_passes, cudaStream_t st, int nviews,
                     const ViewStride& vs) {
    const int blocks = (R + EMIT_FILL_SLOTS - 1) / EMIT_FILL_SLOTS;
    emit_fill_kernel<<<dim3(blocks, nviews > 1 ? nviews : 1), EMIT_FILL_THREADS, 0, st>>>(
        P, R, R_dev, offsets, sorted_idx, g.rect, grid_x, b.tile_buf[0], b.id_buf[0],
        hist_passes > 0 ? b.sort.ghist : nullptr, hist_passes, nviews > 1 ? vs : ViewStride());
    BAGS_LAUNCH_OK("emit_fill_kernel");
    return 0;
}

// =================================================================================================================
// Tile ranges from the boundaries of the sorted tile ids (identifyTileRanges, rasterizer_impl.cu:116-138).
// Untouched tiles keep the (0,0) of the memset, like the reference's cudaMemset.
// =================================================================================================================
constexpr int RANGE_ITEMS = 4;
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int R_cap, const unsigned long long* __restrict__ R_dev, const u32* __restrict__ tile_sorted,
                   uint2* __restrict__ ranges, const ViewStride vs) {
    if (R_dev != nullptr) R_dev = at_view(R_dev, vs.geom, blockIdx.y);
    tile_sorted = at_view(tile_sorted, vs.bin, blockIdx.y);
    ranges = at_view(ranges, vs.img, blockIdx.y);
    const int R = resolve_n(R_cap, R_dev);
    // thread-striped over RANGE_ITEMS * 256 consecutive elements per CTA (coalesced loads)
    const int base = blockIdx.x * (256 * RANGE_ITEMS) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < RANGE_ITEMS; k++) {
        const int i = base + k * 256;
        if (i >= R) break;
        const u32 t = tile_sorted[i];
        if (i == 0) {
            ranges[t].x = 0;
        } else {
            const u32 prev = tile_sorted[i - 1];
            if (t != prev) {
                ranges[prev].y = (u32)i;
                ranges[t].x = (u32)i;
            }
        }
        if (i == R - 1) ranges[t].y = (u32)R;
    }
}

int launch_ranges(int R, const unsigned long long* R_dev, int ntiles, const u32* tile_sorted, uint2* ranges,
                  cudaStream_t st, int nviews, const ViewStride& vs) {
    if (nviews > 1)
        BAGS_CUDA_OK(cudaMemset2DAsync(ranges, vs.img, 0, (size_t)ntiles * sizeof(uint2), (size_t)nviews, st));
    else
        BAGS_CUDA_OK(cudaMemsetAsync(ranges, 0, (size_t)ntiles * sizeof(uint2), st));
    if (R > 0) {
        tile_ranges_kernel<<<dim3((R + 256 * RANGE_ITEMS - 1) / (256 * RANGE_ITEMS), nviews > 1 ? nviews : 1), 256, 0,
                             st>>>(R, R_dev, tile_sorted, ranges, nviews > 1 ? vs : ViewStride());
        BAGS_LAUNCH_OK("tile_ranges_kernel");
    }
    return 0;
}

// =================================================================================================================
// Longest-first launch order of the blend CTAs.  The hardware hands out CTAs in blockIdx order and the blend time of a
// tile is proportional to the length of its list (forward) / its largest n_contrib (backward), which varies by 3x and
// more over an image: in raster order the long tiles that happen to come last run alone at the end of the kernel (list
// scheduling simulation on the bench cameras: makespan 6-20% above the mean load; longest first: 1-3%).  One CTA per
// view: counting sort of the tiles by cost / 4 into 1024 buckets, descending; the order inside a bucket is arbitrary.
// =================================================================================================================
constexpr int ORDER_THREADS = 1024;
constexpr int ORDER_BUCKETS = 1024;
__global__ void __launch_bounds__(ORDER_THREADS)
tile_order_kernel(int ntiles, const uint2* __restrict__ ranges, const u32* __restrict__ cost_in,
                  u32* __restrict__ order, const ViewStride vs) {
    if (ranges != nullptr) ranges = at_view(ranges, vs.img, blockIdx.x);
    if (cost_in != nullptr) cost_in = at_view(cost_in, vs.img, blockIdx.x);
    order = at_view(order, vs.img, blockIdx.x);
    __shared__ u32 s_cnt[ORDER_BUCKETS];
    __shared__ u32 s_wtot[ORDER_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_cnt[tid] = 0;
    __syncthreads();
    auto bucket_of = [&](int t) -> u32 {
        u32 c;
        if (ranges != nullptr) {
            const uint2 r = ranges[t];
            c = r.y - r.x;
        } else {
            c = cost_in[t];
        }
        c >>= 2;
        // descending: bucket 0 holds the most expensive tiles
        return (u32)(ORDER_BUCKETS - 1) - (c < (u32)ORDER_BUCKETS ? c : (u32)(ORDER_BUCKETS - 1));
    };
    for (int t = tid; t < ntiles; t += ORDER_THREADS) atomicAdd(&s_cnt[bucket_of(t)], 1u);
    __syncthreads();
    // exclusive scan of the 1024 bucket counts (one per thread)
    const u32 v = s_cnt[tid];
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (lane == 31) s_wtot[warp] = inc;
    __syncthreads();
    u32 wbase = 0;
    for (int w = 0; w < warp; w++) wbase += s_wtot[w];
    __syncthreads();
    s_cnt[tid] = wbase + inc - v;
    __syncthreads();
    for (int t = tid; t < ntiles; t += ORDER_THREADS) order[atomicAdd(&s_cnt[bucket_of(t)], 1u)] = (u32)t;
}

int launch_tile_order(int ntiles, const uint2* ranges, const u32* cost_in, u32* order, cudaStream_t st, int nviews,
                      const ViewStride& vs) {
    tile_order_kernel<<<nviews > 1 ? nviews : 1, ORDER_THREADS, 0, st>>>(ntiles, ranges, cost_in, order,
                                                                       nviews > 1 ? vs : ViewStride());
    BAGS_LAUNCH_OK("tile_order_kernel");
    return 0;
}

int set_sort_small_tiles(int value) {
    const int prev = g_sort_small_tiles;
    if (value == 0 || value == 1) g_sort_small_tiles = value;
    return prev;
}
int set_sort_hwmatch(int value) {
    const int prev = g_sort_hwmatch;
    if (value == 0 || value == 1) g_sort_hwmatch = value;
    return prev;
}
int set_sort_lookback(int value) {
    const int prev = g_sort_lookback;
    if (value == 4 || value == 8) g_sort_lookback = value;
    return prev;
}

}  // namespace bags

using namespace bags;

extern "C" size_t bags_sort_scratch_bytes(int n) { return SortScratch::bytes(n < 0 ? 0 : n) + 256; }

extern "C" int bags_sort_pairs_u32(int n, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out,
                                   uint32_t* vals_out, int begin_bit, int end_bit, char* scratch, void* stream) {
    if (n < 0 || begin_bit < 0 || end_bit > 32 || begin_bit >= end_bit || !scratch ||
        (n > 0 && (!keys_in || !vals_in || !keys_out || !vals_out))) {
        set_error("bags_sort_pairs_u32: invalid argument");
        return BAGS_E_INVALID;
    }
    if (n >= (1 << 30)) {
        set_error("bags_sort_pairs_u32: n=%d exceeds the 30-bit look-back word", n);
        return BAGS_E_OVERFLOW;
    }
    char* p = (char*)align_up((size_t)scratch);
    SortScratch sc = SortScratch::from(p, n);
    u32 *fk, *fv;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_sort_pairs(n, nullptr, keys_in, vals_in, keys_out, vals_out, begin_bit, end_bit, sc, false,
                               SORT_FRESH, st, &fk, &fv);
    if (rc) return rc;
    if (n > 0 && fk != keys_out) {
        BAGS_CUDA_OK(cudaMemcpyAsync(keys_out, fk, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        BAGS_CUDA_OK(cudaMemcpyAsync(vals_out, fv, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}
