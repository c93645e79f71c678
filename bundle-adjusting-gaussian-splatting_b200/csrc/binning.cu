// Tile binning for sm_100a: hand-written onesweep radix sort, chained-scan tile-instance emission, tile ranges.
//
// The reference bins with duplicateWithKeys -> cub::DeviceRadixSort::SortPairs on 64-bit (tile << 32 | depth) keys
// over 32+bit key bits -> identifyTileRanges (3dgs-pose/cuda_rasterizer/rasterizer_impl.cu:70-138,287-327).
// A stable LSD sort of that key over instances emitted in (Gaussian id, tile) order yields the order
// (tile, depth bits, Gaussian id).  The same order is produced here with far less traffic by a two-level sort:
//
//   1. depth sort   : stable sort of the P Gaussians by their 32 depth bits (culled ones carry 0xFFFFFFFF),
//                     4 x 8-bit onesweep passes over P x 8 B  -> order (depth, id);
//   2. emit         : walk Gaussians in that order, append (tile, id) for every tile of the stored rectangle
//                     (y outer, x inner like rasterizer_impl.cu:98-108); output offsets come from a decoupled
//                     look-back scan fused into the same kernel (no separate InclusiveSum, no offsets array);
//   3. tile sort    : stable sort of the R instances by tile id only, ceil(bit/8) (= 2) passes over R x 8 B
//                     -> order (tile, depth, id), identical to the reference's point_list;
//   4. ranges       : per-tile binary search in the sorted tile ids; empty tiles get (0,0) like the reference's
//                     cudaMemset.
//
// Per tile instance this moves 4 (histogram) + 2 x 16 (passes) + 8 (emit) bytes instead of the reference's
// 12 (emit) + 8 + 6 x 24 (sort) + 8 (ranges).
#include "../../include/bags_b200.h"
#include "common.cuh"

namespace bags {

// =================================================================================================================
// Onesweep radix sort (u32 keys, u32 values), 8-bit digits, one CTA per 8192-key tile, decoupled look-back.
// =================================================================================================================
constexpr size_t SORT_SMEM_BYTES = (SORT_TILE + SORT_WARPS * RADIX + 2 * RADIX + 8) * sizeof(u32);
constexpr u32 FLAG_AGG = 1u << 30;   // tile aggregate available
constexpr u32 FLAG_INCL = 2u << 30;  // inclusive prefix available
constexpr u32 VAL_MASK = (1u << 30) - 1;

__device__ __forceinline__ u32 ld_volatile_u32(const u32* p) {
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(u32* p, u32 v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Histograms of every digit position in one read of the keys.
__global__ void __launch_bounds__(512)
sort_hist_kernel(const u32* __restrict__ keys, int n, int begin_bit, int end_bit, u32* __restrict__ ghist) {
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

// Exclusive scan of 256 values held one per thread by the first 256 threads of the CTA (8 warps).
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

template <bool IOTA>
__global__ void __launch_bounds__(SORT_THREADS)
onesweep_pass_kernel(const u32* __restrict__ keys_in, const u32* __restrict__ vals_in, u32* __restrict__ keys_out,
                     u32* __restrict__ vals_out, int n, int shift, int bits, const u32* __restrict__ ghist,
                     u32* __restrict__ tile_counter, u32* __restrict__ status) {
    extern __shared__ u32 sort_smem[];
    u32* s_buf = sort_smem;                                              // [SORT_TILE] staging, locally sorted order
    u32 (*s_wcnt)[RADIX] = reinterpret_cast<u32 (*)[RADIX]>(s_buf + SORT_TILE);  // [SORT_WARPS][RADIX]
    u32* s_local = s_buf + SORT_TILE + SORT_WARPS * RADIX;               // [RADIX] digit offset inside the tile
    u32* s_gbase = s_local + RADIX;                                      // [RADIX] global index of local position 0
    u32* s_wtot = s_gbase + RADIX;                                       // [8]
    __shared__ u32 s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s_wcnt[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const long long tile_base = (long long)tile * SORT_TILE;
    const int nvalid = (int)min((long long)SORT_TILE, (long long)n - tile_base);
    const u32 mask = (1u << bits) - 1u;

    // ---- load keys, warp-striped: item i of lane l of warp w sits at w*512 + i*32 + l ----
    u32 key[SORT_ITEMS];
    u32 rank[SORT_ITEMS];
    const int wbase = warp * (32 * SORT_ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        int loc = wbase + i * 32;
        key[i] = (loc < nvalid) ? keys_in[tile_base + loc] : 0xFFFFFFFFu;
    }

    // ---- rank inside the warp, in memory order, with match_any ----
    const u32 lt_mask = (1u << lane) - 1u;
    u32* wc = s_wcnt[warp];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        int loc = wbase + i * 32;
        u32 d = (loc < nvalid) ? ((key[i] >> shift) & mask) : (u32)(RADIX - 1);
        u32 peers = __match_any_sync(0xffffffffu, d);
        u32 lower = __popc(peers & lt_mask);
        int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader) {
            old = wc[d];
            wc[d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = old + lower;
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive offsets across warps, tile count, look-back, global base ----
    if (tid < RADIX) {
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
            const u32* look = st - RADIX;
            while (true) {
                u32 s = ld_volatile_u32(look);
                u32 f = s & ~VAL_MASK;
                if (f == 0) continue;
                excl += s & VAL_MASK;
                if (f == FLAG_INCL) break;
                look -= RADIX;
            }
        }
        st_volatile_u32(st, ((excl + sum) & VAL_MASK) | FLAG_INCL);

        u32 local_excl = block_excl_scan_256(sum, s_wtot, tid);
        u32 gh = ghist[tid];
        u32 gexcl = block_excl_scan_256(gh, s_wtot, tid);
        s_local[tid] = local_excl;
        s_gbase[tid] = gexcl + excl - local_excl;
    }
    __syncthreads();

    // ---- scatter keys to their locally sorted position, then write coalesced runs ----
    u32 pos[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        int loc = wbase + i * 32;
        u32 d = (loc < nvalid) ? ((key[i] >> shift) & mask) : (u32)(RADIX - 1);
        pos[i] = s_local[d] + wc[d] + rank[i];
        s_buf[pos[i]] = key[i];
    }
    __syncthreads();
    u32 dst[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        int p = i * SORT_THREADS + tid;
        if (p < nvalid) {
            u32 k = s_buf[p];
            u32 d = (k >> shift) & mask;
            dst[i] = s_gbase[d] + (u32)p;
            keys_out[dst[i]] = k;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        int loc = wbase + i * 32;
        u32 v = 0;
        if (loc < nvalid) v = IOTA ? (u32)(tile_base + loc) : vals_in[tile_base + loc];
        s_buf[pos[i]] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        int p = i * SORT_THREADS + tid;
        if (p < nvalid) vals_out[dst[i]] = s_buf[p];
    }
}

// Sorts bits [begin_bit, end_bit).  Ping-pongs between (keys_in, vals_in) and (keys_out, vals_out); when the number
// of passes is even the result lands in the *_in buffers, so the final location is reported through final_*.
int launch_sort_pairs(int n, u32* keys_in, u32* vals_in, u32* keys_out, u32* vals_out, int begin_bit, int end_bit,
                      SortScratch& sc, bool iota_vals, cudaStream_t st, u32** final_keys, u32** final_vals) {
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
    BAGS_CUDA_OK(cudaMemsetAsync(sc.ghist, 0, sc.zero_bytes, st));
    int hist_blocks = min((n + 512 * 8 - 1) / (512 * 8), 148 * 4);
    sort_hist_kernel<<<hist_blocks, 512, 0, st>>>(keys_in, n, begin_bit, end_bit, sc.ghist);
    BAGS_LAUNCH_OK("sort_hist_kernel");
    const int ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    // per-device attribute; setting it is a cheap driver call, so do it on every sort (multi-device safe)
    BAGS_CUDA_OK(cudaFuncSetAttribute(onesweep_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SORT_SMEM_BYTES));
    BAGS_CUDA_OK(cudaFuncSetAttribute(onesweep_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SORT_SMEM_BYTES));
    u32 *ki = keys_in, *vi = vals_in, *ko = keys_out, *vo = vals_out;
    for (int p = 0; p < npass; p++) {
        int sft = begin_bit + p * RADIX_BITS;
        int bits = min(RADIX_BITS, end_bit - sft);
        u32* status = sc.status + (size_t)p * sc.ntiles * RADIX;
        if (p == 0 && iota_vals)
            onesweep_pass_kernel<true><<<ntiles, SORT_THREADS, SORT_SMEM_BYTES, st>>>(ki, vi, ko, vo, n, sft, bits,
                                                                        sc.ghist + p * RADIX, sc.tile_counter + p,
                                                                        status);
        else
            onesweep_pass_kernel<false><<<ntiles, SORT_THREADS, SORT_SMEM_BYTES, st>>>(ki, vi, ko, vo, n, sft, bits,
                                                                         sc.ghist + p * RADIX, sc.tile_counter + p,
                                                                         status);
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
// Emit (tile, id) instances in depth order; output offsets by a chained (decoupled look-back) scan in the same kernel.
// =================================================================================================================
__global__ void __launch_bounds__(EMIT_THREADS)
emit_instances_kernel(int P, const u32* __restrict__ sorted_idx, const ushort4* __restrict__ rect, int grid_x,
                      unsigned long long* __restrict__ counters, u32* __restrict__ emit_status,
                      u32* __restrict__ tile_out, u32* __restrict__ id_out) {
    __shared__ u32 s_wsum[EMIT_THREADS / 32];
    __shared__ u32 s_tile;
    __shared__ u32 s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (u32)atomicAdd(&counters[1], 1ull);
    __syncthreads();
    const u32 tile = s_tile;
    const int i = tile * EMIT_THREADS + tid;

    u32 g = 0;
    ushort4 r = make_ushort4(0, 0, 0, 0);
    if (i < P) {
        g = sorted_idx[i];
        r = rect[g];
    }
    const u32 cnt = (u32)r.z * (u32)r.w;

    // warp inclusive scan of counts
    u32 inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    u32 wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < EMIT_THREADS / 32; k++) {
        u32 v = s_wsum[k];
        if (k < warp) wbase += v;
        total += v;
    }
    // chained scan across CTAs (one word: 2 flag bits + 30 value bits)
    if (tid == 0) {
        u32 excl = 0;
        if (tile > 0) {
            st_volatile_u32(&emit_status[tile], total | FLAG_AGG);
            int look = (int)tile - 1;
            while (true) {
                u32 s = ld_volatile_u32(&emit_status[look]);
                u32 f = s & ~VAL_MASK;
                if (f == 0) continue;
                excl += s & VAL_MASK;
                if (f == FLAG_INCL) break;
                look--;
            }
        }
        st_volatile_u32(&emit_status[tile], ((excl + total) & VAL_MASK) | FLAG_INCL);
        s_base = excl;
    }
    __syncthreads();
    const u32 start = s_base + wbase + inc - cnt;  // first output slot of this lane's Gaussian

    // warp-cooperative, coalesced emission: slot s of the warp's contiguous output range belongs to the lane whose
    // [start, start+cnt) contains it.
    const u32 wstart = __shfl_sync(0xffffffffu, start, 0);
    const u32 wtotal = __shfl_sync(0xffffffffu, inc, 31);
    const u32 packed_min = (u32)r.x | ((u32)r.y << 16);
    const u32 width = r.z;
    // NB: the loop trip count is warp-uniform; every lane takes part in the shuffles and only the stores are guarded.
    for (u32 sbase = 0; sbase < wtotal; sbase += 32) {
        const u32 s = sbase + lane;
        const bool active = s < wtotal;
        const u32 slot = wstart + (active ? s : 0u);
        // largest lane k with start_k <= slot (starts are non-decreasing)
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = lo + step;   // <= 31
            const u32 sk = __shfl_sync(0xffffffffu, start, cand);
            if (sk <= slot) lo = cand;
        }
        // zero-count lanes share their start with the next lane; the search lands on the last lane of such a run,
        // which is the non-empty owner because slot < wstart + wtotal.
        const u32 ostart = __shfl_sync(0xffffffffu, start, lo);
        const u32 omin = __shfl_sync(0xffffffffu, packed_min, lo);
        const u32 ow = __shfl_sync(0xffffffffu, width, lo);
        const u32 og = __shfl_sync(0xffffffffu, g, lo);
        if (active) {
            const u32 local = slot - ostart;
            const u32 ty = (omin >> 16) + local / ow;
            const u32 tx = (omin & 0xFFFFu) + local % ow;
            tile_out[slot] = ty * (u32)grid_x + tx;
            id_out[slot] = og;
        }
    }
}

int launch_emit(int P, const GeomState& g, const u32* sorted_idx, BinningState& b, int grid_x, cudaStream_t st) {
    const int blocks = (P + EMIT_THREADS - 1) / EMIT_THREADS;
    emit_instances_kernel<<<blocks, EMIT_THREADS, 0, st>>>(P, sorted_idx, g.rect, grid_x, g.counters,
                                                          g.emit_status, b.tile_buf[0], b.id_buf[0]);
    BAGS_LAUNCH_OK("emit_instances_kernel");
    return 0;
}

// =================================================================================================================
// Tile ranges by binary search over the sorted tile ids.
// =================================================================================================================
__device__ __forceinline__ u32 lower_bound_u32(const u32* __restrict__ a, u32 n, u32 key) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void tile_ranges_kernel(int R, int ntiles, const u32* __restrict__ tile_sorted,
                                   uint2* __restrict__ ranges) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    u32 lo = 0, hi = 0;
    if (R > 0) {
        lo = lower_bound_u32(tile_sorted, (u32)R, (u32)t);
        hi = lower_bound_u32(tile_sorted, (u32)R, (u32)t + 1u);
    }
    if (lo == hi) { lo = 0; hi = 0; }   // untouched tiles keep the reference's memset value
    ranges[t] = make_uint2(lo, hi);
}

int launch_ranges(int R, int ntiles, const u32* tile_sorted, uint2* ranges, cudaStream_t st) {
    tile_ranges_kernel<<<(ntiles + 127) / 128, 128, 0, st>>>(R, ntiles, tile_sorted, ranges);
    BAGS_LAUNCH_OK("tile_ranges_kernel");
    return 0;
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
    int rc = launch_sort_pairs(n, keys_in, vals_in, keys_out, vals_out, begin_bit, end_bit, sc, false, st, &fk, &fv);
    if (rc) return rc;
    if (n > 0 && fk != keys_out) {
        BAGS_CUDA_OK(cudaMemcpyAsync(keys_out, fk, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        BAGS_CUDA_OK(cudaMemcpyAsync(vals_out, fv, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}
