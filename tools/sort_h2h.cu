// Head-to-head: the hand-written onesweep of libbags_b200.so (bags_sort_pairs_u32) against cub::DeviceRadixSort (the
// CUB that ships with this CUDA toolkit; the reference's sort, rasterizer_impl.cu:313-318, is its u64/u32 SortPairs) on
// the three problems of one C3 view.  Development / evidence tool, not part of the product:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/sort_h2h.cu -o gpurun_out/sort_h2h \
//        -L bundle-adjusting-gaussian-splatting_b200 -lbags_b200 -Xlinker -rpath -Xlinker $PWD/bundle-adjusting-gaussian-splatting_b200
// Prints one JSON line per problem: median of 50 timed runs after 5 warm-ups, CUDA events, inputs re-copied (untimed)
// before every run, results compared element by element (both sorts are stable, so they must be identical).
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../include/bags_b200.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

static float median(std::vector<float> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

template <typename F, typename G>
static float time_runs(F&& prepare, G&& run, int iters = 50) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::vector<float> ms;
    for (int i = 0; i < iters + 5; i++) {
        prepare();
        CK(cudaEventRecord(a));
        run();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float t; CK(cudaEventElapsedTime(&t, a, b));
        if (i >= 5) ms.push_back(t);
    }
    return median(ms);
}

static void problem_u32(const char* name, int n, int bits, unsigned seed, bool depth_like) {
    std::vector<uint32_t> hk(n), hv(n);
    std::mt19937 rng(seed);
    for (int i = 0; i < n; i++) {
        if (depth_like) {                       // float bits of depths in (0.2, 8), 36 % culled (0xFFFFFFFF)
            float z = 0.2f + 7.8f * (rng() / 4294967296.0f);
            uint32_t u; memcpy(&u, &z, 4);
            hk[i] = (rng() % 100 < 36) ? 0xFFFFFFFFu : u;
        } else {
            hk[i] = rng() % (1u << bits);
        }
        hv[i] = (uint32_t)i;
    }
    uint32_t *k0, *v0, *ka, *va, *kb, *vb, *kc, *vc;
    size_t B = (size_t)n * 4;
    CK(cudaMalloc(&k0, B)); CK(cudaMalloc(&v0, B)); CK(cudaMalloc(&ka, B)); CK(cudaMalloc(&va, B));
    CK(cudaMalloc(&kb, B)); CK(cudaMalloc(&vb, B)); CK(cudaMalloc(&kc, B)); CK(cudaMalloc(&vc, B));
    CK(cudaMemcpy(k0, hk.data(), B, cudaMemcpyHostToDevice)); CK(cudaMemcpy(v0, hv.data(), B, cudaMemcpyHostToDevice));
    char* scratch; CK(cudaMalloc(&scratch, bags_sort_scratch_bytes(n)));
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, kc, v0, vc, n, 0, bits);
    void* temp; CK(cudaMalloc(&temp, tb));
    auto prep = [&] { CK(cudaMemcpyAsync(ka, k0, B, cudaMemcpyDeviceToDevice)); CK(cudaMemcpyAsync(va, v0, B, cudaMemcpyDeviceToDevice)); };
    auto ours = [&] { if (bags_sort_pairs_u32(n, ka, va, kb, vb, 0, bits, scratch, nullptr) < 0) { printf("ours failed: %s\n", bags_last_error()); exit(1); } };
    auto cubr = [&] { cub::DeviceRadixSort::SortPairs(temp, tb, ka, kc, va, vc, n, 0, bits); };
    float t_ours = time_runs(prep, ours);
    float t_cub = time_runs(prep, cubr);
    prep(); ours(); prep(); cubr(); CK(cudaDeviceSynchronize());
    std::vector<uint32_t> r1(n), r2(n), s1(n), s2(n);
    CK(cudaMemcpy(r1.data(), kb, B, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r2.data(), kc, B, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(s1.data(), vb, B, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(s2.data(), vc, B, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int i = 0; i < n; i++) bad += (r1[i] != r2[i]) || (s1[i] != s2[i]);
    printf("{\"problem\": \"%s\", \"n\": %d, \"key_bits\": %d, \"ours_us\": %.1f, \"cub_us\": %.1f, \"cub_over_ours\": %.3f, \"mismatches\": %lld, "
           "\"ours\": \"bags_sort_pairs_u32 (histogram + %d onesweep passes, incl. its memset)\", \"cub\": \"cub::DeviceRadixSort::SortPairs<u32,u32> bits [0,%d)\"}\n",
           name, n, bits, t_ours * 1e3, t_cub * 1e3, t_cub / t_ours, bad, (bits + 7) / 8, bits);
    cudaFree(k0); cudaFree(v0); cudaFree(ka); cudaFree(va); cudaFree(kb); cudaFree(vb); cudaFree(kc); cudaFree(vc);
    cudaFree(scratch); cudaFree(temp);
}

// what the reference sorts: R pairs of (u64 tile << 32 | depth bits, u32 id) over 32 + 13 key bits
static void problem_u64(int n, int tile_bits, unsigned seed) {
    std::vector<uint64_t> hk(n);
    std::vector<uint32_t> hv(n);
    std::mt19937 rng(seed);
    for (int i = 0; i < n; i++) {
        float z = 0.2f + 7.8f * (rng() / 4294967296.0f);
        uint32_t u; memcpy(&u, &z, 4);
        hk[i] = ((uint64_t)(rng() % (1u << tile_bits)) << 32) | u;
        hv[i] = (uint32_t)i;
    }
    uint64_t *k0, *ka, *kc; uint32_t *v0, *va, *vc;
    CK(cudaMalloc(&k0, (size_t)n * 8)); CK(cudaMalloc(&ka, (size_t)n * 8)); CK(cudaMalloc(&kc, (size_t)n * 8));
    CK(cudaMalloc(&v0, (size_t)n * 4)); CK(cudaMalloc(&va, (size_t)n * 4)); CK(cudaMalloc(&vc, (size_t)n * 4));
    CK(cudaMemcpy(k0, hk.data(), (size_t)n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(v0, hv.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, kc, v0, vc, n, 0, 32 + tile_bits);
    void* temp; CK(cudaMalloc(&temp, tb));
    auto prep = [&] { CK(cudaMemcpyAsync(ka, k0, (size_t)n * 8, cudaMemcpyDeviceToDevice)); CK(cudaMemcpyAsync(va, v0, (size_t)n * 4, cudaMemcpyDeviceToDevice)); };
    auto cubr = [&] { cub::DeviceRadixSort::SortPairs(temp, tb, ka, kc, va, vc, n, 0, 32 + tile_bits); };
    float t = time_runs(prep, cubr);
    printf("{\"problem\": \"reference_sort_u64\", \"n\": %d, \"key_bits\": %d, \"cub_us\": %.1f, \"cub\": \"cub::DeviceRadixSort::SortPairs<u64,u32> bits [0,%d) "
           "(rasterizer_impl.cu:313-318); ours replaces it with the 1M x 32-bit depth sort + the %d-bit tile sort above\"}\n",
           n, 32 + tile_bits, t * 1e3, 32 + tile_bits, tile_bits);
    cudaFree(k0); cudaFree(ka); cudaFree(kc); cudaFree(v0); cudaFree(va); cudaFree(vc); cudaFree(temp);
}

int main(int argc, char** argv) {
    int R = argc > 1 ? atoi(argv[1]) : 4747220;     // mean tile instances of the bench views at C3
    problem_u32("depth_sort_1M_x_32bit", 1000000, 32, 1, true);
    problem_u32("tile_sort_R_x_13bit", R, 13, 2, false);
    problem_u32("tile_sort_8M_x_13bit", 8000000, 13, 3, false);
    problem_u32("depth_sort_6M_x_32bit", 6000000, 32, 4, true);
    problem_u64(R, 13, 5);
    return 0;
}
