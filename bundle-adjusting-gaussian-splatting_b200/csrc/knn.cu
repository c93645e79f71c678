// simple-knn replacement (distCUDA2): placeholder translation unit, filled in by knn kernels below.
#include "../../include/bags_b200.h"
#include "common.cuh"
using namespace bags;
extern "C" size_t bags_knn_scratch_bytes(int P) { return 256 + (size_t)(P < 0 ? 0 : P) * 64; }
extern "C" int bags_knn_mean_dist2(int P, const float* points, float* mean_dist2, char* scratch, void* stream) {
    (void)P; (void)points; (void)mean_dist2; (void)scratch; (void)stream;
    set_error("bags_knn_mean_dist2: not implemented yet");
    return BAGS_E_INVALID;
}
