// K2 / K3 (tcgen05) - placeholder until the tensor-core kernels land; fails loudly.
#include "common.cuh"
extern "C" int64_t cmc_csd_workspace_bytes(int, int, int, int) { return 16; }
extern "C" int cmc_csd_msc(const float*, const float*, int, int, int, int, int64_t, int64_t, float*, float*,
                           float*, float*, void*, int64_t, void*) {
    cmc::set_error("cmc_csd_msc: not built yet");
    return CMC_EUNSUPPORTED;
}
extern "C" int64_t cmc_surrogate_workspace_bytes(int, int, int, int, int, int64_t) { return 16; }
extern "C" int cmc_surrogate_null(const void*, int, int, int, int, int, int, const int32_t*, uint64_t, int64_t,
                                  int64_t, const float*, uint32_t*, float*, void*, int64_t, void*) {
    cmc::set_error("cmc_surrogate_null: not built yet");
    return CMC_EUNSUPPORTED;
}
