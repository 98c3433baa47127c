// extend.cu -- extension / consensus / period scan (placeholder until the kernels land)
#include "common.cuh"
using namespace bwtk;
#define NOTYET(name) do { set_error(name ": kernel not built yet"); return BWTK_EINTERNAL; } while (0)
extern "C" int32_t bwtk_extend_batch(const uint8_t *, int64_t, const int32_t *, const int32_t *, const int32_t *, int64_t, int32_t, int32_t *, void *) { NOTYET("extend_batch"); }
extern "C" int32_t bwtk_consensus_batch(const uint8_t *, int64_t, const int32_t *, const int32_t *, const int32_t *, const int64_t *, int64_t, uint8_t *, int32_t *, void *) { NOTYET("consensus_batch"); }
extern "C" int32_t bwtk_period_scan(const uint8_t *, int64_t, int64_t, int64_t, int32_t, int64_t, int64_t, double, const uint8_t *, const double *, int64_t, int32_t *, int64_t, int64_t *, int64_t *, void *) { NOTYET("period_scan"); }
