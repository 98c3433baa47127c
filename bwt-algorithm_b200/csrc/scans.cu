// scans.cu -- detectors (placeholder until the kernels land)
#include "common.cuh"
using namespace bwtk;
#define NOTYET(name) do { set_error(name ": kernel not built yet"); return BWTK_EINTERNAL; } while (0)
extern "C" int64_t bwtk_tier1_workspace_bytes(int64_t) { return 0; }
extern "C" int32_t bwtk_tier1_scan(const uint8_t *, int64_t, int32_t, int32_t, int32_t, double, int32_t *, int64_t, int64_t *, uint8_t *, void *, int64_t, void *) { NOTYET("tier1_scan"); }
extern "C" int64_t bwtk_strict_workspace_bytes(int64_t, int64_t) { return 0; }
extern "C" int32_t bwtk_strict_scan(const uint8_t *, int64_t, int64_t, int64_t, int64_t, int64_t, int32_t *, int64_t, int64_t *, void *, int64_t, void *) { NOTYET("strict_scan"); }
extern "C" int64_t bwtk_plateau_workspace_bytes(int64_t) { return 0; }
extern "C" int32_t bwtk_lcp_plateaus(const uint8_t *, int64_t, const int32_t *, const int32_t *, int64_t, int64_t, int64_t, int64_t, int32_t *, int64_t, int64_t *, int64_t *, void *, int64_t, void *) { NOTYET("lcp_plateaus"); }
