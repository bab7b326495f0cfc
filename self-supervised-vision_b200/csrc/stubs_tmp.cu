// TEMPORARY: entry points not implemented yet return SSVB_ERR_UNSUPPORTED (removed as they land).
#include "host_util.h"
extern "C" {
size_t ssvb_barlow_saved_bytes(int64_t, int64_t) { return 0; }
size_t ssvb_barlow_workspace_bytes(int64_t, int64_t) { return 0; }
int ssvb_barlow_fwd(const float*, const float*, int64_t, int64_t, int64_t, int64_t, int, float, float*, void*, void*, size_t, void*) { return SSVB_ERR_UNSUPPORTED; }
int ssvb_barlow_bwd(const float*, const float*, int64_t, int64_t, int64_t, int64_t, int, float, const float*, const void*, float*, float*, int64_t, int64_t, void*, size_t, void*) { return SSVB_ERR_UNSUPPORTED; }
size_t ssvb_swav_saved_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }
size_t ssvb_swav_workspace_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }
int ssvb_swav_fwd(const float*, const float*, const float*, const float*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, float, float, int, float*, void*, void*, size_t, void*) { return SSVB_ERR_UNSUPPORTED; }
int ssvb_swav_bwd(const float*, const float*, const float*, const float*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, float, const float*, const void*, float*, float*, float*, int64_t, int64_t, int64_t, void*, size_t, void*) { return SSVB_ERR_UNSUPPORTED; }
}
