// Library-level entry points: version, error strings, device check.
#include "host_util.h"

#include <atomic>
#include <mutex>

namespace ssvb {
namespace {
constexpr int kMaxRec = 8192;
std::atomic<long long> g_launches{0};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
cudaEvent_t g_ev0[kMaxRec], g_ev1[kMaxRec];
int g_kind[kMaxRec];
int g_nrec = 0, g_nalloc = 0;
}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int prof_begin(int kind, cudaStream_t s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_nrec >= kMaxRec) return -1;
  const int slot = g_nrec++;
  if (slot >= g_nalloc) {
    cudaEventCreate(&g_ev0[slot]);
    cudaEventCreate(&g_ev1[slot]);
    g_nalloc = slot + 1;
  }
  g_kind[slot] = kind;
  cudaEventRecord(g_ev0[slot], s);
  return slot;
}
void prof_end(int slot, cudaStream_t s) {
  if (slot >= 0) cudaEventRecord(g_ev1[slot], s);
}
}  // namespace ssvb

extern "C" {

// ---- measurement hooks for bench.py (not part of the loss path; OFF by default) -------------------------
long long ssvb_launch_count(int reset) {
  return reset ? ssvb::g_launches.exchange(0) : ssvb::g_launches.load();
}
int ssvb_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(ssvb::g_prof_mu);
  ssvb::g_nrec = 0;
  ssvb::g_prof_on.store(on ? 1 : 0);
  return SSVB_OK;
}
// total device milliseconds and launch count of the kernels of `kind` recorded since profile_enable(1);
// synchronises on the recorded events.
int ssvb_profile_summary(int kind, double* total_ms, long long* count) {
  if (!total_ms || !count) return SSVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ssvb::g_prof_mu);
  double tot = 0.0;
  long long n = 0;
  for (int i = 0; i < ssvb::g_nrec; ++i) {
    if (ssvb::g_kind[i] != kind) continue;
    if (cudaEventSynchronize(ssvb::g_ev1[i]) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ssvb::g_ev0[i], ssvb::g_ev1[i]) == cudaSuccess) {
      tot += ms;
      ++n;
    }
  }
  *total_ms = tot;
  *count = n;
  return SSVB_OK;
}

int ssvb_version(void) { return SSVB_VERSION; }

const char* ssvb_strerror(int rc) {
  switch (rc) {
    case SSVB_OK: return "ok";
    case SSVB_ERR_INVALID: return "ssv_b200: invalid argument (null pointer, non-positive size or bad flag)";
    case SSVB_ERR_ALIGNMENT: return "ssv_b200: pointer / leading dimension not 16-byte aligned (need d % 4 == 0)";
    case SSVB_ERR_UNSUPPORTED: return "ssv_b200: shape not supported by the sm_100a kernels";
    case SSVB_ERR_WORKSPACE: return "ssv_b200: workspace / saved buffer too small";
    case SSVB_ERR_ARCH: return "ssv_b200: current device is not a Blackwell sm_100 GPU (no fallback path exists)";
    case SSVB_ERR_DRIVER: return "ssv_b200: cuTensorMapEncodeTiled unavailable or failed";
    default: break;
  }
  if (rc > 0) return cudaGetErrorString(static_cast<cudaError_t>(rc));
  return "ssv_b200: unknown error";
}

int ssvb_device_check(void) { return ssvb::check_device_sm100(); }

}  // extern "C"
