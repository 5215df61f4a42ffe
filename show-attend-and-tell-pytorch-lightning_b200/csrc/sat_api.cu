// Library-level plumbing of libsat_b200.so: version, error string, ABI self-description.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "sat_common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_sat_launches = 0;

void sat_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- optional per-kernel event timing (bench.py roofline) --------------------------------------
#include <vector>
int g_sat_prof_kind = 0;
static std::vector<cudaEvent_t> g_prof_events;
void sat_prof_mark(cudaStream_t st) {
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_prof_events.push_back(e);
}

// L2 persistence carve-out, opt-in with SAT_ATT_L2=1.  Measured on B200 at BASELINE configs[1] (annotations 51 MB of the
// 126 MB L2, pinned for the time loop): attention forward 20.6 vs 20.3 us, backward 25.1 vs 24.1 us, decoder step 2.30 vs
// 2.20 ms WITH vs without the window -- the kernels are latency-bound at this batch, not DRAM-bound, and the persisting
// carve-out takes L2 away from everything else.  Off by default.
size_t sat_l2_persist_limit() {
  static size_t limit[64];
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  dev &= 63;
  if (!done[dev]) {
    done[dev] = true;
    limit[dev] = 0;
    const char* e = getenv("SAT_ATT_L2");
    if (e != nullptr && atoi(e) != 0) {
      int maxp = 0, maxw = 0;
      cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
      cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      if (maxp > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp) == cudaSuccess)
        limit[dev] = (size_t)(maxp < maxw || maxw == 0 ? maxp : maxw);
      cudaGetLastError();
    }
  }
  return limit[dev];
}

extern "C" {

int sat_profile_begin(int kind) {
  for (auto e : g_prof_events) cudaEventDestroy(e);
  g_prof_events.clear();
  g_sat_prof_kind = kind;
  return 0;
}

int sat_profile_end(float* total_ms, int* count) {
  g_sat_prof_kind = 0;
  float tot = 0.0f;
  int n = 0;
  for (size_t i = 0; i + 1 < g_prof_events.size(); i += 2) {
    SAT_CUDA(cudaEventSynchronize(g_prof_events[i + 1]));
    float ms = 0.0f;
    SAT_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i], g_prof_events[i + 1]));
    tot += ms;
    ++n;
  }
  for (auto e : g_prof_events) cudaEventDestroy(e);
  g_prof_events.clear();
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  return 0;
}

int sat_version(void) { return SAT_ABI_VERSION; }

float sat_dropout_multiplier(float p, uint64_t seed, uint32_t stream, uint64_t idx) { return sat_dropout_scale(p, seed, stream, idx); }

const char* sat_last_error(void) { return g_err; }

unsigned long long sat_launch_count(void) { return g_sat_launches; }

int sat_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(SatDims);
    case 1: return (int)sizeof(SatWeights);
    case 2: return (int)sizeof(SatTrainBuffers);
    case 3: return (int)sizeof(SatDecodeBuffers);
    case 4: return (int)sizeof(SatMasterWeights);
    case 5: return (int)sizeof(SatParamGrads);
    default: return -1;
  }
}

}  // extern "C"
