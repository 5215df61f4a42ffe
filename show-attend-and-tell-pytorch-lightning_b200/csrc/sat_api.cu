// Library-level plumbing of libsat_b200.so: version, error string, ABI self-description.
#include <stdarg.h>
#include <string.h>

#include "sat_common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_sat_launches = 0;

void sat_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

int sat_version(void) { return SAT_ABI_VERSION; }

const char* sat_last_error(void) { return g_err; }

unsigned long long sat_launch_count(void) { return g_sat_launches; }

int sat_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(SatDims);
    case 1: return (int)sizeof(SatWeights);
    case 2: return (int)sizeof(SatTrainBuffers);
    default: return -1;
  }
}

}  // extern "C"
