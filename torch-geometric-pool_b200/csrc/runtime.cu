// Process-level helpers of the library: launch counter, debug-sync switch.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace tgp {

long long& launch_counter() {
  static long long n = 0;
  return n;
}

bool debug_sync_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TGPB200_DEBUG_SYNC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

void report_launch_failure(const char* name, cudaError_t e) {
  fprintf(stderr, "[tgp_b200] kernel %s failed: %s\n", name, cudaGetErrorString(e));
  fflush(stderr);
}

}  // namespace tgp

extern "C" {
long long tgpb200_debug_launch_count(void) { return tgp::launch_counter(); }
}
