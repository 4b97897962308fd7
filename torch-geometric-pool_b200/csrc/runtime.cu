// Process-level helpers of the library: launch counter, debug-sync switch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

static char g_filter[128] = {0};
static cudaEvent_t g_ev[2][512];
static int g_ev_n = 0;
static bool g_ev_init = false;

bool timing_match(const char* name) { return g_filter[0] != 0 && strstr(name, g_filter) != nullptr && g_ev_n < 512; }
void timing_begin(cudaStream_t st) {
  if (!g_ev_init) {
    for (int i = 0; i < 512; ++i) {
      cudaEventCreate(&g_ev[0][i]);
      cudaEventCreate(&g_ev[1][i]);
    }
    g_ev_init = true;
  }
  cudaEventRecord(g_ev[0][g_ev_n], st);
}
void timing_end(cudaStream_t st) {
  cudaEventRecord(g_ev[1][g_ev_n], st);
  ++g_ev_n;
}

}  // namespace tgp

extern "C" {
long long tgpb200_debug_launch_count(void) { return tgp::launch_counter(); }

// Start timing every kernel whose name contains `filter` (NULL or "" stops); resets the recorded pairs.
void tgpb200_debug_time_kernel(const char* filter) {
  tgp::g_ev_n = 0;
  if (filter == nullptr) {
    tgp::g_filter[0] = 0;
    return;
  }
  strncpy(tgp::g_filter, filter, sizeof(tgp::g_filter) - 1);
}
// Mean duration (ms) of the recorded launches (synchronises on their events); *count receives how many.
double tgpb200_debug_kernel_time_ms(int* count) {
  double tot = 0;
  int n = tgp::g_ev_n;
  for (int i = 0; i < n; ++i) {
    cudaEventSynchronize(tgp::g_ev[1][i]);
    float ms = 0;
    cudaEventElapsedTime(&ms, tgp::g_ev[0][i], tgp::g_ev[1][i]);
    tot += ms;
  }
  if (count) *count = n;
  return n > 0 ? tot / n : 0.0;
}
}
