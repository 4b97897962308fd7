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

// Diagnostics only (bench.py): event pairs around the kernels whose name matches the filter ("*" = every kernel).
// Off unless tgpb200_debug_time_kernel() was called; never active during a CUDA-graph capture (bench.py times its
// kernels in a separate eager pass).
constexpr int kMaxEv = 8192;
static char g_filter[128] = {0};
static cudaEvent_t g_ev[2][kMaxEv];
static const char* g_ev_name[kMaxEv];
static int g_ev_n = 0;
static int g_ev_made = 0;

bool timing_match(const char* name) {
  if (g_filter[0] == 0 || g_ev_n >= kMaxEv) return false;
  if (g_filter[0] == '*' || strstr(name, g_filter) != nullptr) {
    g_ev_name[g_ev_n] = name;  // kernel names are string literals
    return true;
  }
  return false;
}
void timing_begin(cudaStream_t st) {
  while (g_ev_made <= g_ev_n) {
    cudaEventCreate(&g_ev[0][g_ev_made]);
    cudaEventCreate(&g_ev[1][g_ev_made]);
    ++g_ev_made;
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
// Trace of the recorded launches in launch order: one "name<TAB>ms" line per launch.  Returns the number of bytes
// written (0-terminated, truncated to `cap`).
size_t tgpb200_debug_kernel_times(char* buf, size_t cap) {
  const int n = tgp::g_ev_n;
  if (!buf || cap == 0) return 0;
  size_t off = 0;
  buf[0] = 0;
  for (int i = 0; i < n; ++i) {
    cudaEventSynchronize(tgp::g_ev[1][i]);
    float ms = 0;
    cudaEventElapsedTime(&ms, tgp::g_ev[0][i], tgp::g_ev[1][i]);
    int w = snprintf(buf + off, cap - off, "%s\t%.6f\n", tgp::g_ev_name[i], ms);
    if (w < 0 || (size_t)w >= cap - off) break;
    off += (size_t)w;
  }
  return off;
}
}
