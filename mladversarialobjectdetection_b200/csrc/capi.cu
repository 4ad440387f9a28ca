// Error plumbing and small host utilities of libeotpatch (see include/eotpatch.h).
#include "eot_common.cuh"

#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace eot {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return EOT_ERR_CUDA;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  return n;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("EOT_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

static bool stage_times_enabled() {
  static const bool on = [] { const char* e = getenv("EOT_KERNEL_TIMES"); return e && e[0] == '1'; }();
  return on;
}

StageTimer::StageTimer(cudaStream_t stream, const char* w) : on(false), st(stream), what(w), n(0) {
  if (!stage_times_enabled()) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  on = true;
  mark("start");
}

void StageTimer::mark(const char* stage) {
  if (!on || n >= 16) return;
  cudaEventCreate(&ev[n]);
  cudaEventRecord(ev[n], st);
  names[n++] = stage;
}

StageTimer::~StageTimer() {
  if (!on) return;
  cudaStreamSynchronize(st);
  float total = 0.0f;
  cudaEventElapsedTime(&total, ev[0], ev[n - 1]);
  fprintf(stderr, "[eot] %s: %.1f us =", what, total * 1e3f);
  for (int i = 1; i < n; ++i) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
    fprintf(stderr, " %s %.1f", names[i], ms * 1e3f);
  }
  fprintf(stderr, "\n");
  for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
}

}  // namespace eot

extern "C" const char* eot_last_error(void) { return eot::g_err; }
extern "C" int eot_version(void) { return 101; }
extern "C" uint64_t eot_launch_count(void) { return eot::g_launches.load(std::memory_order_relaxed); }
