// Host-side plumbing of the C ABI: error reporting, version, workspace sizing.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include "tiles.cuh"

namespace b200ssl {
namespace {
thread_local char g_err[512] = "";
unsigned long long* g_timing = nullptr;
}
int g_head_sm_budget = 0;
unsigned long long* debug_timing_buffer(int kernel_tag) { return g_timing ? g_timing + (size_t)kernel_tag * kDebugRegion : nullptr; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}
}  // namespace b200ssl

using namespace b200ssl;

extern "C" int b200ssl_version(void) { return B200SSL_VERSION; }

extern "C" void b200ssl_debug_set_timing_buffer(void* device_u64) { g_timing = static_cast<unsigned long long*>(device_u64); }

extern "C" const char* b200ssl_last_error_string(void) { return g_err; }

extern "C" int b200ssl_set_head_sm_budget(int32_t sms) {
  const int prev = g_head_sm_budget;
  g_head_sm_budget = sms > 0 && sms < kNumSMs ? sms : 0;
  return prev;
}

extern "C" size_t b200ssl_workspace_bytes(int64_t rows, int32_t classes, int64_t bank_rows) {
  if (rows < 1) rows = 1;
  if (classes < 2) classes = 2;
  size_t need = sizeof(float) * 3 * kMaxRowCtas;                           // row-kernel partials
  const size_t da = sizeof(float) * (size_t)classes * kNumSMs;              // DA column partials
  if (da > need) need = da;
  const long long row_tiles = (rows + kTM - 1) / kTM;
  const size_t contrast = sizeof(float) * contrast_workspace_floats(rows, B200SSL_MAX_EMB_DIM);
  if (contrast > need) need = contrast;
  const size_t contrast_tc = sizeof(float) * contrast_tc_workspace_floats(rows);
  if (contrast_tc > need) need = contrast_tc;
  if (bank_rows > 0) {
    int tps = 0;
    const int nsplit = smooth_nsplit(rows, bank_rows, &tps);
    const size_t sm = nsplit > 1 ? (size_t)nsplit * row_tiles * kTM * ((1 + classes + 3) & ~3) * sizeof(float) : 0;
    if (sm > need) need = sm;
    const size_t tc = sizeof(float) * smooth_tc_workspace_floats(rows, bank_rows, classes);
    if (tc > need) need = tc;
    const size_t tc32 = smooth_tc_f32_workspace_bytes(rows, bank_rows, classes);     // fold partials + bf16 hi / mid operand copies
    if (classes <= 31 && tc32 > need) need = tc32;
  }
  return kWsHeaderBytes + ((need + 255) & ~(size_t)255);
}
