// error plumbing + version for the C-ABI
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace v1t {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { int phase; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;      // records of the current window
static std::vector<ProfRec> g_prof_pool; // recycled events

ProfScope::ProfScope(int phase, cudaStream_t stream) : slot(-1), st(stream) {
  if (!g_prof_on) return;
  ProfRec r;
  if (!g_prof_pool.empty()) { r = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else { cudaEventCreate(&r.a); cudaEventCreate(&r.b); }
  r.phase = phase;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}
}  // namespace v1t

extern "C" uint64_t v1t_launch_count(void) { return v1t::g_launches.load(); }
extern "C" int v1t_prof_enable(int on) { v1t::g_prof_on = on != 0; return V1T_OK; }
extern "C" int v1t_prof_reset(void) {
  for (auto& r : v1t::g_prof) v1t::g_prof_pool.push_back(r);
  v1t::g_prof.clear();
  return V1T_OK;
}
extern "C" int v1t_prof_read(int phase, float* total_ms, int* count) {
  V1T_CHECK_ARG(total_ms && count, "prof_read: null output");
  float tot = 0.f; int n = 0;
  for (auto& r : v1t::g_prof) {
    if (r.phase != phase) continue;
    V1T_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    V1T_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    tot += ms; ++n;
  }
  *total_ms = tot; *count = n;
  return V1T_OK;
}

extern "C" const char* v1t_last_error(void) { return v1t::g_err; }
extern "C" int v1t_version(void) { return 1; }

extern "C" int v1t_dropout_mask(float* out, int64_t n, uint64_t seed, uint32_t site, float p, void* stream) {
  V1T_CHECK_ARG(out && n >= 0 && p >= 0.f && p < 1.f, "dropout_mask: bad argument");
  if (n == 0) return V1T_OK;
  return v1t::dropout_mask(out, n, v1t::DropSpec{seed, site, p}, (cudaStream_t)stream);
}
