#include "host_util.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace ovla {

static thread_local char g_err[1024] = "";
static long long g_launches = 0;

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
const char* last_error() { return g_err; }
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches; }
void reset_launch_count() { g_launches = 0; }


namespace {
struct ProfRec {
  int cat;
  double flops, bytes;
  cudaEvent_t a, b;
};
bool g_prof_on = false;
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_free;
cudaEvent_t get_event() {
  if (!g_free.empty()) {
    cudaEvent_t e = g_free.back();
    g_free.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OVLA_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

void prof_enable(bool on) { g_prof_on = on; }
bool prof_enabled() { return g_prof_on; }
void prof_begin(int cat, double flops, double bytes, cudaStream_t st) {
  ProfRec r{cat, flops, bytes, get_event(), get_event()};
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, st);
}
int prof_collect(long long* launches, double* ms, double* flops, double* bytes) {
  for (int i = 0; i < kNumCat; ++i) { launches[i] = 0; ms[i] = flops[i] = bytes[i] = 0.0; }
  for (ProfRec& r : g_recs) {
    CUDA_TRY(cudaEventSynchronize(r.b));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
    launches[r.cat] += 1;
    ms[r.cat] += t;
    flops[r.cat] += r.flops;
    bytes[r.cat] += r.bytes;
    g_free.push_back(r.a);
    g_free.push_back(r.b);
  }
  g_recs.clear();
  return 0;
}

}  // namespace ovla
