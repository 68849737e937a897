// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's
// roofline.achieved is "algorithmic work / kernel duration, measured live with CUDA events").
// Disabled by default: no events are recorded and the hot path pays one branch per launch.
#include "common.cuh"
#include <vector>

namespace updes {

struct ProfRecord { cudaEvent_t a, b; double work; int cat; };
static bool g_prof_on = false;
static std::vector<ProfRecord> g_records;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}

void prof_begin(int cat, double work, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfRecord r; r.a = get_event(); r.b = get_event(); r.work = work; r.cat = cat;
  cudaEventRecord(r.a, st);
  g_records.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!g_prof_on || g_records.empty()) return;
  cudaEventRecord(g_records.back().b, st);
}

}  // namespace updes

extern "C" int updes_profile_enable(int on) {
  using namespace updes;
  g_prof_on = on != 0;
  for (auto &r : g_records) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
  g_records.clear();
  return 0;
}

// Sum over the records of class `cat` since the last enable: milliseconds, work units, launches.
// Synchronises on the recorded events.
extern "C" int updes_profile_read(int cat, double *ms, double *work, int64_t *count) {
  using namespace updes;
  double tms = 0, tw = 0; int64_t c = 0;
  for (auto &r : g_records) {
    if (r.cat != cat) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return (int)cudaGetLastError();
    float f = 0; cudaEventElapsedTime(&f, r.a, r.b);
    tms += f; tw += r.work; c++;
  }
  if (ms) *ms = tms;
  if (work) *work = tw;
  if (count) *count = c;
  return 0;
}

// Per-launch records of class `cat` since the last enable, in launch order: ms[i], work[i] for i < min(count, max).
// Returns the number of records of that class (or a negative CUDA error).
extern "C" int64_t updes_profile_records(int cat, double *ms, double *work, int64_t max) {
  using namespace updes;
  int64_t c = 0;
  for (auto &r : g_records) {
    if (r.cat != cat) continue;
    if (c < max) {
      if (cudaEventSynchronize(r.b) != cudaSuccess) return -(int64_t)cudaGetLastError();
      float f = 0; cudaEventElapsedTime(&f, r.a, r.b);
      if (ms) ms[c] = f;
      if (work) work[c] = r.work;
    }
    c++;
  }
  return c;
}
