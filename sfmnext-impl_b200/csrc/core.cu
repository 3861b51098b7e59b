// Error reporting and device checks for libsqlx.
#include "common.cuh"

#include <atomic>

namespace sqlx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();  // clear the non-sticky launch error so later calls are not poisoned
    return SQLX_ECUDA;
  }
  return SQLX_OK;
}

}  // namespace sqlx

#include <mutex>
#include <set>
#include <utility>

namespace sqlx {
int ensure_dyn_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;      // (kernel, device) pairs already configured
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return check_launch("cudaGetDevice");
  std::lock_guard<std::mutex> lk(mu);
  const auto key = std::make_pair(kernel, dev);
  if (done.count(key)) return SQLX_OK;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %zu) failed on device %d: %s", bytes, dev,
              cudaGetErrorString(cudaGetLastError()));
    return SQLX_ECUDA;
  }
  done.insert(key);
  return SQLX_OK;
}
}  // namespace sqlx

extern "C" const char* sqlx_last_error(void) { return sqlx::g_err; }

extern "C" int sqlx_version(void) { return 100; }

/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
extern "C" unsigned long long sqlx_launch_count(void) { return sqlx::g_launches.load(); }

extern "C" int sqlx_device_ok(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    cudaGetLastError();
    sqlx::set_error("cudaGetDeviceProperties(%d) failed", device);
    return 0;
  }
  return prop.major == 10 ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// Per-kernel device timing (bench.py's roofline leg): CUDA events recorded on the launching stream
// around the named kernels while profiling is enabled.  Off by default: one relaxed atomic load per launch.
// ------------------------------------------------------------------------------------------------
#include <atomic>
#include <mutex>
#include <string.h>
#include <vector>

namespace sqlx {

namespace {
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
struct Pending {
  const char* name;
  cudaEvent_t a, b;
};
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_free_events;

cudaEvent_t get_event() {
  if (!g_free_events.empty()) {
    cudaEvent_t e = g_free_events.back();
    g_free_events.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

ProfScope::ProfScope(const char* name, cudaStream_t st) : idx_(-1), st_(st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return;  // events inside a graph capture would become graph nodes: skip
  }
  std::lock_guard<std::mutex> lk(g_prof_mu);
  Pending p{name, get_event(), get_event()};
  cudaEventRecord(p.a, st);
  idx_ = (int)g_pending.size();
  g_pending.push_back(p);
}

ProfScope::~ProfScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_pending[idx_].b, st_);
}

}  // namespace sqlx

extern "C" int sqlx_profile_enable(int on) {
  sqlx::g_prof_on.store(on ? 1 : 0);
  return SQLX_OK;
}

// Waits for every pending event pair, then writes one line per kernel name: "<name> <launches> <total_ms>\n".
// Returns the number of bytes written (excluding the terminating NUL), or a negative SQLX_E* code.
extern "C" int sqlx_profile_report(char* buf, size_t buf_bytes) {
  using namespace sqlx;
  SQLX_REQUIRE(buf && buf_bytes > 0, "NULL / empty buffer");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Row { const char* name; int count; double ms; };
  std::vector<Row> rows;
  for (const Pending& p : g_pending) {
    if (cudaEventSynchronize(p.b) != cudaSuccess) {
      set_error("cudaEventSynchronize: %s", cudaGetErrorString(cudaGetLastError()));
      return SQLX_ECUDA;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    bool found = false;
    for (Row& r : rows)
      if (strcmp(r.name, p.name) == 0) { r.count++; r.ms += ms; found = true; break; }
    if (!found) rows.push_back({p.name, 1, (double)ms});
    g_free_events.push_back(p.a);
    g_free_events.push_back(p.b);
  }
  g_pending.clear();
  size_t off = 0;
  buf[0] = 0;
  for (const Row& r : rows) {
    const int w = snprintf(buf + off, buf_bytes - off, "%s %d %.6f\n", r.name, r.count, r.ms);
    if (w < 0 || (size_t)w >= buf_bytes - off) break;
    off += (size_t)w;
  }
  return (int)off;
}
