// rs_api.cu — the C ABI declared in include/rag_b200.h: argument validation, kernel-family
// choice, workspace ownership and status/last-error plumbing.  No exceptions cross the ABI.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/rag_b200.h"
#include "comm.h"
#include "kernels.h"
#include "tc5_host.h"

namespace {
thread_local std::string g_create_error;
}

// Worker threads that stage a host-resident document list into pinned memory (rs_maxsim_list).  They live as long as
// the handle: creating threads per call and binding each to the device cost ~0.1 ms of a 0.7 ms call, and made six
// threads slower than three.  A worker spins on the job counter for a short while after a job (the next rerank call
// usually follows at once) and then parks on a condition variable.
class StagePool {
 public:
  StagePool(int workers, int device) : n_(workers) {
    for (int w = 0; w < workers; ++w) threads_.emplace_back([this, w, device] { loop(w, device); });
  }
  ~StagePool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      gen_.fetch_add(1);
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int workers() const { return n_; }
  // runs job(0..workers) — part `workers` on the calling thread — and returns when every part is done
  void run(const std::function<void(int)>& job) {
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &job;
      pending_.store(n_);
      gen_.fetch_add(1);
    }
    cv_.notify_all();
    job(n_);
    while (pending_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  }

 private:
  void loop(int w, int device) {
    cudaSetDevice(device);
    uint64_t seen = 0;
    for (;;) {
      // hot phase: the next job usually arrives within microseconds of the last
      const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(300);
      while (gen_.load(std::memory_order_acquire) == seen && std::chrono::steady_clock::now() < until) {
      }
      if (gen_.load(std::memory_order_acquire) == seen) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_.load() != seen; });
      }
      const std::function<void(int)>* job;
      {
        std::lock_guard<std::mutex> lk(m_);
        seen = gen_.load();
        if (stop_) return;
        job = job_;
      }
      (*job)(w);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  int n_;
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> pending_{0};
  const std::function<void(int)>* job_ = nullptr;
  bool stop_ = false;
};

struct rs_handle {
  int device = 0;
  int num_sms = 0;
  // Stream ordering: every launch on this handle shares the workspaces below, so consecutive calls must not overlap
  // on the device.  Calls on ONE stream are ordered by the stream; when the stream changes, the new stream waits for
  // an event recorded on the previous one (order_after_last).
  cudaStream_t last_stream = nullptr;
  // row tensor map of the last filtered scan's corpus (tile::gather4, dense_scan.cu)
  alignas(64) CUtensorMap gmap;
  const void* gmap_corpus = nullptr;
  int64_t gmap_n = -1;
  int32_t gmap_d = -1;
  bool gmap_ok = false;
  int last_dense_redo = 0;  // queries the last batched k > 128 call re-ran through the scan
  StagePool* stage_pool = nullptr;  // created by the first rs_maxsim_list call with a large host-resident list
  bool has_last = false;
  cudaEvent_t order_ev = nullptr;
  // dense scan workspace
  uint64_t* ws_keys = nullptr;  // 2 x [num_sms, 2048]: consecutive scans alternate (they may overlap under PDL)
  unsigned* ticket = nullptr;   // 2 counters, 128 bytes apart
  uint64_t scan_seq = 0;
  // scan work counters: launch i uses counter (i mod kScanCounters), which the launch itself leaves at zero
  unsigned long long* unit_ctr = nullptr;
  uint64_t* scan_trace = nullptr;  // diagnostics: caller's device buffer [8][num_sms][8], see rs_set_scan_trace
  // *_host staging
  void* pinned = nullptr;      // mapped pinned memory: staged inputs, and results written by the kernel itself
  void* pinned_dev = nullptr;  // device alias of `pinned`
  size_t pinned_bytes = 0;
  void* dev_stage = nullptr;
  size_t dev_stage_bytes = 0;
  // packed tokens of rs_maxsim_list (documents + query in the compute dtype)
  void* list_buf = nullptr;
  size_t list_buf_bytes = 0;
  // small device scratch for rs_filter_mask (values, offsets)
  int32_t* filt_dev = nullptr;
  size_t filt_dev_ints = 0;
  // tcgen05 paths (tensor-map cache, scratch)
  rs::Tc5State* tc5 = nullptr;
  // multi-GPU exchange over peer memory (comm.cu)
  rs::CommState* comm = nullptr;
  float* shard_scores = nullptr;  // local top-k of rs_dense_topk_sharded_host before the exchange
  int64_t* shard_ids = nullptr;
  size_t shard_pairs = 0;
  // per-call statistics (rs_set_profiling)
  bool profiling = false;
  cudaEvent_t prof_ev[3] = {nullptr, nullptr, nullptr};  // start, before the merge launch, end
  bool prof_mid = false, prof_pending = false;
  rs_call_stats prof{};
  int dense_impl = RS_DENSE_AUTO, maxsim_impl = RS_MAXSIM_AUTO;
  int last_dense_impl = 0, last_maxsim_impl = 0;
  int64_t launches = 0;
  std::string err;
};

namespace {

constexpr int kMaxK = 2048;
// A launch that has started but not completed keeps at least one CTA resident, and at most 148 SMs x 4 CTAs
// (448 threads each) fit on the device, so launches i and i + 1024 are never in flight together: by the time a
// counter is used again its previous launch has completed, i.e. its merging CTA has reset it.
constexpr int kScanCounters = 1024;

int fail(rs_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_create_error = buf;
  return code;
}

int cuda_fail(rs_handle* h, cudaError_t e, const char* what) {
  return fail(h, RS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Order the work about to be enqueued on `st` after everything this handle enqueued before.  Same stream: nothing to
// do.  Different stream: event on the previous stream, wait on the new one.  (A previous stream the caller has
// destroyed in the meantime has no pending work left to wait for; the failed record is ignored.)
void order_after_last(rs_handle* h, cudaStream_t st) {
  if (h->has_last && h->last_stream != st) {
    if (cudaEventRecord(h->order_ev, h->last_stream) == cudaSuccess)
      cudaStreamWaitEvent(st, h->order_ev, 0);
    else
      cudaGetLastError();
  }
  h->last_stream = st;
  h->has_last = true;
}

int ensure_staging(rs_handle* h, size_t host_bytes, size_t dev_bytes) {
  if (host_bytes > h->pinned_bytes) {
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr;
    h->pinned_bytes = 0;
    size_t want = host_bytes < (1u << 20) ? (1u << 20) : host_bytes;
    cudaError_t e = cudaHostAlloc(&h->pinned, want, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->pinned_dev, h->pinned, 0);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaHostAlloc(staging)");
    h->pinned_bytes = want;
  }
  if (dev_bytes > h->dev_stage_bytes) {
    if (h->dev_stage) cudaFree(h->dev_stage);
    h->dev_stage = nullptr;
    h->dev_stage_bytes = 0;
    size_t want = dev_bytes < (1u << 20) ? (1u << 20) : dev_bytes;
    cudaError_t e = cudaMalloc(&h->dev_stage, want);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc(staging)");
    h->dev_stage_bytes = want;
  }
  return RS_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bracket of one profiled call: the constructor records the start event, finish() the end event and the counters.
struct ProfScope {
  rs_handle* h;
  cudaStream_t st;
  int64_t launches0;
  bool on;
  ProfScope(rs_handle* handle, cudaStream_t stream) : h(handle), st(stream), launches0(handle->launches), on(handle->profiling) {
    if (on) {
      cudaEventRecord(h->prof_ev[0], st);
      h->prof_mid = false;
    }
  }
  void finish(int entry, int family, int queries, int64_t bytes, double flops) {
    if (!on) return;
    cudaEventRecord(h->prof_ev[2], st);
    h->prof = rs_call_stats{};
    h->prof.entry = entry;
    h->prof.kernel_family = family;
    h->prof.launches = (int32_t)(h->launches - launches0);
    h->prof.queries = queries;
    h->prof.bytes_scanned = bytes;
    h->prof.flops = flops;
    h->prof_pending = true;
  }
};

}  // namespace

extern "C" {

int rs_abi_version(void) { return RS_ABI_VERSION; }

int rs_create(int device, rs_handle** out) {
  if (!out) return fail(nullptr, RS_ERR_INVALID_ARG, "rs_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return fail(nullptr, RS_ERR_NO_DEVICE,
                "rs_create: no CUDA device visible (%s); this engine has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= count) return fail(nullptr, RS_ERR_INVALID_ARG, "rs_create: device %d out of range", device);
  cudaDeviceProp prop{};
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, RS_ERR_NO_DEVICE,
                "rs_create: device %d is sm_%d%d; the kernels are built for sm_100a only and there is no fallback",
                device, prop.major, prop.minor);
  DeviceGuard guard(device);
  rs_handle* h = new (std::nothrow) rs_handle();
  if (!h) return fail(nullptr, RS_ERR_NOMEM, "rs_create: out of host memory");
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  e = cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&h->ws_keys, (size_t)2 * h->num_sms * kMaxK * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->ticket, 256);
  if (e == cudaSuccess) e = cudaMemset(h->ticket, 0, 256);
  if (e == cudaSuccess) e = cudaMalloc(&h->unit_ctr, kScanCounters * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->unit_ctr, 0, kScanCounters * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    int rc = cuda_fail(nullptr, e, "rs_create: workspace allocation");
    rs_destroy(h);
    return rc;
  }
  h->tc5 = rs::tc5_create(device, h->num_sms);
  h->comm = rs::comm_create(device, h->num_sms);
  *out = h;
  return RS_OK;
}

int rs_destroy(rs_handle* h) {
  if (!h) return RS_OK;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  delete h->stage_pool;
  if (h->tc5) rs::tc5_destroy(h->tc5);
  if (h->comm) rs::comm_destroy(h->comm);
  if (h->shard_scores) cudaFree(h->shard_scores);
  if (h->shard_ids) cudaFree(h->shard_ids);
  if (h->ws_keys) cudaFree(h->ws_keys);
  if (h->ticket) cudaFree(h->ticket);
  if (h->unit_ctr) cudaFree(h->unit_ctr);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->dev_stage) cudaFree(h->dev_stage);
  if (h->filt_dev) cudaFree(h->filt_dev);
  if (h->list_buf) cudaFree(h->list_buf);
  if (h->order_ev) cudaEventDestroy(h->order_ev);
  for (cudaEvent_t ev : h->prof_ev)
    if (ev) cudaEventDestroy(ev);
  delete h;
  return RS_OK;
}

const char* rs_last_error(const rs_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int64_t rs_launch_count(const rs_handle* h) { return h ? h->launches : 0; }

int rs_set_dense_impl(rs_handle* h, int impl) {
  if (!h || impl < RS_DENSE_AUTO || impl > RS_DENSE_TCGEN05) return fail(h, RS_ERR_INVALID_ARG, "rs_set_dense_impl: bad argument");
  h->dense_impl = impl;
  return RS_OK;
}
int rs_set_maxsim_impl(rs_handle* h, int impl) {
  if (!h || impl < RS_MAXSIM_AUTO || impl > RS_MAXSIM_TCGEN05_CAND) return fail(h, RS_ERR_INVALID_ARG, "rs_set_maxsim_impl: bad argument");
  h->maxsim_impl = impl;
  return RS_OK;
}
int rs_set_scan_trace(rs_handle* h, uint64_t* trace_dev) {
  if (!h) return fail(h, RS_ERR_INVALID_ARG, "rs_set_scan_trace: null handle");
  h->scan_trace = trace_dev;
  return RS_OK;
}
int rs_scan_plan(int32_t d, int32_t k, int64_t* out7) {
  if (!out7 || d <= 0 || (d % 8) != 0 || d > 4096 || k < 1 || k > kMaxK) return RS_ERR_INVALID_ARG;
  rs::scan_plan_query(d, k, out7);
  return RS_OK;
}
int rs_scan_plan_chained(int32_t d, int32_t k, int64_t* out8) {
  if (!out8 || d <= 0 || (d % 8) != 0 || d > 4096 || k < 1 || k > kMaxK) return RS_ERR_INVALID_ARG;
  rs::scan_plan_query(d, k, out8, /*chained=*/true);
  return RS_OK;
}
int rs_set_profiling(rs_handle* h, int on) {
  if (!h) return RS_ERR_INVALID_ARG;
  DeviceGuard guard(h->device);
  if (on && !h->prof_ev[0]) {
    for (cudaEvent_t& ev : h->prof_ev) {
      cudaError_t e = cudaEventCreate(&ev);
      if (e != cudaSuccess) return cuda_fail(h, e, "rs_set_profiling: cudaEventCreate");
    }
  }
  h->profiling = on != 0;
  h->prof_pending = false;
  return RS_OK;
}
int rs_last_call_stats(rs_handle* h, rs_call_stats* out) {
  if (!h || !out) return RS_ERR_INVALID_ARG;
  if (!h->prof_pending) return fail(h, RS_ERR_INVALID_ARG, "rs_last_call_stats: no profiled call (rs_set_profiling(h, 1) first)");
  DeviceGuard guard(h->device);
  cudaError_t e = cudaEventSynchronize(h->prof_ev[2]);
  if (e != cudaSuccess) return cuda_fail(h, e, "rs_last_call_stats: cudaEventSynchronize");
  float ms = 0.f, mm = 0.f;
  cudaEventElapsedTime(&ms, h->prof_ev[0], h->prof_ev[2]);
  if (h->prof_mid) cudaEventElapsedTime(&mm, h->prof_ev[1], h->prof_ev[2]);
  h->prof.device_ms = ms;
  h->prof.merge_ms = mm;
  *out = h->prof;
  return RS_OK;
}
int rs_last_dense_impl(const rs_handle* h) { return h ? h->last_dense_impl : 0; }
int rs_last_dense_redo(const rs_handle* h) { return h ? h->last_dense_redo : 0; }
int rs_last_maxsim_impl(const rs_handle* h) { return h ? h->last_maxsim_impl : 0; }

// ------------------------------------------------------------------------------ dense
int rs_dense_topk(rs_handle* h, const void* corpus, int64_t n, int32_t d, int32_t dtype, const float* inv_norm,
                  int32_t metric, const void* queries, int32_t nq, const uint32_t* mask, int64_t mask_stride_words,
                  int32_t k, int64_t id_base, float* out_scores, int64_t* out_ids, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (n < 0 || nq < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk: negative size (n=%lld nq=%d)", (long long)n, nq);
  if (nq == 0) return RS_OK;
  if (!queries || !out_scores || !out_ids) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk: NULL queries/outputs");
  if (n > 0 && !corpus) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk: NULL corpus");
  if (dtype != RS_F16 && dtype != RS_BF16)
    return fail(h, RS_ERR_UNSUPPORTED, "rs_dense_topk: corpus dtype must be RS_F16 or RS_BF16 (got %d)", dtype);
  if (metric != RS_METRIC_IP && metric != RS_METRIC_COSINE) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk: bad metric %d", metric);
  if (d <= 0 || (d % 8) != 0 || d > 4096)
    return fail(h, RS_ERR_UNSUPPORTED, "rs_dense_topk: d must be a multiple of 8 in [8, 4096] (got %d)", d);
  if (k < 1 || k > kMaxK) return fail(h, RS_ERR_UNSUPPORTED, "rs_dense_topk: k must be in [1, %d] (got %d)", kMaxK, k);
  if (n >= (1ll << 32)) return fail(h, RS_ERR_UNSUPPORTED, "rs_dense_topk: n must be < 2^32 rows per shard");
  if (!aligned16(corpus) || !aligned16(queries))
    return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk: corpus and queries must be 16-byte aligned");
  if (mask_stride_words < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk: negative mask stride");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  order_after_last(h, st);
  ProfScope prof(h, st);
  const int64_t scan_bytes = n * (int64_t)d * 2;

  // one single-query scan launch (query qi of the call); `chained`: it is one of several launches of this call
  const CUtensorMap* gather_map = nullptr;
  auto scan_one = [&](int qi, bool pdl, bool chained) -> int {
    rs::ScanParams p{};
    p.corpus = corpus;
    p.query = static_cast<const uint8_t*>(queries) + (size_t)qi * d * 2;
    p.inv_norm = (metric == RS_METRIC_COSINE) ? inv_norm : nullptr;
    p.mask = mask ? mask + (size_t)qi * mask_stride_words : nullptr;
    p.n = n;
    p.d = d;
    p.k = k;
    p.metric = metric;
    p.id_base = id_base;
    if (h->scan_trace) p.trace = h->scan_trace + (h->scan_seq & 7) * (size_t)h->num_sms * 8;
    const int buf = (int)(h->scan_seq++ & 1);
    p.ws_keys = h->ws_keys + (size_t)buf * h->num_sms * kMaxK;
    p.ticket = h->ticket + buf * 32;
    p.out_scores = out_scores + (size_t)qi * k;
    p.out_ids = out_ids + (size_t)qi * k;
    const int ctr = (int)((h->scan_seq - 1) & (kScanCounters - 1));
    p.unit_counter = h->unit_ctr + ctr;
    cudaError_t e = rs::launch_dense_scan(p, dtype, h->num_sms, pdl, st, gather_map, chained);
    if (e != cudaSuccess) return cuda_fail(h, e, "dense_scan_kernel launch");
    h->launches += 1;
    return RS_OK;
  };
  // Filtered scans fetch the passing rows of sparse mask words four per TMA instruction (tile::gather4) through a
  // row tensor map of the corpus; encoded on the host (~1 us) when the corpus changes, kept in the handle.
  auto prepare_gather_map = [&]() {
    static const bool no_gather4 = getenv("RS_SCAN_NO_GATHER4") != nullptr;
    if (mask != nullptr && n > 0 && !no_gather4 && rs::scan_gather4_supported(d)) {
      if (h->gmap_corpus != corpus || h->gmap_n != n || h->gmap_d != d) {
        std::string err;
        h->gmap_ok = rs::tc5_encode_rows(h->tc5, &h->gmap, corpus, (uint64_t)n, (uint32_t)d * 2u, &err);
        h->gmap_corpus = corpus;
        h->gmap_n = n;
        h->gmap_d = d;
      }
      if (h->gmap_ok) gather_map = &h->gmap;
    }
  };

  int impl = h->dense_impl;
  if (impl == RS_DENSE_AUTO)
    impl = rs::tc5_dense_supported(h->tc5, n, d, nq, k, mask, mask_stride_words, /*worthwhile=*/true) ? RS_DENSE_TCGEN05 : RS_DENSE_SCAN;
  if (impl == RS_DENSE_TCGEN05) {
    if (!rs::tc5_dense_supported(h->tc5, n, d, nq, k, mask, mask_stride_words))
      return fail(h, RS_ERR_UNSUPPORTED,
                  "rs_dense_topk: tcgen05 batched path needs nq >= 2, n >= 256, d %% 64 == 0, k <= 128 (k <= 1024 when the "
                  "corpus ranges can hold twice their share of the answer)");
    int launched = 0;
    std::string err;
    std::vector<int> redo;
    int rc = rs::tc5_dense_topk(h->tc5, corpus, n, d, dtype, inv_norm, metric, queries, nq, mask, mask_stride_words, k, id_base, out_scores,
                                out_ids, st, &launched, &err, &redo);
    h->launches += launched;
    h->last_dense_impl = RS_DENSE_TCGEN05;
    if (rc != RS_OK) return fail(h, rc, "rs_dense_topk(tcgen05): %s", err.c_str());
    // k > 128: queries for which a corpus range may have held more of the answer than its list keeps are re-run
    // through the exact single-query scan (dense_tc5.cu, dense_overflow_check_kernel)
    h->last_dense_redo = (int)redo.size();
    if (!redo.empty()) {
      prepare_gather_map();
      for (size_t i = 0; i < redo.size(); ++i) {
        rc = scan_one(redo[i], /*pdl=*/i > 0, /*chained=*/redo.size() > 1);
        if (rc != RS_OK) return rc;
      }
    }
    prof.finish(RS_CALL_DENSE_TOPK, RS_DENSE_TCGEN05, nq, scan_bytes, 2.0 * nq * (double)n * d);
    return RS_OK;
  }

  h->last_dense_impl = RS_DENSE_SCAN;
  prepare_gather_map();
  for (int qi = 0; qi < nq; ++qi) {
    int rc = scan_one(qi, /*pdl=*/qi > 0, /*chained=*/nq > 1);
    if (rc != RS_OK) return rc;
  }
  prof.finish(RS_CALL_DENSE_TOPK, RS_DENSE_SCAN, nq, scan_bytes * nq, 2.0 * nq * (double)n * d);
  return RS_OK;
}

int rs_dense_topk_host(rs_handle* h, const void* corpus, int64_t n, int32_t d, int32_t dtype, const float* inv_norm,
                       int32_t metric, const void* queries_host, int32_t nq, const uint32_t* mask_host,
                       const uint32_t* mask_dev, int64_t mask_stride_words, int32_t k, int64_t id_base,
                       float* out_scores_host, int64_t* out_ids_host, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (nq < 0 || n < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk_host: negative size");
  if (nq == 0) return RS_OK;
  if (!queries_host || !out_scores_host || !out_ids_host) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk_host: NULL host buffer");
  if (d <= 0 || k < 1 || k > kMaxK) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk_host: bad d/k");
  DeviceGuard guard(h->device);
  // Everything runs on the CALLER's stream: the copy of the queries, the scan and the final synchronise are ordered
  // after whatever the caller enqueued there before (rs_filter_mask building mask_dev, writes to the corpus).
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t q_bytes = align_up((size_t)nq * d * 2, 256);
  const size_t words_per_mask = (size_t)((n + 31) / 32);
  const size_t n_masks = mask_host ? (mask_stride_words ? (size_t)nq : 1) : 0;
  const size_t m_bytes = align_up(mask_host ? ((n_masks - 1) * (size_t)mask_stride_words + words_per_mask) * 4 : 0, 256);
  const size_t os_bytes = align_up((size_t)nq * k * 4, 256);
  const size_t oi_bytes = align_up((size_t)nq * k * 8, 256);
  const size_t total = q_bytes + m_bytes + os_bytes + oi_bytes;
  int rc = ensure_staging(h, total, total);
  if (rc != RS_OK) return rc;
  uint8_t* hp = static_cast<uint8_t*>(h->pinned);
  uint8_t* dp = static_cast<uint8_t*>(h->dev_stage);
  memcpy(hp, queries_host, (size_t)nq * d * 2);
  if (mask_host) memcpy(hp + q_bytes, mask_host, ((n_masks - 1) * (size_t)mask_stride_words + words_per_mask) * 4);
  cudaError_t e = cudaMemcpyAsync(dp, hp, q_bytes + m_bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(h, e, "H2D(queries, mask)");
  const uint32_t* mask = mask_host ? reinterpret_cast<const uint32_t*>(dp + q_bytes) : mask_dev;
  // The k result pairs are written by the kernel straight into mapped pinned memory: no device-to-host copy to
  // enqueue and wait for, the stream synchronise below is the only round trip after the launch.
  uint8_t* hp_dev = static_cast<uint8_t*>(h->pinned_dev);
  float* d_scores = reinterpret_cast<float*>(hp_dev + q_bytes + m_bytes);
  int64_t* d_ids = reinterpret_cast<int64_t*>(hp_dev + q_bytes + m_bytes + os_bytes);
  rc = rs_dense_topk(h, corpus, n, d, dtype, inv_norm, metric, dp, nq, mask, mask_stride_words, k, id_base, d_scores,
                     d_ids, stream);
  if (rc != RS_OK) return rc;
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(h, e, "rs_dense_topk_host: stream synchronize");
  memcpy(out_scores_host, hp + q_bytes + m_bytes, (size_t)nq * k * 4);
  memcpy(out_ids_host, hp + q_bytes + m_bytes + os_bytes, (size_t)nq * k * 8);
  return RS_OK;
}

int rs_topk_merge(rs_handle* h, const float* scores, const int64_t* ids, int32_t nlists, int32_t nq, int32_t k_in,
                  int32_t k_out, int64_t score_list_stride, int64_t id_list_stride, float* out_scores, int64_t* out_ids,
                  void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (nq == 0) return RS_OK;
  if (!scores || !ids || !out_scores || !out_ids) return fail(h, RS_ERR_INVALID_ARG, "rs_topk_merge: NULL buffer");
  if (nlists < 1 || nq < 0 || k_in < 1 || k_out < 1 || score_list_stride < 0 || id_list_stride < 0)
    return fail(h, RS_ERR_INVALID_ARG, "rs_topk_merge: bad sizes");
  if ((long long)nlists * k_in > 16384 || k_out > 16384)
    return fail(h, RS_ERR_UNSUPPORTED, "rs_topk_merge: nlists * k_in must be <= 16384 (got %lld)", (long long)nlists * k_in);
  DeviceGuard guard(h->device);
  cudaError_t e = rs::launch_topk_merge(scores, ids, nlists, nq, k_in, k_out, score_list_stride, id_list_stride, out_scores,
                                        out_ids, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(h, e, "topk_merge_kernel launch");
  h->launches += 1;
  return RS_OK;
}

// ------------------------------------------------------------------------------ MaxSim
int rs_maxsim(rs_handle* h, const void* q, int32_t nq, int32_t lq, int32_t d, int32_t dtype, const float* q_weight,
              const void* doc_tokens, int64_t n_tokens, const int32_t* doc_offsets, int32_t nd, const int32_t* cand,
              int32_t nc, float* out_scores, int32_t* out_argmax, float* out_tokmax, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (nq < 0 || nd < 0 || nc < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim: negative size");
  const int ndo = cand ? nc : nd;
  if (nq == 0 || ndo == 0) return RS_OK;  // empty documents -> [] (rerankers.py:281-282,364-365)
  if (!q || !doc_tokens || !doc_offsets || !out_scores) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim: NULL buffer");
  if (lq < 1 || d < 1) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim: lq and d must be positive");
  if (n_tokens < 1 || n_tokens >= (1ll << 31)) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim: n_tokens must be in [1, 2^31)");
  if (dtype != RS_F16 && dtype != RS_BF16 && dtype != RS_F32) return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim: bad dtype %d", dtype);
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  order_after_last(h, st);
  ProfScope prof(h, st);
  const double pair_tokens = (double)nq * ndo * ((double)n_tokens / (nd > 0 ? nd : 1));  // (query, document) token pairs
  const double ms_flops = 2.0 * pair_tokens * lq * d;
  const int64_t ms_bytes = (int64_t)((cand ? pair_tokens : (double)n_tokens) * d * (dtype == RS_F32 ? 4 : 2));
  rs::MaxSimParams p{q, q_weight, doc_tokens, doc_offsets, cand, out_scores, out_argmax, out_tokmax, n_tokens, nq, lq, d, nd, nc};

  if (dtype == RS_F32) {
    if (h->maxsim_impl != RS_MAXSIM_AUTO && h->maxsim_impl != RS_MAXSIM_SIMT)
      return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim: fp32 inputs run only on the SIMT path");
    if (rs::maxsim_simt_smem_bytes(lq, d) > 200 * 1024)
      return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim(fp32): lq * d too large for shared memory (lq=%d d=%d)", lq, d);
    cudaError_t e = rs::launch_maxsim_simt(p, st);
    if (e != cudaSuccess) return cuda_fail(h, e, "maxsim_simt_kernel launch");
    h->launches += 1;
    h->last_maxsim_impl = RS_MAXSIM_SIMT;
    prof.finish(RS_CALL_MAXSIM, RS_MAXSIM_SIMT, nq, ms_bytes, ms_flops);
    return RS_OK;
  }
  if (h->maxsim_impl == RS_MAXSIM_SIMT) return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim: SIMT path takes fp32 inputs only");
  if (!aligned16(q) || !aligned16(doc_tokens)) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim: q and doc_tokens must be 16-byte aligned");

  // Kernel family by shape: shared candidates with at least one full 128-row tile of query tokens and d in
  // {64, 128} -> the compute-bound tcgen05 kernel; per-query candidates, too few query tokens, d = 768 (any multiple
  // of 64 up to 1024) or the arg-max / per-token-max outputs of the explanations path -> the document-streaming
  // tcgen05 kernel; what is left (d % 64 != 0, query tiles too large for shared memory) -> the mma.sync kernel.
  int impl = h->maxsim_impl;
  const bool tc5_ok = rs::tc5_maxsim_supported(h->tc5, nq, lq, d, nd, cand, out_argmax) && out_tokmax == nullptr;
  const bool cand_ok = rs::tc5_maxsim_cand_supported(h->tc5, nq, lq, d, nd, ndo);
  if (impl == RS_MAXSIM_AUTO) impl = tc5_ok ? RS_MAXSIM_TCGEN05 : (cand_ok ? RS_MAXSIM_TCGEN05_CAND : RS_MAXSIM_MMA);
  if (impl == RS_MAXSIM_TCGEN05) {
    if (!tc5_ok)
      return fail(h, RS_ERR_UNSUPPORTED,
                  "rs_maxsim: shared-candidate tcgen05 path needs cand == NULL, no argmax, lq <= 128, d in {64,128}, "
                  "nq * lq_pad >= 128");
    int launched = 0;
    std::string err;
    int rc = rs::tc5_maxsim(h->tc5, p, dtype, st, &launched, &err);
    h->launches += launched;
    h->last_maxsim_impl = RS_MAXSIM_TCGEN05;
    if (rc != RS_OK) return fail(h, rc, "rs_maxsim(tcgen05): %s", err.c_str());
    prof.finish(RS_CALL_MAXSIM, RS_MAXSIM_TCGEN05, nq, ms_bytes, ms_flops);
    return RS_OK;
  }
  if (impl == RS_MAXSIM_TCGEN05_CAND) {
    if (!cand_ok)
      return fail(h, RS_ERR_UNSUPPORTED,
                  "rs_maxsim: candidate tcgen05 path needs lq <= 128, d a multiple of 64 up to 1024 and lq_pad * d <= 24576");
    int launched = 0;
    std::string err;
    int rc = rs::tc5_maxsim_cand(h->tc5, p, dtype, st, &launched, &err);
    h->launches += launched;
    h->last_maxsim_impl = RS_MAXSIM_TCGEN05_CAND;
    if (rc != RS_OK) return fail(h, rc, "rs_maxsim(tcgen05 candidates): %s", err.c_str());
    prof.finish(RS_CALL_MAXSIM, RS_MAXSIM_TCGEN05_CAND, nq, ms_bytes, ms_flops);
    return RS_OK;
  }
  if ((d % 16) != 0) return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim: d must be a multiple of 16 for fp16/bf16 (got %d)", d);
  if (lq > 128) return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim: lq must be <= 128 (got %d)", lq);
  if (rs::maxsim_mma_smem_bytes(lq, d) > 220 * 1024)
    return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim: lq=%d d=%d does not fit shared memory", lq, d);
  cudaError_t e = rs::launch_maxsim_mma(p, dtype, h->num_sms, st);
  if (e != cudaSuccess) return cuda_fail(h, e, "maxsim_mma_kernel launch");
  h->launches += 1;
  h->last_maxsim_impl = RS_MAXSIM_MMA;
  prof.finish(RS_CALL_MAXSIM, RS_MAXSIM_MMA, nq, ms_bytes, ms_flops);
  return RS_OK;
}

int rs_rerank_postprocess(rs_handle* h, const float* scores, const float* other, int32_t nq, int32_t n, float w_a,
                          float w_b, int32_t top_k, int32_t* out_idx, float* out_scores, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (nq == 0 || top_k == 0) return RS_OK;
  if (!scores || !out_idx || !out_scores) return fail(h, RS_ERR_INVALID_ARG, "rs_rerank_postprocess: NULL buffer");
  if (nq < 0 || n < 1 || top_k < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_rerank_postprocess: bad sizes");
  if (n > 16384) return fail(h, RS_ERR_UNSUPPORTED, "rs_rerank_postprocess: n must be <= 16384 (got %d)", n);
  DeviceGuard guard(h->device);
  cudaError_t e = rs::launch_rerank_postprocess(scores, other, nq, n, w_a, w_b, top_k, out_idx, out_scores,
                                                static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(h, e, "rerank_postprocess_kernel launch");
  h->launches += 1;
  return RS_OK;
}

int rs_owned_candidates(rs_handle* h, const void* cand, int32_t cand_is_i64, int64_t n, int32_t world, int32_t rank,
                        int64_t pool, int32_t* out_local, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (n < 0 || world < 1 || rank < 0 || rank >= world || pool < 0)
    return fail(h, RS_ERR_INVALID_ARG, "rs_owned_candidates: bad n / world / rank / pool");
  if (n == 0) return RS_OK;
  if (!cand || !out_local) return fail(h, RS_ERR_INVALID_ARG, "rs_owned_candidates: NULL argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  order_after_last(h, st);
  cudaError_t e = rs::launch_owned_candidates(cand, cand_is_i64, n, world, rank, pool, out_local, st);
  if (e != cudaSuccess) return cuda_fail(h, e, "owned_candidates_kernel launch");
  h->launches += 1;
  return RS_OK;
}

int rs_filter_mask(rs_handle* h, const int32_t* const* cols, int32_t nclauses, const int32_t* values,
                   const int32_t* val_offsets, const uint32_t* tombstone, int64_t n, uint32_t* out_mask, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (n < 0 || nclauses < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_filter_mask: negative size");
  if (n == 0) return RS_OK;
  if (!out_mask) return fail(h, RS_ERR_INVALID_ARG, "rs_filter_mask: NULL out_mask");
  if (nclauses > 16) return fail(h, RS_ERR_UNSUPPORTED, "rs_filter_mask: at most 16 clauses (got %d)", nclauses);
  if (nclauses > 0 && (!cols || !values || !val_offsets)) return fail(h, RS_ERR_INVALID_ARG, "rs_filter_mask: NULL clause arrays");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  order_after_last(h, st);
  const int nvals = nclauses > 0 ? val_offsets[nclauses] : 0;
  const size_t ints = (size_t)nvals + nclauses + 1;
  if (ints > h->filt_dev_ints) {
    if (h->filt_dev) cudaFree(h->filt_dev);
    h->filt_dev = nullptr;
    h->filt_dev_ints = 0;
    size_t want = ints < 4096 ? 4096 : ints;
    cudaError_t e = cudaMalloc(&h->filt_dev, want * sizeof(int32_t));
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc(filter scratch)");
    h->filt_dev_ints = want;
  }
  std::vector<int32_t> stage(ints);
  for (int i = 0; i <= nclauses; ++i) stage[i] = nclauses > 0 ? val_offsets[i] : 0;
  for (int i = 0; i < nvals; ++i) stage[nclauses + 1 + i] = values[i];
  // pageable -> device copy: the runtime stages it before returning, so `stage` may go away
  cudaError_t e = cudaMemcpyAsync(h->filt_dev, stage.data(), ints * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(h, e, "H2D(filter clauses)");
  e = rs::launch_filter_mask(cols, nclauses, h->filt_dev + nclauses + 1, h->filt_dev, tombstone, n, out_mask, h->num_sms, st);
  if (e != cudaSuccess) return cuda_fail(h, e, "filter_mask_kernel launch");
  h->launches += 1;
  return RS_OK;
}

// ------------------------------------------------------------------------------ multi-GPU exchange
int rs_comm_export(rs_handle* h, int32_t world, int32_t rank, int64_t slot_bytes, void* out_handle) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (!out_handle || slot_bytes < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_comm_export: bad argument");
  DeviceGuard guard(h->device);
  std::string err;
  const int rc = rs::comm_export(h->comm, world, rank, (size_t)slot_bytes, out_handle, &err);
  if (rc != 0) return fail(h, rc, "rs_comm_export: %s", err.c_str());
  return RS_OK;
}

int rs_comm_open(rs_handle* h, const void* handles) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (!handles) return fail(h, RS_ERR_INVALID_ARG, "rs_comm_open: NULL handles");
  DeviceGuard guard(h->device);
  std::string err;
  const int rc = rs::comm_open(h->comm, handles, &err);
  if (rc != 0) return fail(h, rc, "rs_comm_open: %s", err.c_str());
  return RS_OK;
}

int rs_comm_close(rs_handle* h) {
  if (!h) return RS_ERR_INVALID_ARG;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  rs::comm_close(h->comm);
  return RS_OK;
}

int rs_comm_info(const rs_handle* h, int32_t* world, int32_t* rank, int64_t* slot_bytes) {
  if (!h) return RS_ERR_INVALID_ARG;
  const bool open = rs::comm_is_open(h->comm);
  if (world) *world = open ? rs::comm_world(h->comm) : 0;
  if (rank) *rank = open ? rs::comm_rank(h->comm) : 0;
  if (slot_bytes) *slot_bytes = open ? (int64_t)rs::comm_slot_bytes(h->comm) : 0;
  return RS_OK;
}

int rs_allgather_topk(rs_handle* h, const float* local_scores, const int64_t* local_ids, int32_t nq, int32_t k_in,
                      int32_t k_out, float* out_scores, int64_t* out_ids, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (nq == 0) return RS_OK;
  if (!local_scores || !local_ids || !out_scores || !out_ids || nq < 0 || k_in < 1 || k_out < 1)
    return fail(h, RS_ERR_INVALID_ARG, "rs_allgather_topk: bad argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  order_after_last(h, st);
  std::string err;
  const int rc = rs::comm_allgather_topk(h->comm, local_scores, local_ids, nq, k_in, k_out, out_scores, out_ids, st, &err);
  if (rc != 0) return fail(h, rc, "rs_allgather_topk: %s", err.c_str());
  h->launches += 1;
  return RS_OK;
}

int rs_allgather(rs_handle* h, const void* local, int64_t bytes, void* out, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (bytes == 0) return RS_OK;
  if (!local || !out || bytes < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_allgather: bad argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  order_after_last(h, st);
  std::string err;
  const int rc = rs::comm_allgather(h->comm, local, (size_t)bytes, out, st, &err);
  if (rc != 0) return fail(h, rc, "rs_allgather: %s", err.c_str());
  h->launches += 1;
  return RS_OK;
}

int rs_allreduce_max_f32(rs_handle* h, const float* local, int64_t n, float* out, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (n == 0) return RS_OK;
  if (!local || !out || n < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_allreduce_max_f32: bad argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  order_after_last(h, st);
  std::string err;
  const int rc = rs::comm_allreduce_max(h->comm, local, (size_t)n, out, st, &err);
  if (rc != 0) return fail(h, rc, "rs_allreduce_max_f32: %s", err.c_str());
  h->launches += 1;
  return RS_OK;
}

// One request against a row-sharded corpus, host buffers in and out: H2D(queries) -> local top-k -> ONE fused
// push + merge kernel over peer memory writing the merged result into mapped pinned memory -> synchronise.
int rs_dense_topk_sharded_host(rs_handle* h, const void* corpus, int64_t n, int32_t d, int32_t dtype, const float* inv_norm,
                               int32_t metric, const void* queries_host, int32_t nq, const uint32_t* mask_dev,
                               int64_t mask_stride_words, int32_t k, int64_t id_base, float* out_scores_host,
                               int64_t* out_ids_host, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (nq < 0 || n < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk_sharded_host: negative size");
  if (nq == 0) return RS_OK;
  if (!queries_host || !out_scores_host || !out_ids_host) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk_sharded_host: NULL host buffer");
  if (d <= 0 || k < 1 || k > kMaxK) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk_sharded_host: bad d/k");
  if (!rs::comm_is_open(h->comm)) return fail(h, RS_ERR_INVALID_ARG, "rs_dense_topk_sharded_host: no exchange open (rs_comm_export + rs_comm_open)");
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t q_bytes = align_up((size_t)nq * d * 2, 256);
  const size_t os_bytes = align_up((size_t)nq * k * 4, 256);
  const size_t oi_bytes = align_up((size_t)nq * k * 8, 256);
  int rc = ensure_staging(h, q_bytes + os_bytes + oi_bytes, q_bytes);
  if (rc != RS_OK) return rc;
  const size_t pairs = (size_t)nq * k;
  if (pairs > h->shard_pairs) {
    if (h->shard_scores) cudaFree(h->shard_scores);
    if (h->shard_ids) cudaFree(h->shard_ids);
    h->shard_scores = nullptr;
    h->shard_ids = nullptr;
    h->shard_pairs = 0;
    const size_t want = pairs < 4096 ? 4096 : pairs;
    cudaError_t e = cudaMalloc(&h->shard_scores, want * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&h->shard_ids, want * sizeof(int64_t));
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc(shard top-k)");
    h->shard_pairs = want;
  }
  uint8_t* hp = static_cast<uint8_t*>(h->pinned);
  uint8_t* dp = static_cast<uint8_t*>(h->dev_stage);
  memcpy(hp, queries_host, (size_t)nq * d * 2);
  cudaError_t e = cudaMemcpyAsync(dp, hp, q_bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(h, e, "H2D(queries)");
  rc = rs_dense_topk(h, corpus, n, d, dtype, inv_norm, metric, dp, nq, mask_dev, mask_stride_words, k, id_base,
                     h->shard_scores, h->shard_ids, stream);
  if (rc != RS_OK) return rc;
  uint8_t* hp_dev = static_cast<uint8_t*>(h->pinned_dev);
  rc = rs_allgather_topk(h, h->shard_scores, h->shard_ids, nq, k, k, reinterpret_cast<float*>(hp_dev + q_bytes),
                         reinterpret_cast<int64_t*>(hp_dev + q_bytes + os_bytes), stream);
  if (rc != RS_OK) return rc;
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(h, e, "rs_dense_topk_sharded_host: stream synchronize");
  memcpy(out_scores_host, hp + q_bytes, (size_t)nq * k * 4);
  memcpy(out_ids_host, hp + q_bytes + os_bytes, (size_t)nq * k * 8);
  return RS_OK;
}

// ------------------------------------------------------------------------------ MaxSim over a document LIST
// The exact call shape of ColBERTReranker._compute_maxsim_scores (rerankers.py:215-265): one query, documents as
// separately allocated [Ld_i, D] matrices, a list of floats back.  Everything between the caller's pointers and the
// scores happens here: one staged upload (pointer table / offsets / weights, plus the raw tokens when they live in
// host memory — copied into pinned memory by a few threads), one gather-and-convert launch, one rs_maxsim launch
// writing into mapped pinned memory, one synchronise.
int rs_maxsim_list(rs_handle* h, const void* q, int32_t q_on_host, int32_t lq, int32_t d, int32_t src_dtype,
                   int32_t compute_dtype, const float* q_weight_host, const void* const* docs, const int32_t* doc_lens,
                   int32_t nd, int32_t docs_on_host, float* out_scores_host, void* stream) {
  if (!h) return RS_ERR_INVALID_ARG;
  if (nd < 0) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim_list: negative size");
  if (nd == 0) return RS_OK;
  if (!q || !docs || !doc_lens || !out_scores_host) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim_list: NULL argument");
  if (lq < 1 || d < 8 || (d % 8) != 0) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim_list: lq >= 1 and d a multiple of 8 required");
  for (int dt : {src_dtype, compute_dtype})
    if (dt != RS_F16 && dt != RS_BF16 && dt != RS_F32) return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim_list: bad dtype %d", dt);
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t es = src_dtype == RS_F32 ? 4 : 2, ec = compute_dtype == RS_F32 ? 4 : 2;
  int64_t total_rows = 0;
  int max_len = lq;
  for (int i = 0; i < nd; ++i) {
    if (doc_lens[i] < 0 || !docs[i]) return fail(h, RS_ERR_INVALID_ARG, "rs_maxsim_list: document %d is NULL or has a negative length", i);
    total_rows += doc_lens[i];
    if (doc_lens[i] > max_len) max_len = doc_lens[i];
  }
  if (total_rows + lq >= (1ll << 31)) return fail(h, RS_ERR_UNSUPPORTED, "rs_maxsim_list: too many tokens");
  // staged layout (same in pinned host memory and in the device mirror); the query travels as document nd
  const int n_src = nd + 1;
  const size_t o_ptr = 0, o_soff = align_up(o_ptr + (size_t)n_src * 8, 16), o_off = align_up(o_soff + (size_t)n_src * 8, 16);
  const size_t o_w = align_up(o_off + (size_t)(n_src + 1) * 4, 16), o_raw = align_up(o_w + (size_t)lq * 4, 256);
  size_t raw_bytes = 0;
  std::vector<int64_t> soff(n_src, 0);
  for (int i = 0; i < n_src; ++i) {
    const bool on_host = i < nd ? docs_on_host != 0 : q_on_host != 0;
    const size_t bytes = (size_t)(i < nd ? doc_lens[i] : lq) * d * es;
    if (on_host) {
      soff[i] = (int64_t)(o_raw + raw_bytes);
      raw_bytes += align_up(bytes, 16);
    }
  }
  const size_t in_bytes = o_raw + raw_bytes, o_out = align_up(in_bytes, 256), host_bytes = o_out + align_up((size_t)nd * 4, 256);
  int rc = ensure_staging(h, host_bytes, in_bytes);
  if (rc != RS_OK) return rc;
  const size_t packed_bytes = (size_t)(total_rows + lq) * d * ec;
  if (packed_bytes > h->list_buf_bytes) {
    if (h->list_buf) cudaFree(h->list_buf);
    h->list_buf = nullptr;
    h->list_buf_bytes = 0;
    const size_t want = packed_bytes < (4u << 20) ? (4u << 20) : packed_bytes + packed_bytes / 4;
    cudaError_t e = cudaMalloc(&h->list_buf, want);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc(packed documents)");
    h->list_buf_bytes = want;
  }
  uint8_t* hp = static_cast<uint8_t*>(h->pinned);
  uint8_t* dp = static_cast<uint8_t*>(h->dev_stage);
  const void** ptrs = reinterpret_cast<const void**>(hp + o_ptr);
  int32_t* offs = reinterpret_cast<int32_t*>(hp + o_off);
  offs[0] = 0;
  for (int i = 0; i < n_src; ++i) {
    ptrs[i] = i < nd ? docs[i] : q;
    offs[i + 1] = offs[i] + (i < nd ? doc_lens[i] : lq);
  }
  memcpy(hp + o_soff, soff.data(), (size_t)n_src * 8);
  if (q_weight_host) memcpy(hp + o_w, q_weight_host, (size_t)lq * 4);
  // Upload: the tables (and the query) first, then the documents.  Pageable -> pinned runs at ~10 GB/s per thread,
  // less than the DMA engine moves, so a large list is split over worker threads, and every thread hands each
  // ~512 KB it has staged to the copy engine at once: the H2D transfer of one piece overlaps the staging of the
  // next instead of starting when the whole list has been copied (round 2: 0.94 ms per call for 9.2 MB).
  if (q_on_host) memcpy(hp + soff[nd], q, (size_t)lq * d * es);
  order_after_last(h, st);
  cudaError_t e = cudaMemcpyAsync(dp, hp, o_raw, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && q_on_host)
    e = cudaMemcpyAsync(dp + soff[nd], hp + soff[nd], (size_t)lq * d * es, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(h, e, "H2D(document list tables)");
  if (docs_on_host && nd > 0) {
    std::atomic<int> copy_err{0};
    const int device = h->device;
    auto copy_range = [&](int a, int b, bool set_device) {
      if (set_device) cudaSetDevice(device);
      constexpr size_t kPiece = 512u << 10;
      int first = a;
      for (int i = a; i < b; ++i) {
        memcpy(hp + soff[i], docs[i], (size_t)doc_lens[i] * d * es);
        const size_t end = (size_t)soff[i] + align_up((size_t)doc_lens[i] * d * es, 16);
        if (end - (size_t)soff[first] >= kPiece || i + 1 == b) {
          const cudaError_t ce = cudaMemcpyAsync(dp + soff[first], hp + soff[first], end - (size_t)soff[first],
                                                 cudaMemcpyHostToDevice, st);
          if (ce != cudaSuccess) copy_err.store((int)ce);
          first = i + 1;
        }
      }
    };
    // parts = staging threads incl. the caller.  Threads created per call (round 2, first half): 1 / 3 / 6 threads ->
    // 0.85 / 0.72 / 1.08 ms per call for 9.2 MB of fp32 documents; the handle's persistent workers: 1 / 3 / 6 / 8 / 12
    // parts -> 0.84 / 0.47 / 0.49 / 0.50 / 0.54 ms (profiles/r02_config1_staging_pool.txt).
    static const int forced_threads = getenv("RS_LIST_THREADS") ? atoi(getenv("RS_LIST_THREADS")) : 0;
    const int want = forced_threads > 0 ? forced_threads : (raw_bytes > (1u << 20) ? 3 : 1);
    const int parts = std::max(1, std::min(want, nd));
    if (parts == 1) {
      copy_range(0, nd, false);
    } else {
      if (h->stage_pool && h->stage_pool->workers() != parts - 1) {
        delete h->stage_pool;
        h->stage_pool = nullptr;
      }
      if (!h->stage_pool) h->stage_pool = new (std::nothrow) StagePool(parts - 1, device);
      if (!h->stage_pool) {
        copy_range(0, nd, false);
      } else {
        const std::function<void(int)> job = [&](int part) {
          copy_range((int)((int64_t)nd * part / parts), (int)((int64_t)nd * (part + 1) / parts), false);
        };
        h->stage_pool->run(job);
      }
    }
    if (copy_err.load() != 0) return cuda_fail(h, (cudaError_t)copy_err.load(), "H2D(document list)");
  }
  // documents and the query are never mixed host / device in one table: two launches when they differ
  const bool same_side = (docs_on_host != 0) == (q_on_host != 0);
  auto gather = [&](int first, int count, bool on_host) {
    return rs::launch_gather_docs(on_host ? nullptr : reinterpret_cast<const void* const*>(dp + o_ptr) + first, on_host ? dp : nullptr,
                                  reinterpret_cast<const int64_t*>(dp + o_soff) + first, reinterpret_cast<const int32_t*>(dp + o_off) + first,
                                  count, max_len, d, src_dtype, compute_dtype, h->list_buf, st);
  };
  if (same_side) {
    e = gather(0, n_src, docs_on_host != 0);
    h->launches += 1;
  } else {
    e = gather(0, nd, docs_on_host != 0);
    if (e == cudaSuccess) e = gather(nd, 1, q_on_host != 0);
    h->launches += 2;
  }
  if (e != cudaSuccess) return cuda_fail(h, e, "gather_docs_kernel launch");
  const uint8_t* packed = static_cast<const uint8_t*>(h->list_buf);
  float* out_dev = reinterpret_cast<float*>(static_cast<uint8_t*>(h->pinned_dev) + o_out);
  rc = rs_maxsim(h, packed + (size_t)total_rows * d * ec, 1, lq, d, compute_dtype,
                 q_weight_host ? reinterpret_cast<const float*>(dp + o_w) : nullptr, packed, total_rows > 0 ? total_rows : 1,
                 reinterpret_cast<const int32_t*>(dp + o_off), nd, nullptr, 0, out_dev, nullptr, nullptr, stream);
  if (rc != RS_OK) return rc;
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(h, e, "rs_maxsim_list: stream synchronize");
  memcpy(out_scores_host, hp + o_out, (size_t)nd * 4);
  return RS_OK;
}

}  // extern "C"
