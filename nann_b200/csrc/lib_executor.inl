// lib_executor.inl -- closed-loop load generator + request executor: what blaze-benchmark does around
// Session::Run (blaze-benchmark/benchmark/core/benchmark.cc:101-146, model.cc:19-53,192-234,
// predict_request_producer.cc:49-80, predict_request_consumer.cc:17-54, metrics.cc:5-94), re-done for
// one CUDA context: `predictor_num` searchers (the reference's sessions on virtual GPUs) each own a
// CUDA stream and a consumer thread; producers enqueue requests; a consumer takes up to
// max_batch_size queued requests and runs them as ONE nann_search_batch call (max_batch_size=1 is the
// reference's behaviour: every request is its own run).  Meters/histograms carry the reference's
// names and quantiles (cppmetrics ConsoleReporter: count, mean rate; min max mean stddev median 75%
// 95% 98% 99% 99.9%).

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace nann {

struct ExecRequest {
  int64_t query;
  std::chrono::steady_clock::time_point enq;
};

struct ExecShared {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<ExecRequest> queue;
  bool stop = false;
  // metrics
  std::mutex mmu;
  std::vector<float> lat_us, e2e_us, batch;
  int64_t ok = 0, failures = 0, dropped = 0;
};

static void hist_stats(std::vector<float>& v, double out[11]) {
  for (int i = 0; i < 11; ++i) out[i] = 0;
  if (v.empty()) return;
  std::sort(v.begin(), v.end());
  const size_t n = v.size();
  double sum = 0, sq = 0;
  for (float x : v) { sum += x; sq += (double)x * x; }
  const double mean = sum / n;
  auto q = [&](double p) {  // cppmetrics Snapshot::getValue: linear interpolation at p*(n+1)
    double pos = p * (n + 1);
    if (pos < 1) return (double)v[0];
    if (pos >= n) return (double)v[n - 1];
    const double lo = v[(size_t)pos - 1], hi = v[(size_t)pos];
    return lo + (pos - std::floor(pos)) * (hi - lo);
  };
  out[0] = (double)n; out[1] = v[0]; out[2] = v[n - 1]; out[3] = mean;
  out[4] = n > 1 ? std::sqrt(std::max(0.0, (sq - n * mean * mean) / (n - 1))) : 0.0;
  out[5] = q(0.5); out[6] = q(0.75); out[7] = q(0.95); out[8] = q(0.98); out[9] = q(0.99); out[10] = q(0.999);
}

static void print_hist(FILE* f, const char* name, const double h[11]) {
  fprintf(f, "%s:\n             count = %lld\n               min = %.0f\n               max = %.0f\n              mean = %.2f\n"
             "            stddev = %.2f\n            median = %.2f\n              75%% <= %.2f\n              95%% <= %.2f\n"
             "              98%% <= %.2f\n              99%% <= %.2f\n            99.9%% <= %.2f\n",
          name, (long long)h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10]);
}

}  // namespace nann

extern "C" {

nann_status nann_executor_run(const nann_index_t* ix, nann_scorer_t* scorer, const nann_bench_conf_t* conf,
                              const int32_t level_topn[6], const float* queries, int64_t n_queries,
                              int print_reports, nann_bench_report_t* report) {
  NANN_TRY(require_device());
  if (!ix || !scorer || !conf || !level_topn || !queries || n_queries <= 0 || !report)
    return fail(NANN_INVALID_ARGUMENT, "nann_executor_run: null or empty argument");
  if (is_device_ptr(queries)) return fail(NANN_INVALID_ARGUMENT, "queries must be host memory (requests arrive on the host)");
  const int P = std::max(1, conf->predictor_num), T = std::max(1, conf->bench_thread_count);
  const int MB = std::max(1, conf->max_batch_size);
  if (MB > 65535) return fail(NANN_UNIMPLEMENTED, "max_batch_size > 65535");
  const int uf = nann_scorer_user_floats(scorer);
  const int k = level_topn[5];
  NANN_CUDA(cudaSetDevice(ix->device));

  struct Predictor {
    nann_searcher_t* se = nullptr; cudaStream_t st = nullptr; float* users = nullptr; int64_t* ids = nullptr;
    float* sc = nullptr; int32_t* status = nullptr;
  };
  std::vector<Predictor> preds(P);
  nann_status rc = NANN_OK;
  for (int p = 0; p < P && rc == NANN_OK; ++p) {
    rc = nann_searcher_create(ix, scorer, MB, level_topn, &preds[p].se);
    if (rc != NANN_OK) break;
    if (cudaStreamCreateWithFlags(&preds[p].st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&preds[p].users, (size_t)MB * uf * 4) != cudaSuccess ||
        cudaMallocHost(&preds[p].ids, (size_t)MB * std::max(k, 1) * 8) != cudaSuccess ||
        cudaMallocHost(&preds[p].sc, (size_t)MB * std::max(k, 1) * 4) != cudaSuccess ||
        cudaMallocHost(&preds[p].status, (size_t)MB * 4) != cudaSuccess)
      rc = fail(NANN_RESOURCE_EXHAUSTED, "executor: pinned buffers / stream allocation failed");
  }
  auto cleanup = [&]() {
    for (auto& p : preds) {
      nann_searcher_destroy(p.se);
      if (p.st) cudaStreamDestroy(p.st);
      cudaFreeHost(p.users); cudaFreeHost(p.ids); cudaFreeHost(p.sc); cudaFreeHost(p.status);
    }
  };
  if (rc != NANN_OK) { cleanup(); return rc; }

  // Warm-up (Model::Warmup, model.cc:364-382): one call per predictor
  for (int p = 0; p < P; ++p) {
    const int wb = (int)std::min<int64_t>(MB, n_queries);
    memcpy(preds[p].users, queries, (size_t)wb * uf * 4);
    rc = nann_search_batch(preds[p].se, preds[p].users, wb, level_topn, preds[p].ids, preds[p].sc, preds[p].status,
                           nullptr, preds[p].st);
    if (rc != NANN_OK) { cleanup(); return rc; }
  }

  ExecShared sh;
  const auto t_start = std::chrono::steady_clock::now();
  const auto t_end = t_start + std::chrono::milliseconds((int64_t)(1000.0 * std::max(conf->duration_s, 0.001)));
  const size_t keep_full = (size_t)std::max(2 * MB * P, 64);
  std::atomic<int64_t> next_q{0};
  std::atomic<int> first_error{NANN_OK};

  auto producer = [&](int t) {
    // qps <= 0: keep the queue topped up (the reference enqueues one request per 100 us per thread,
    // predict_request_producer.cc:72-78, which caps the offered load; a closed loop needs no such cap).
    // qps > 0: pace this thread at qps / bench_thread_count.
    const double per_thread = conf->qps > 0 ? conf->qps / T : 0.0;
    auto next_time = std::chrono::steady_clock::now();
    (void)t;
    while (std::chrono::steady_clock::now() < t_end) {
      if (per_thread > 0) {
        std::this_thread::sleep_until(next_time);
        next_time += std::chrono::nanoseconds((int64_t)(1e9 / per_thread));
        std::lock_guard<std::mutex> l(sh.mu);
        sh.queue.push_back({next_q.fetch_add(1) % n_queries, std::chrono::steady_clock::now()});
        sh.cv.notify_one();
      } else {
        std::unique_lock<std::mutex> l(sh.mu);
        if (sh.queue.size() >= keep_full) {
          l.unlock();
          std::this_thread::sleep_for(std::chrono::microseconds(20));
          continue;
        }
        const auto now = std::chrono::steady_clock::now();
        for (int i = 0; i < MB && sh.queue.size() < keep_full; ++i)
          sh.queue.push_back({next_q.fetch_add(1) % n_queries, now});
        sh.cv.notify_all();
      }
    }
  };

  auto consumer = [&](int p) {
    Predictor& pr = preds[p];
    cudaSetDevice(ix->device);
    std::vector<ExecRequest> batch;
    while (true) {
      batch.clear();
      {
        std::unique_lock<std::mutex> l(sh.mu);
        sh.cv.wait_for(l, std::chrono::microseconds(1000), [&] { return sh.stop || !sh.queue.empty(); });
        if (sh.stop) return;
        if (sh.queue.empty()) continue;
        if (conf->max_queue_size > 0 && (int)sh.queue.size() > conf->max_queue_size) {  // predict_request_consumer.cc:31-35
          sh.queue.pop_front();
          std::lock_guard<std::mutex> ml(sh.mmu);
          sh.dropped++;
          continue;
        }
        if ((int)sh.queue.size() < MB && conf->batch_timeout_us > 0) {
          sh.cv.wait_for(l, std::chrono::microseconds(conf->batch_timeout_us), [&] { return sh.stop || (int)sh.queue.size() >= MB; });
          if (sh.stop) return;
        }
        while (!sh.queue.empty() && (int)batch.size() < MB) { batch.push_back(sh.queue.front()); sh.queue.pop_front(); }
      }
      const int B = (int)batch.size();
      for (int i = 0; i < B; ++i) memcpy(pr.users + (size_t)i * uf, queries + batch[i].query * uf, (size_t)uf * 4);
      const auto bef = std::chrono::steady_clock::now();
      nann_status s = nann_search_batch(pr.se, pr.users, B, level_topn, pr.ids, pr.sc, pr.status, nullptr, pr.st);
      const auto aft = std::chrono::steady_clock::now();
      const float dur = (float)std::chrono::duration_cast<std::chrono::nanoseconds>(aft - bef).count() / 1000.f;
      std::lock_guard<std::mutex> ml(sh.mmu);
      if (s != NANN_OK) {
        int expected = NANN_OK;
        first_error.compare_exchange_strong(expected, s);
        sh.failures += B;
        continue;
      }
      sh.batch.push_back((float)B);
      for (int i = 0; i < B; ++i) {
        if (pr.status[i] == NANN_OK) sh.ok++; else sh.failures++;
        sh.lat_us.push_back(dur);   // latency of the run the request was part of (consumer.cc:41-50)
        sh.e2e_us.push_back((float)std::chrono::duration_cast<std::chrono::nanoseconds>(aft - batch[i].enq).count() / 1000.f);
      }
    }
  };

  std::vector<std::thread> threads;
  for (int p = 0; p < P; ++p) threads.emplace_back(consumer, p);
  for (int t = 0; t < T; ++t) threads.emplace_back(producer, t);

  // ConsoleReporter: every report_interval_s (3 s in the reference, metrics.cc:15)
  const int interval = conf->report_interval_s > 0 ? conf->report_interval_s : 3;
  auto next_report = t_start + std::chrono::seconds(interval);
  while (std::chrono::steady_clock::now() < t_end) {
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    if (print_reports && std::chrono::steady_clock::now() >= next_report) {
      next_report += std::chrono::seconds(interval);
      std::lock_guard<std::mutex> ml(sh.mmu);
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
      fprintf(stderr, "-- Meters --\nnann_throughput:\n             count = %lld\n         mean rate = %.2f events/second\n",
              (long long)sh.ok, sh.ok / el);
    }
  }
  {
    std::lock_guard<std::mutex> l(sh.mu);
    sh.stop = true;
    sh.cv.notify_all();
  }
  for (auto& th : threads) th.join();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();

  memset(report, 0, sizeof(*report));
  report->seconds = secs;
  report->throughput_count = sh.ok;
  report->mean_rate = sh.ok / secs;
  report->failures = sh.failures;
  report->get_predictor_failures = sh.dropped;
  hist_stats(sh.lat_us, report->latency_us);
  hist_stats(sh.e2e_us, report->e2e_latency_us);
  hist_stats(sh.batch, report->batchsize);
  if (print_reports) {
    fprintf(stderr, "-- Meters --\nnann_throughput:\n             count = %lld\n         mean rate = %.2f events/second\n"
                    "nann_failures:\n             count = %lld\nnann_get_predictor_failures:\n             count = %lld\n-- Histograms --\n",
            (long long)sh.ok, report->mean_rate, (long long)sh.failures, (long long)sh.dropped);
    print_hist(stderr, "nann_latency", report->latency_us);
    print_hist(stderr, "nann_batchsize", report->batchsize);
    print_hist(stderr, "nann_e2e_latency", report->e2e_latency_us);
  }
  cleanup();
  if (first_error.load() != NANN_OK && sh.ok == 0) return fail((nann_status)first_error.load(), "every search call failed");
  return NANN_OK;
}

}  // extern "C"
