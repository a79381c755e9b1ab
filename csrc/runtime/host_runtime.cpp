// Host runtime primitives behind AsyncContext (C ABI, loaded through ctypes; ctypes drops the
// GIL around every call so blocking take() never stalls other Python threads).
//
//  * ResultQueue  - blocking MPMC queue of fixed-size result records; the analogue of
//                   ASYNCcontext.ResultList: LinkedBlockingQueue (core/rdd/ASYNCcontext.scala:15).
//                   Payload tensors stay in Python (slot table); the queue carries metadata.
//  * LogicalClock - the CurrentTime/LastTime counters (ASYNCcontext.scala:24-54) as atomics
//                   (fixes the non-volatile race noted in SURVEY.md §5.2).
//  * WorkerTable  - ASYNCcontext.STAT / workerState (core/rdd/workerState.scala:14-87) as a
//                   fixed array of atomics: staleness, average task time, availability, task
//                   count, heartbeat; O(P) global queries available_workers / max_staleness.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Record {
  int64_t payload_id;
  int64_t ts;          // logical clock at submit
  int64_t staleness;   // clock at merge - ts
  int64_t batch_size;  // true number of records reduced (the reference always reports Int.MinValue)
  int32_t worker;
  int32_t flags;
  double t_submit_ms;
  double t_finish_ms;
};

struct ResultQueue {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Record> q;
  bool closed = false;
};

struct WorkerSlot {
  std::atomic<int64_t> staleness{0};
  std::atomic<double> avg_task_time_ms{0.0};
  std::atomic<int32_t> available{1};
  std::atomic<int64_t> num_tasks{0};
  std::atomic<int32_t> has_entry{0};
  std::atomic<int32_t> alive{1};
  std::atomic<double> heartbeat_ms{0.0};
};

struct WorkerTable {
  std::vector<WorkerSlot> slots;
  explicit WorkerTable(int p) : slots(static_cast<size_t>(p)) {}
};

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

// ---- ResultQueue -------------------------------------------------------------------------
void* afb_queue_create() { return new ResultQueue(); }
void afb_queue_destroy(void* h) { delete static_cast<ResultQueue*>(h); }

void afb_queue_put(void* h, const Record* r) {
  auto* q = static_cast<ResultQueue*>(h);
  {
    std::lock_guard<std::mutex> g(q->mu);
    q->q.push_back(*r);
  }
  q->cv.notify_one();
}

// returns 1 on success, 0 on timeout, -1 if closed and empty.  timeout_ms < 0 => wait forever.
int afb_queue_take(void* h, Record* out, double timeout_ms) {
  auto* q = static_cast<ResultQueue*>(h);
  std::unique_lock<std::mutex> lk(q->mu);
  auto ready = [&] { return !q->q.empty() || q->closed; };
  if (timeout_ms < 0) {
    q->cv.wait(lk, ready);
  } else if (!q->cv.wait_for(lk, std::chrono::duration<double, std::milli>(timeout_ms), ready)) {
    return 0;
  }
  if (q->q.empty()) return -1;
  *out = q->q.front();
  q->q.pop_front();
  return 1;
}

// Batch take: blocks like afb_queue_take for the FIRST record, then drains up to max_n records that are already
// queued under the same lock acquisition.  Returns the number of records written, 0 on timeout, -1 if closed and empty.
int afb_queue_take_batch(void* h, Record* out, int max_n, double timeout_ms) {
  auto* q = static_cast<ResultQueue*>(h);
  if (max_n <= 0) return 0;
  std::unique_lock<std::mutex> lk(q->mu);
  auto ready = [&] { return !q->q.empty() || q->closed; };
  if (timeout_ms < 0) {
    q->cv.wait(lk, ready);
  } else if (!q->cv.wait_for(lk, std::chrono::duration<double, std::milli>(timeout_ms), ready)) {
    return 0;
  }
  if (q->q.empty()) return -1;
  int n = 0;
  while (n < max_n && !q->q.empty()) {
    out[n++] = q->q.front();
    q->q.pop_front();
  }
  return n;
}

int64_t afb_queue_size(void* h) {
  auto* q = static_cast<ResultQueue*>(h);
  std::lock_guard<std::mutex> g(q->mu);
  return static_cast<int64_t>(q->q.size());
}

void afb_queue_close(void* h) {
  auto* q = static_cast<ResultQueue*>(h);
  {
    std::lock_guard<std::mutex> g(q->mu);
    q->closed = true;
  }
  q->cv.notify_all();
}

void afb_queue_reopen(void* h) {
  auto* q = static_cast<ResultQueue*>(h);
  std::lock_guard<std::mutex> g(q->mu);
  q->closed = false;
}

// ---- LogicalClock ------------------------------------------------------------------------
void* afb_clock_create() { return new std::atomic<int64_t>(0); }
void afb_clock_destroy(void* h) { delete static_cast<std::atomic<int64_t>*>(h); }
int64_t afb_clock_get(void* h) { return static_cast<std::atomic<int64_t>*>(h)->load(std::memory_order_acquire); }
int64_t afb_clock_add(void* h, int64_t d) {  // returns the value BEFORE the add
  return static_cast<std::atomic<int64_t>*>(h)->fetch_add(d, std::memory_order_acq_rel);
}
void afb_clock_set(void* h, int64_t v) { static_cast<std::atomic<int64_t>*>(h)->store(v, std::memory_order_release); }

// ---- WorkerTable -------------------------------------------------------------------------
void* afb_table_create(int p) { return new WorkerTable(p); }
void afb_table_destroy(void* h) { delete static_cast<WorkerTable*>(h); }
int afb_table_size(void* h) { return static_cast<int>(static_cast<WorkerTable*>(h)->slots.size()); }

void afb_table_set_state(void* h, int p, int64_t staleness, double avg_ms, int available, int64_t num_tasks) {
  auto& s = static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)];
  s.staleness.store(staleness, std::memory_order_relaxed);
  s.avg_task_time_ms.store(avg_ms, std::memory_order_relaxed);
  s.num_tasks.store(num_tasks, std::memory_order_relaxed);
  s.available.store(available, std::memory_order_relaxed);
  s.heartbeat_ms.store(now_ms(), std::memory_order_relaxed);
  s.has_entry.store(1, std::memory_order_release);
}

// Folds one finished task into the running mean (fixes reference quirk B.19 where numTasks is
// always 1).  Returns the new task count.
int64_t afb_table_task_done(void* h, int p, int64_t staleness, double task_ms) {
  auto& s = static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)];
  int64_t n = s.num_tasks.fetch_add(1, std::memory_order_acq_rel) + 1;
  double prev = s.avg_task_time_ms.load(std::memory_order_relaxed);
  s.avg_task_time_ms.store(prev + (task_ms - prev) / static_cast<double>(n), std::memory_order_relaxed);
  s.staleness.store(staleness, std::memory_order_relaxed);
  s.heartbeat_ms.store(now_ms(), std::memory_order_relaxed);
  s.has_entry.store(1, std::memory_order_release);
  s.available.store(1, std::memory_order_release);
  return n;
}

void afb_table_set_available(void* h, int p, int available) {
  auto& s = static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)];
  s.available.store(available, std::memory_order_release);
  s.has_entry.store(1, std::memory_order_release);
}
void afb_table_set_alive(void* h, int p, int alive) {
  static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].alive.store(alive, std::memory_order_release);
}
void afb_table_heartbeat(void* h, int p) {
  static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].heartbeat_ms.store(now_ms(), std::memory_order_relaxed);
}

int afb_table_has_entry(void* h, int p) { return static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].has_entry.load(std::memory_order_acquire); }
int afb_table_available(void* h, int p) { return static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].available.load(std::memory_order_acquire); }
int afb_table_alive(void* h, int p) { return static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].alive.load(std::memory_order_acquire); }
int64_t afb_table_staleness(void* h, int p) { return static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].staleness.load(std::memory_order_relaxed); }
double afb_table_avg_ms(void* h, int p) { return static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].avg_task_time_ms.load(std::memory_order_relaxed); }
int64_t afb_table_num_tasks(void* h, int p) { return static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].num_tasks.load(std::memory_order_relaxed); }
double afb_table_heartbeat_age_ms(void* h, int p) {
  return now_ms() - static_cast<WorkerTable*>(h)->slots[static_cast<size_t>(p)].heartbeat_ms.load(std::memory_order_relaxed);
}

// workerState.getAvailableWorkers (workerState.scala:63-72): counts entries whose availability is true.
int afb_table_available_workers(void* h) {
  auto* t = static_cast<WorkerTable*>(h);
  int n = 0;
  for (auto& s : t->slots)
    if (s.has_entry.load(std::memory_order_acquire) && s.alive.load(std::memory_order_relaxed) &&
        s.available.load(std::memory_order_acquire)) ++n;
  return n;
}

// workerState.getMaxStaleness (workerState.scala:74-82)
int64_t afb_table_max_staleness(void* h) {
  auto* t = static_cast<WorkerTable*>(h);
  int64_t m = 0;
  for (auto& s : t->slots)
    if (s.has_entry.load(std::memory_order_acquire)) m = std::max<int64_t>(m, s.staleness.load(std::memory_order_relaxed));
  return m;
}

double afb_now_ms() { return now_ms(); }

// Native self-benchmark of the result queue: `producers` threads put `per_producer` records each while the calling
// thread takes them (single records, or batches of `batch` when batch > 1).  Returns the elapsed seconds.
double afb_queue_bench(int producers, int64_t per_producer, int batch) {
  auto* q = static_cast<ResultQueue*>(afb_queue_create());
  std::vector<std::thread> th;
  const double t0 = now_ms();
  for (int p = 0; p < producers; ++p)
    th.emplace_back([q, p, per_producer] {
      Record r{};
      r.worker = p;
      for (int64_t i = 0; i < per_producer; ++i) { r.payload_id = i; afb_queue_put(q, &r); }
    });
  const int64_t total = static_cast<int64_t>(producers) * per_producer;
  int64_t got = 0;
  std::vector<Record> buf(static_cast<size_t>(batch > 1 ? batch : 1));
  while (got < total) {
    int n = (batch > 1) ? afb_queue_take_batch(q, buf.data(), batch, -1.0) : afb_queue_take(q, buf.data(), -1.0);
    if (n > 0) got += n;
  }
  const double dt = (now_ms() - t0) * 1e-3;
  for (auto& t : th) t.join();
  afb_queue_destroy(q);
  return dt;
}

}  // extern "C"
