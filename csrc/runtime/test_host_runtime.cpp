// Multi-threaded stress test of the host runtime primitives (ResultQueue / LogicalClock / WorkerTable).
// Built and run by tests/test_native_runtime.py, once plain and once with -fsanitize=thread
// ("TSAN builds of the C++ host runtime").  Exit code 0 = all invariants hold.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

struct Record {
  int64_t payload_id, ts, staleness, batch_size;
  int32_t worker, flags;
  double t_submit_ms, t_finish_ms;
};

extern "C" {
void* afb_queue_create();
void afb_queue_destroy(void*);
void afb_queue_put(void*, const Record*);
int afb_queue_take(void*, Record*, double);
int64_t afb_queue_size(void*);
void afb_queue_close(void*);
void* afb_clock_create();
void afb_clock_destroy(void*);
int64_t afb_clock_get(void*);
int64_t afb_clock_add(void*, int64_t);
void* afb_table_create(int);
void afb_table_destroy(void*);
int64_t afb_table_task_done(void*, int, int64_t, double);
void afb_table_set_available(void*, int, int);
int afb_table_available_workers(void*);
int64_t afb_table_max_staleness(void*);
int64_t afb_table_num_tasks(void*, int);
double afb_table_avg_ms(void*, int);
}

#define CHECK(c)                                                       \
  do {                                                                 \
    if (!(c)) { fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); return 1; } \
  } while (0)

int main() {
  const int P = 8, PER = 5000;
  void* q = afb_queue_create();
  void* clk = afb_clock_create();
  void* tab = afb_table_create(P);
  std::atomic<int64_t> sum_in{0}, sum_out{0}, taken{0};

  // P producers ("task completions") + 2 consumers ("updater threads") + 1 observer (barrier predicate reader)
  std::vector<std::thread> th;
  for (int p = 0; p < P; ++p)
    th.emplace_back([&, p] {
      for (int i = 0; i < PER; ++i) {
        afb_table_set_available(tab, p, 0);
        int64_t ts = afb_clock_get(clk);
        int64_t before = afb_clock_add(clk, 1);
        Record r{static_cast<int64_t>(p) * PER + i + 1, ts, before - ts, 10, p, 0, 0.0, 1.0};
        sum_in += r.payload_id;
        afb_table_task_done(tab, p, r.staleness, 2.0);
        afb_queue_put(q, &r);
      }
    });
  std::atomic<bool> stop{false};
  for (int c = 0; c < 2; ++c)
    th.emplace_back([&] {
      Record r;
      while (true) {
        int rc = afb_queue_take(q, &r, 50.0);
        if (rc == 1) { sum_out += r.payload_id; ++taken; if (r.staleness < 0) abort(); }
        else if (rc == -1) break;
        else if (stop.load() && afb_queue_size(q) == 0) break;
      }
    });
  std::thread obs([&] {
    while (!stop.load()) {
      int a = afb_table_available_workers(tab);
      int64_t m = afb_table_max_staleness(tab);
      if (a < 0 || a > P || m < 0) abort();
    }
  });
  for (int p = 0; p < P; ++p) th[p].join();
  stop = true;
  afb_queue_close(q);
  for (size_t i = P; i < th.size(); ++i) th[i].join();
  obs.join();

  CHECK(taken.load() == static_cast<int64_t>(P) * PER);          // no lost / duplicated results
  CHECK(sum_in.load() == sum_out.load());
  CHECK(afb_clock_get(clk) == static_cast<int64_t>(P) * PER);     // the logical clock counts every merge exactly once
  for (int p = 0; p < P; ++p) {
    CHECK(afb_table_num_tasks(tab, p) == PER);                    // running task count accumulates (quirk B.19 fixed)
    CHECK(afb_table_avg_ms(tab, p) > 1.99 && afb_table_avg_ms(tab, p) < 2.01);
  }
  CHECK(afb_table_available_workers(tab) == P);
  afb_queue_destroy(q);
  afb_clock_destroy(clk);
  afb_table_destroy(tab);
  printf("host runtime stress: OK (%lld results)\n", static_cast<long long>(taken.load()));
  return 0;
}
