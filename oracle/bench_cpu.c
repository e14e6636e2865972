/* bench_cpu.c — C timing harness for the reference's CPU path (TEST / BENCH INFRASTRUCTURE, like the rest of oracle/).
 *
 * time the reference loops from C (ctypes adds microseconds per call), pinned to one core
 * (sched_setaffinity), warm-up then clock_gettime per step.  The loop under test is passed in as a function pointer
 * (oracle/_ref's arithmetic_binary_avx2 — the reference's own instruction stream — or the restatement), so this file
 * contains no arithmetic of its own.
 *
 * bench_add_f64_chunked: compute.Add(float64, float64) over a chunked column the way the reference's executor runs
 * it — iterateExecSpans (arrow/compute/executor.go:757-863) yields one span per min(remaining of each argument's
 * current chunk), executeSpans (:598-623) calls the kernel once per span into slices of one contiguous output; a
 * CallFunction executes on one goroutine (exec.go:164-170).
 */
#define _GNU_SOURCE
#include <sched.h>
#include <stdint.h>
#include <stddef.h>
#include <time.h>

typedef void (*arith_fn)(int type, int8_t op, const void* l, const void* r, void* out, int len);

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Pin the calling thread to `core` (>= 0); returns 0 on success. */
int bench_pin_to_core(int core) {
  cpu_set_t set;
  CPU_ZERO(&set);
  CPU_SET(core, &set);
  return sched_setaffinity(0, sizeof(set), &set);
}

/* One step = one Add over rows [lo, hi) of the chunked column (left chunks of lc rows, right chunks of rc rows).
 * Runs `warmup` untimed steps, then `steps` timed ones; step times (seconds) go to times[steps].  Returns spans/step. */
int64_t bench_add_f64_chunked(arith_fn fn, int type, int op, const double* l, const double* r, double* out, int64_t lo, int64_t hi,
                              int64_t lc, int64_t rc, int warmup, int steps, double* times) {
  int64_t spans = 0;
  for (int s = -warmup; s < steps; ++s) {
    const double t0 = now_s();
    int64_t pos = lo, k = 0;
    while (pos < hi) {
      int64_t ln = lc - pos % lc;
      const int64_t rr = rc - pos % rc;
      if (rr < ln) ln = rr;
      if (hi - pos < ln) ln = hi - pos;
      fn(type, (int8_t)op, l + pos, r + pos, out + pos, (int)ln);
      pos += ln;
      ++k;
    }
    if (s >= 0) times[s] = now_s() - t0;
    spans = k;
  }
  return spans;
}
