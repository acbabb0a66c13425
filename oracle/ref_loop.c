/*
 * ref_loop.c -- CPU ORACLE, TEST/BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline and --impl reference).
 *
 * Restates the reference's parallel restart driver, crates/optik/src/lib.rs:297-413: a pool of worker threads
 * pulls restart indices i from a shared counter (stand-in for rayon's work-stealing `(0..max_restarts)
 * .into_par_iter()`, :298-301), each runs one attempt from seed i (:360-370), successes are classified (:376-379),
 * Speed raises a shared should_exit flag (:381-384), and the result is selected as Quality = arg-min ||q-x0||
 * (:398-407) or Speed = lowest converged index (deterministic form of find_any, :409-412).
 * The inner solve is the fp64 LM twin (solver_twin.c); the reference's NLopt SLSQP is an un-vendored dependency
 * and cannot be built here -- every number from this file is a "port", not the Rust reference.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "twin_math.h"

#define MAX_DOF 32

/* from solver_twin.c */
typedef struct twin_params twin_params;
typedef struct twin_chain twin_chain;
int twin_chain_init(twin_chain* c, const double* chain, int njoints, const double* ee_offset);
int twin_attempt(const twin_chain* c, const twin_params* P, se3t tgt, const double* q_init, double* q_out,
                 double* f_out, int* evals_out);
int twin_status_success(const twin_params* P, int status);
void oracle_restart_seed(uint64_t restart, const double* lb, const double* ub, int n, double* q);
size_t twin_chain_sizeof(void);
int twin_chain_n(const twin_chain* c);
const double* twin_chain_slb(const twin_chain* c);
const double* twin_chain_sub(const twin_chain* c);

typedef struct {
  const twin_chain* c;
  const twin_params* P;
  se3t tgt;
  const double* x0;
  int n, mode;
  uint64_t r_begin, r_end;
  atomic_ullong next;
  atomic_int should_exit;
  /* per-worker results */
  int have;
  double best_score, best_f, best_q[MAX_DOF];
  uint64_t best_r;
  int best_status;
  uint64_t attempts, evals, converged;
} job_t;

typedef struct { job_t* shared; job_t local; } worker_t;

static void* worker_main(void* arg) {
  worker_t* w = (worker_t*)arg;
  job_t* S = w->shared;
  job_t* L = &w->local;
  L->have = 0; L->attempts = L->evals = L->converged = 0;
  for (;;) {
    if (S->mode == 2 && atomic_load_explicit(&S->should_exit, memory_order_relaxed)) break;
    uint64_t r = atomic_fetch_add(&S->next, 1);
    if (r >= S->r_end) break;
    double qi[MAX_DOF], q[MAX_DOF], f;
    int evals;
    if (r == 0) memcpy(qi, S->x0, sizeof(double) * S->n);
    else oracle_restart_seed(r, twin_chain_slb(S->c), twin_chain_sub(S->c), S->n, qi);
    int st = twin_attempt(S->c, S->P, S->tgt, qi, q, &f, &evals);
    L->attempts++; L->evals += (uint64_t)evals;
    if (twin_status_success(S->P, st)) {
      L->converged++;
      double score = 0.0;
      if (S->mode == 1) for (int j = 0; j < S->n; j++) score += (q[j] - S->x0[j]) * (q[j] - S->x0[j]);
      else score = (double)r;
      if (!L->have || score < L->best_score) {
        L->have = 1; L->best_score = score; L->best_f = f; L->best_r = r; L->best_status = st;
        memcpy(L->best_q, q, sizeof(double) * S->n);
      }
      if (S->mode == 2) atomic_store_explicit(&S->should_exit, 1, memory_order_relaxed);
    }
  }
  return NULL;
}

/* Robot::ik for one target with `threads` workers. stats[3] += {attempts, evaluations, converged}. Returns 1 if found. */
int ref_ik_threaded(const double* chain, int njoints, const double* ee_offset, const twin_params* P,
                    const double* target, const double* x0, uint64_t r_begin, uint64_t r_end, int mode, int threads,
                    double* q_out, double* f_out, uint64_t* restart_out, uint64_t* stats) {
  twin_chain* c = (twin_chain*)malloc(twin_chain_sizeof());
  if (twin_chain_init(c, chain, njoints, ee_offset)) { free(c); return -1; }
  job_t S;
  memset(&S, 0, sizeof(S));
  S.c = c; S.P = P; S.x0 = x0; S.n = twin_chain_n(c); S.mode = mode; S.r_begin = r_begin; S.r_end = r_end;
  S.tgt.q.x = target[0]; S.tgt.q.y = target[1]; S.tgt.q.z = target[2]; S.tgt.q.w = target[3];
  S.tgt.t = v3_make(target[4], target[5], target[6]);
  atomic_init(&S.next, r_begin);
  atomic_init(&S.should_exit, 0);
  if (threads < 1) threads = 1;
  worker_t* ws = (worker_t*)calloc((size_t)threads, sizeof(worker_t));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  for (int i = 0; i < threads; i++) { ws[i].shared = &S; pthread_create(&th[i], NULL, worker_main, &ws[i]); }
  int have = 0;
  double best = 0;
  for (int i = 0; i < threads; i++) {
    pthread_join(th[i], NULL);
    job_t* L = &ws[i].local;
    if (stats) { stats[0] += L->attempts; stats[1] += L->evals; stats[2] += L->converged; }
    if (L->have && (!have || L->best_score < best)) {
      have = 1; best = L->best_score;
      memcpy(q_out, L->best_q, sizeof(double) * S.n); *f_out = L->best_f; *restart_out = L->best_r;
    }
  }
  free(ws); free(th); free(c);
  return have;
}

/* ---- a batch of independent targets on the CPU: what a caller of the reference does with many IK problems
 * (examples/example.rs:23-36 loops over targets) when it wants throughput -- one target per worker thread, each
 * running Robot::ik with set_parallelism(1) semantics (restarts in index order, Speed stops at the first success,
 * lib.rs:397-413).  Workers pull target indices from a shared counter.  No speculative work is wasted, so this is
 * the most favourable CPU arrangement for a Speed-mode batch. */
int twin_ik_c(const double* chain, int njoints, const double* ee_offset, const twin_params* P, const double* target,
              const double* x0, uint64_t r_begin, uint64_t r_end, int mode, double* q_out, double* f_out,
              int* status_out, uint64_t* restart_out);

typedef struct {
  const double* chain; int njoints; const twin_params* P;
  const double* targets; const double* x0; int n; uint64_t T, R; int mode;
  double* q_out; double* f_out; int* found_out;
  atomic_ullong next;
} batch_t;

static void* batch_worker(void* arg) {
  batch_t* B = (batch_t*)arg;
  for (;;) {
    uint64_t t = atomic_fetch_add(&B->next, 1);
    if (t >= B->T) break;
    int st; uint64_t rs; double f;
    int have = twin_ik_c(B->chain, B->njoints, NULL, B->P, B->targets + 8 * t, B->x0 + (size_t)B->n * t, 0, B->R, B->mode,
                         B->q_out + (size_t)B->n * t, &f, &st, &rs);
    B->f_out[t] = f;
    B->found_out[t] = have > 0;
  }
  return NULL;
}

int ref_batch_threaded(const double* chain, int njoints, const twin_params* P, const double* targets, const double* x0, int n,
                       uint64_t T, uint64_t R, int mode, int threads, double* q_out, double* f_out, int* found_out) {
  batch_t B;
  memset(&B, 0, sizeof(B));
  B.chain = chain; B.njoints = njoints; B.P = P; B.targets = targets; B.x0 = x0; B.n = n; B.T = T; B.R = R; B.mode = mode;
  B.q_out = q_out; B.f_out = f_out; B.found_out = found_out;
  atomic_init(&B.next, 0);
  if (threads < 1) threads = 1;
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, batch_worker, &B);
  for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
  free(th);
  return 0;
}
