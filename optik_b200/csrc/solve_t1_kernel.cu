// solve_t1_kernel.cu -- the restart solve with ONE THREAD PER RESTART SEED (throughput layout, n <= 8).
//
// Same path, objective, stop rules, seeds and LM step as solve_kernel.cu (reference: crates/optik/src/lib.rs:297-413,
// objective.rs:40-110, kinematics.rs:123-196, math.rs:40-203), but the evaluation order is sequential per seed:
//   * one BACKWARD recursion in the target's frame on the inverse pose C_j = B_j^-1 (B_{j-1} = L_j B_j, B_n = tip)
//     yields joint j's body-Jacobian column [t_C x (R_C a_j); R_C a_j] = [R_Bj^T (a_j x p_Bj); R_Bj^T a_j] with one
//     rotation and one cross product, and the pose error X = C_0^-1  (as eval_kernel)
//   * on accept the body columns become task columns by two explicit 3x3 matrices (Jlog6 = [[J, C J],[0, J]], 27 fma
//     per column)
//   * columns live in a per-thread shared-memory row (128-bit accesses, odd 16-byte row stride => conflict-free);
//     two rows per thread: the current point's task columns and the trial point's body columns (swapped on accept)
//   * the 6x6 Gram matrix is accumulated joint by joint with fma; the LDL^T solve is thread-private.
// Why it exists: in the tile layout every lane of a tile repeats the tile-uniform math (log map, LDL^T), so a
// 7-DOF seed costs ~240 fp64 issue slots per evaluation; here it costs ~60.  The tile kernel remains the
// low-latency / long-chain layout (and the literal "one warp per seed" configuration); this one is the batch layout.
// Threads never communicate: no shuffles, no block barriers in the loop; a thread refills itself with the next
// (target, chunk) job when its attempt ends (flattened state machine), so lanes of a warp stay busy.
#include <cuda_runtime.h>

#include "ik_math.cuh"
#include "ldl6.cuh"
#include "solver_params.h"

namespace optik {

constexpr int T1_THREADS = 128;

DEV int t1_row_units(int n) { return (3 * n) | 1; }  // 16-byte units per column row, forced odd

// One ChaCha8 block -> the first 8 u64 of stream `stream` (enough for n <= 8 joints)
DEV void chacha8_block(const uint32_t* key, uint64_t stream, uint64_t* out8) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                    key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    0u, 0u, (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll 1  // rolled: the block is drawn once per attempt, its code should not crowd the evaluation loop out of the instruction cache
  for (int r = 0; r < 4; r++) {
    OPTIK_QR(x[0], x[4], x[8], x[12]) OPTIK_QR(x[1], x[5], x[9], x[13])
    OPTIK_QR(x[2], x[6], x[10], x[14]) OPTIK_QR(x[3], x[7], x[11], x[15])
    OPTIK_QR(x[0], x[5], x[10], x[15]) OPTIK_QR(x[1], x[6], x[11], x[12])
    OPTIK_QR(x[2], x[7], x[8], x[13]) OPTIK_QR(x[3], x[4], x[9], x[14])
  }
#pragma unroll
  for (int k = 0; k < 8; k++)
    out8[k] = (uint64_t)(x[2 * k] + s[2 * k]) | ((uint64_t)(x[2 * k + 1] + s[2 * k + 1]) << 32);
}

__global__ void __launch_bounds__(T1_THREADS, 2) solve_t1_kernel(const __grid_constant__ SolveParams P) {
  extern __shared__ __align__(128) double smem[];
  const int n = P.n;
  // layout: chain blob | tip^-1 pose8 | mbarrier (16 B) | rows[2][T1_THREADS][units] (16 B units) | qc[n][T] | qt[n][T] |
  //         qnext[n][T] | per-joint constants origin_q (x) (axis, 0) [n][4]
  double* s_chain = smem;
  double* s_tip = smem + OPTIK_CHAIN_STRIDE * n + 8;
  uint64_t* s_bar = (uint64_t*)(s_tip + 8);
  const int units = t1_row_units(n);
  double2* s_rows = (double2*)(s_tip + 10);
  double* s_qc = (double*)(s_rows + 2 * T1_THREADS * units);
  double* s_qt = s_qc + n * T1_THREADS;
  double* s_qn = s_qt + n * T1_THREADS;
  double* s_oa = s_qn + n * T1_THREADS;
  stage_chain_tma(s_chain, s_bar, P.chain, P.chain_bytes);
  if (threadIdx.x == 0) {  // (fixed tip joint * ee_offset)^-1, once per block
    const se3 tip = se3mul(load_pose8(s_chain + OPTIK_CHAIN_STRIDE * n), load_pose8(P.ee_offset));
    const qt iq = qconj(tip.q);
    const v3 it = neg3(qrot(iq, tip.t));
    s_tip[0] = iq.x; s_tip[1] = iq.y; s_tip[2] = iq.z; s_tip[3] = iq.w;
    s_tip[4] = it.x; s_tip[5] = it.y; s_tip[6] = it.z; s_tip[7] = 0.0;
  }
  if (threadIdx.x < n) {  // origin_q (x) (axis, 0): L_j.q = cos * origin_q + sin * this
    const double* jc = s_chain + OPTIK_CHAIN_STRIDE * threadIdx.x;
    qt oq, qa;
    oq.x = jc[4]; oq.y = jc[5]; oq.z = jc[6]; oq.w = jc[7];
    qa.x = jc[8]; qa.y = jc[9]; qa.z = jc[10]; qa.w = 0.0;
    const qt oa = qmul(oq, qa);
    double* d = s_oa + 4 * threadIdx.x;
    d[0] = oa.x; d[1] = oa.y; d[2] = oa.z; d[3] = oa.w;
  }
  __syncthreads();

  const int tid = threadIdx.x;
  double* qc = s_qc + tid;  // qc[j * T1_THREADS]
  double* qt_ = s_qt + tid;
  double* qn = s_qn + tid;  // prefetched seed of this thread's next attempt
  double2* const row0 = s_rows + (size_t)tid * units;  // rows[k] = row0 + k * row_stride
  const size_t row_stride = (size_t)T1_THREADS * units;
  const unsigned long long njobs = P.T * (unsigned long long)P.C;
  const bool speed = (P.mode == 2);
  const unsigned long long t_start = P.max_ns ? globaltimer_ns() : 0ull;
  constexpr unsigned FULL = 0xffffffffu;
  const bool prefetch = njobs >= 8ull * T1_THREADS * gridDim.x;

  // ---- per-thread state.  A "job" is one (target, chunk): restarts r_begin+c, +C, ... run in index order.
  unsigned long long job = 0, tgt_id = 0, src_id = 0, r_idx = 0, nxt_job = 0, nxt_r = 0;
  bool running = false, job_open = false, best_has = false, rec_any = false, have_next = false, queue_done = false, done = false;
  se3 O0;  // origin of joint 0 pre-multiplied by T_tgt^-1 (FK in the target's frame)
  O0.q.x = O0.q.y = O0.q.z = 0; O0.q.w = 1; O0.t = mk3(0, 0, 0);
  qt O0a = O0.q;  // O0.q (x) (axis_0, 0)
  double fc = 0.0, rc[6] = {0, 0, 0, 0, 0, 0}, lambda = P.lambda0, best_score = 0.0;
  int have_cur = 0, slow = 0, evals = 0, job_evals = 0, cur = 0;
  unsigned n_attempts = 0, n_evals = 0, n_conv = 0;

  for (;;) {
    // ---------------- seed pipeline.  Drawing a restart seed (one ChaCha8 block, ~800 integer instructions) by a lone
    // lane would cost a full warp issue slot per instruction, so seeds are PREFETCHED: whenever some lane is idle
    // without a prefetched attempt, every lane lacking one refills in the same pass (lib.rs:360-370 per lane).
    // With few jobs per thread (prefetch == 0) only idle lanes pull from the QUEUE, so that no lane hoards a job another
    // could run.
    if (__any_sync(FULL, !running && !have_next && !queue_done)) {
      // the next restart of a thread's OWN chunk can always be drawn ahead (nobody else could run it); only pulling
      // a new job off the queue early is restricted to launches with many jobs per thread
      const bool own_next = job_open && r_idx + P.C < P.r_end;
      if (!have_next && !queue_done && (prefetch || !running || own_next)) {
        bool got = false;
        if (own_next) {  // next restart of my own chunk
          nxt_job = job; nxt_r = r_idx + P.C; got = true;
        } else {
          nxt_job = atomicAdd(P.queue, 1ull);  // dynamic job queue
          if (nxt_job >= njobs) queue_done = true;
          else {
            const unsigned long long c = (P.C == 1) ? 0ull : (P.T == 1 ? nxt_job : nxt_job % P.C);
            nxt_r = P.r_begin + c;
            got = nxt_r < P.r_end;
            if (!got) {  // chunk without any restart (C > R): empty record, job consumed
              const unsigned long long t = (P.C == 1) ? nxt_job : (P.T == 1 ? 0ull : nxt_job / P.C);
              const unsigned long long ts = P.tlist ? (unsigned long long)P.tlist[t] : t;
              for (int j = 0; j < n; j++) P.cand_q[nxt_job * n + j] = P.x0[ts * n + j];
              P.cand_f[nxt_job] = 0.0; P.cand_score[nxt_job] = 0.0; P.cand_restart[nxt_job] = nxt_r;
              P.cand_status[nxt_job] = OPTIK_ST_SKIPPED; P.cand_evals[nxt_job] = 0;
            }
          }
        }
        if (got) {
          if (nxt_r != 0) {
            uint64_t u[8];
            chacha8_block(P.key, nxt_r, u);
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (j < n) {
                const double* jc = s_chain + OPTIK_CHAIN_STRIDE * j;
                qn[j * T1_THREADS] = fmin(fmax(uniform_f64(u[j], jc[14], jc[15]), jc[12]), jc[13]);
              }
          } else {
            const unsigned long long t = (P.C == 1) ? nxt_job : (P.T == 1 ? 0ull : nxt_job / P.C);
            const unsigned long long ts = P.tlist ? (unsigned long long)P.tlist[t] : t;
            for (int j = 0; j < n; j++) {
              const double* jc = s_chain + OPTIK_CHAIN_STRIDE * j;
              qn[j * T1_THREADS] = fmin(fmax(P.x0[ts * n + j], jc[12]), jc[13]);
            }
          }
          have_next = true;
        }
      }
    }
    // ---------------- transitions of idle threads: start the prefetched attempt
    if (!running && !done) {
      if (have_next) {
        have_next = false;
        if (!job_open || nxt_job != job) {  // open a new job
          if (job_open) P.cand_evals[job] = job_evals;
          job = nxt_job;
          tgt_id = (P.C == 1) ? job : (P.T == 1 ? 0ull : job / P.C);
          src_id = P.tlist ? (unsigned long long)P.tlist[tgt_id] : tgt_id;  // row of targets / x0 (phased batches)
          const se3 tgt = load_pose8(P.targets + 8 * src_id);
          se3 ti, o0;
          ti.q = qconj(tgt.q);
          ti.t = neg3(qrot(ti.q, tgt.t));
          o0.q.x = s_chain[4]; o0.q.y = s_chain[5]; o0.q.z = s_chain[6]; o0.q.w = s_chain[7];
          o0.t = mk3(s_chain[0], s_chain[1], s_chain[2]);
          O0 = se3mul(ti, o0);
          qt a0;
          a0.x = s_chain[8]; a0.y = s_chain[9]; a0.z = s_chain[10]; a0.w = 0.0;
          O0a = qmul(O0.q, a0);
          best_has = false; rec_any = false; best_score = 0.0; job_evals = 0;
          for (int j = 0; j < n; j++) P.cand_q[job * n + j] = P.x0[src_id * n + j];  // record if no attempt runs
          P.cand_f[job] = 0.0; P.cand_score[job] = 0.0; P.cand_restart[job] = nxt_r;
          P.cand_status[job] = OPTIK_ST_SKIPPED;
          job_open = true;
        }
        r_idx = nxt_r;
        bool skip = false;
        if (speed && P.found) skip = *((volatile unsigned long long*)(P.found + tgt_id)) < r_idx;
        if (!skip && P.max_ns) skip = (globaltimer_ns() - t_start) > P.max_ns;
        if (skip) {  // every later restart of this chunk is skipped too: the job ends
          P.cand_evals[job] = job_evals;
          job_open = false;
        } else {
          for (int j = 0; j < n; j++) qt_[j * T1_THREADS] = qn[j * T1_THREADS];
          have_cur = 0; slow = 0; evals = 0; lambda = P.lambda0;
          running = true;
        }
      } else if (queue_done) {
        if (job_open) { P.cand_evals[job] = job_evals; job_open = false; }
        done = true;
      }
    }
    if (__all_sync(FULL, done)) break;
    if (!running) continue;

    // ---------------- evaluate the trial point: backward recursion in the target's frame
    double2* trow = row0 + (size_t)(cur ^ 1) * row_stride;  // trial body columns
    se3 Ci = load_pose8(s_tip);  // C_n = tip^-1
#pragma unroll 1
    for (int j = n - 1; j >= 0; j--) {
      const double2* jc = (const double2*)(s_chain + OPTIK_CHAIN_STRIDE * j);
      const double2 c1 = jc[1], c4 = jc[4], c5 = jc[5];
      const int type = (int)c1.y;
      const v3 ax = mk3(c4.x, c4.y, c5.x);
      v3 lin, ang = qrot(Ci.q, ax);
      if (type == 0) lin = cross3(Ci.t, ang);
      else { lin = ang; ang = mk3(0, 0, 0); }
      trow[3 * j + 0] = make_double2(lin.x, lin.y);
      trow[3 * j + 1] = make_double2(lin.z, ang.x);
      trow[3 * j + 2] = make_double2(ang.y, ang.z);
      se3 O;
      qt oa;
      if (j == 0) { O = O0; oa = O0a; }
      else {
        const double2 c0 = jc[0], c2 = jc[2], c3 = jc[3];
        const double2 a0 = ((const double2*)s_oa)[2 * j], a1 = ((const double2*)s_oa)[2 * j + 1];
        O.t = mk3(c0.x, c0.y, c1.x);
        O.q.x = c2.x; O.q.y = c2.y; O.q.z = c3.x; O.q.w = c3.y;
        oa.x = a0.x; oa.y = a0.y; oa.z = a1.x; oa.w = a1.y;
      }
      const double qj = qt_[j * T1_THREADS];
      qt lq;  // conj(L_j.q)
      v3 lt = O.t;
      if (type == 0) {
        double s, c;
        dsincos(0.5 * qj, s, c);
        lq.x = -fma(c, O.q.x, s * oa.x); lq.y = -fma(c, O.q.y, s * oa.y); lq.z = -fma(c, O.q.z, s * oa.z);
        lq.w = fma(c, O.q.w, s * oa.w);
      } else {
        lq = qconj(O.q);
        lt = add3(O.t, qrot(O.q, scale3(ax, qj)));
      }
      Ci.q = qmul(Ci.q, lq);
      Ci.t = sub3(Ci.t, qrot(Ci.q, lt));
    }
    se3 B;  // X = C_0^-1
    B.q = qconj(Ci.q);
    B.t = neg3(qrot(B.q, Ci.t));
    ErrCoef ec;
    v3 elin;
    error_terms(B.q, B.t, ec, elin);
    v3 rl = elin, ra = ec.w;
    qt tq;
    if (P.weighted) {
      const double* tp = P.targets + 8 * src_id;
      tq.x = tp[0]; tq.y = tp[1]; tq.z = tp[2]; tq.w = tp[3];
      rl = weight3(tq, P.wl, elin); ra = weight3(tq, P.wa, ec.w);
    }
    double rt[6] = {rl.x, rl.y, rl.z, ra.x, ra.y, ra.z};
    const double ft = dot6(rt, rt);

    // ---------------- bookkeeping (mirrors NLopt's stop tests as the reference configures them, lib.rs:345-347)
    int status = OPTIK_ST_NONE;
    bool accept = false;
    evals++;
    if (ft != ft) status = OPTIK_ST_NAN;
    else if (ft < P.tol_f) status = OPTIK_ST_STOPVAL;
    else if (!have_cur) accept = true;
    else if (ft < fc) {
      accept = true;
      const double df = fc - ft;
      if (df < P.tol_df_eff) status = OPTIK_ST_FTOL;
      else if (P.tol_dx > 0.0) {
        double dx = 0.0;
        for (int j = 0; j < n; j++) dx = fmax(dx, fabs(qt_[j * T1_THREADS] - qc[j * T1_THREADS]));
        if (dx < P.tol_dx) status = OPTIK_ST_XTOL;
      }
      slow = (df < P.stall_rel * fc) ? slow + 1 : 0;
      if (status == OPTIK_ST_NONE && slow >= P.stall_count) status = OPTIK_ST_STUCK;
      lambda = fmax(lambda * P.lambda_dec, P.lambda_min);
    } else {
      lambda = lambda * P.lambda_inc;
      if (lambda > P.lambda_max) status = OPTIK_ST_STUCK;
    }
    if (status == OPTIK_ST_NONE && evals >= P.max_evals) status = OPTIK_ST_ITERCAP;
    if (status == OPTIK_ST_NONE && P.max_ns && (globaltimer_ns() - t_start) > P.max_ns) status = OPTIK_ST_SKIPPED;
    if (status == OPTIK_ST_NONE && speed && P.found && *((volatile unsigned long long*)(P.found + tgt_id)) < r_idx)
      status = OPTIK_ST_SKIPPED;  // should_exit (lib.rs:308)

    if (status != OPTIK_ST_NONE) {  // attempt over
      const bool success = (P.tol_f >= 0.0 && status == OPTIK_ST_STOPVAL) ||
                           (P.tol_df_user >= 0.0 && status == OPTIK_ST_FTOL) ||
                           (P.tol_dx >= 0.0 && status == OPTIK_ST_XTOL);  // lib.rs:376-379
      double score = 0.0;
      if (success && !speed)  // Quality score ||q - x0||^2 (lib.rs:402-407)
        for (int j = 0; j < n; j++) {
          const double d = qt_[j * T1_THREADS] - P.x0[src_id * n + j];
          score = fma(d, d, score);
        }
      n_attempts++; n_evals += evals; job_evals += evals;
      const bool record = success ? (!best_has || score < best_score) : !rec_any;  // failures: keep the first
      if (record) {
        rec_any = true;
        for (int j = 0; j < n; j++) P.cand_q[job * n + j] = qt_[j * T1_THREADS];
        P.cand_f[job] = ft; P.cand_score[job] = score; P.cand_restart[job] = r_idx; P.cand_status[job] = status;
      }
      if (success) {
        n_conv++;
        if (record) { best_has = true; best_score = score; }
        if (speed) {  // first success ends the chunk (lib.rs:381-387, 411)
          if (P.found) atomicMin(P.found + tgt_id, r_idx);
          P.cand_evals[job] = job_evals;
          job_open = false;
          if (have_next && nxt_job == job) have_next = false;  // the prefetched restart of this chunk is moot
        }
      }
      running = false;
      continue;
    }

    if (accept) {  // current point <- trial point; body columns -> task columns, in place
      for (int j = 0; j < n; j++) qc[j * T1_THREADS] = qt_[j * T1_THREADS];
      fc = ft; have_cur = 1;
#pragma unroll
      for (int i = 0; i < 6; i++) rc[i] = rt[i];
      double Jm[9], CJ[9];
      task_mats(ec, Jm, CJ);
#pragma unroll 1
      for (int j = 0; j < n; j++) {
        const double2 a0 = trow[3 * j + 0], a1 = trow[3 * j + 1], a2 = trow[3 * j + 2];
        v3 top, bot;
        task_col_m(Jm, CJ, mk3(a0.x, a0.y, a1.x), mk3(a1.y, a2.x, a2.y), top, bot);
        if (P.weighted) { top = weight3(tq, P.wl, top); bot = weight3(tq, P.wa, bot); }
        trow[3 * j + 0] = make_double2(top.x, top.y);
        trow[3 * j + 1] = make_double2(top.z, bot.x);
        trow[3 * j + 2] = make_double2(bot.y, bot.z);
      }
      cur ^= 1;
    }

    // ---------------- LM step from the current point: y = (Jm Jm^T + lambda I)^-1 r ; dq = -Jm^T y ; project on bounds
    const double2* crow = row0 + (size_t)cur * row_stride;
    double Ap[21];
#pragma unroll
    for (int e = 0; e < 21; e++) Ap[e] = 0.0;
    unsigned free_mask = 0;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
      const double2 a0 = crow[3 * j + 0], a1 = crow[3 * j + 1], a2 = crow[3 * j + 2];
      const double c[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
      const double* jc = s_chain + OPTIK_CHAIN_STRIDE * j;
      const double g = dot6(rc, c);
      const double qcj = qc[j * T1_THREADS];
      const bool pinned = (qcj <= jc[12] && g > 0.0) || (qcj >= jc[13] && g < 0.0);
      const double m = pinned ? 0.0 : 1.0;
      free_mask |= pinned ? 0u : (1u << j);
#pragma unroll
      for (int a = 0; a < 6; a++) {
        const double jm = m * c[a];
#pragma unroll
        for (int b = 0; b <= a; b++) Ap[a * (a + 1) / 2 + b] = fma(jm, c[b], Ap[a * (a + 1) / 2 + b]);
      }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) Ap[a * (a + 1) / 2 + a] = Ap[a * (a + 1) / 2 + a] + lambda;
    double y[6];
    ldl6_solve(Ap, rc, y);
#pragma unroll 1
    for (int j = 0; j < n; j++) {
      const double2 a0 = crow[3 * j + 0], a1 = crow[3 * j + 1], a2 = crow[3 * j + 2];
      const double m = ((free_mask >> j) & 1u) ? 1.0 : 0.0;
      const double Jm[6] = {m * a0.x, m * a0.y, m * a1.x, m * a1.y, m * a2.x, m * a2.y};
      const double* jc = s_chain + OPTIK_CHAIN_STRIDE * j;
      qt_[j * T1_THREADS] = fmin(fmax(qc[j * T1_THREADS] - dot6(Jm, y), jc[12]), jc[13]);
    }
  }

  if (P.counters) {  // one atomic triple per warp: every thread of the warp reaches this point exactly once
    __syncwarp();
    unsigned a = n_attempts, e = n_evals, c = n_conv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); e += __shfl_xor_sync(0xffffffffu, e, o); c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(P.counters + 0, (unsigned long long)a);
      atomicAdd(P.counters + 1, (unsigned long long)e);
      atomicAdd(P.counters + 2, (unsigned long long)c);
    }
  }
}

}  // namespace optik

extern "C" int optik_t1_smem_bytes(int n) {
  const size_t units = (size_t)((3 * n) | 1);
  return (int)(sizeof(double) * (OPTIK_CHAIN_STRIDE * n + 8 + 8 + 2) + 16 * 2 * optik::T1_THREADS * units +
               sizeof(double) * 3 * n * optik::T1_THREADS + sizeof(double) * 4 * n);
}
extern "C" int optik_launch_solve_t1(const SolveParams* p, int blocks, void* stream) {
  const int smem = optik_t1_smem_bytes(p->n);
  cudaError_t e = cudaFuncSetAttribute(optik::solve_t1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  optik::solve_t1_kernel<<<blocks, optik::T1_THREADS, smem, (cudaStream_t)stream>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int optik_solve_t1_occupancy(int n, int* blocks_per_sm) {
  const int smem = optik_t1_smem_bytes(n);
  cudaError_t e = cudaFuncSetAttribute(optik::solve_t1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, optik::solve_t1_kernel, optik::T1_THREADS, smem);
}
