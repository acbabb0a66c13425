// solve_kernel.cu -- the random-restart IK solve as ONE sm_100a kernel.
//
// Replaces, in the reference (kylc/optik @ 355e463):
//   crates/optik/src/lib.rs:297-395   rayon fan-out over restart index i            -> CUDA grid of tiles
//   crates/optik/src/lib.rs:302-372   per-restart NLopt SLSQP solve                  -> in-tile projected LM (dual 6x6)
//   crates/optik/src/lib.rs:360-370   ChaCha8Rng(42).set_stream(i) seed draw         -> chacha8_u64 per lane
//   crates/optik/src/lib.rs:305-337   objective callback = FK + gradient + cost     -> evaluate() below
//   crates/optik/src/lib.rs:376-387   success classification, should_exit flag       -> status codes, `found[]`
//   crates/optik/src/lib.rs:397-413   Quality min_by_key / Speed find_any            -> per-chunk best + select_kernel
//
// Mapping: a TILE of 8/16/32 lanes owns one restart attempt; lane j owns joint j.
//   FK        = lane-local origin_j*R(axis_j,q_j), then a Kogge-Stone scan of SE(3) products over shuffles
//   Jacobian  = one body-frame column per lane, mapped through Jlog6 (and the weights) in registers
//   LM step   = (J J^T + lambda I) y = r solved redundantly in every lane after a butterfly all-reduce of the
//               21 unique entries of J J^T  (identical to the n x n normal equations by the push-through identity)
// The chain (128 B per joint) is staged once per block into shared memory by a 1-D TMA bulk copy.
// TILE=32 is "one warp per restart seed"; TILE=8 packs four 7-DOF attempts into a warp.  All tiles of a warp run
// one uniform instruction stream (a flattened state machine), so packed tiles never serialise each other.
// Tensor cores are deliberately unused: 6 x n is not a dense contraction.
#include <cuda_runtime.h>

#include "ik_math.cuh"
#include "solver_params.h"

#ifndef OPTIK_SOLVE_MIN_BLOCKS
#define OPTIK_SOLVE_MIN_BLOCKS 3
#endif

namespace optik {

// ------------------------------------------------------------------ tile collectives
constexpr unsigned FULL = 0xffffffffu;
template <int TILE>
DEV double tile_sum(double v) {  // xor butterfly, low strides first: every lane gets the same pairwise tree sum
#pragma unroll
  for (int s = 1; s < TILE; s <<= 1) v = v + __shfl_xor_sync(FULL, v, s, TILE);
  return v;
}
template <int TILE>
DEV double tile_max(double v) {
#pragma unroll
  for (int s = 1; s < TILE; s <<= 1) v = fmax(v, __shfl_xor_sync(FULL, v, s, TILE));
  return v;
}
template <int TILE>
DEV se3 shfl_up_se3(se3 T, int d) {
  se3 r;
  r.q.x = __shfl_up_sync(FULL, T.q.x, d, TILE); r.q.y = __shfl_up_sync(FULL, T.q.y, d, TILE);
  r.q.z = __shfl_up_sync(FULL, T.q.z, d, TILE); r.q.w = __shfl_up_sync(FULL, T.q.w, d, TILE);
  r.t.x = __shfl_up_sync(FULL, T.t.x, d, TILE); r.t.y = __shfl_up_sync(FULL, T.t.y, d, TILE);
  r.t.z = __shfl_up_sync(FULL, T.t.z, d, TILE);
  return r;
}
template <int TILE>
DEV se3 shfl_se3(se3 T, int src) {
  se3 r;
  r.q.x = __shfl_sync(FULL, T.q.x, src, TILE); r.q.y = __shfl_sync(FULL, T.q.y, src, TILE);
  r.q.z = __shfl_sync(FULL, T.q.z, src, TILE); r.q.w = __shfl_sync(FULL, T.q.w, src, TILE);
  r.t.x = __shfl_sync(FULL, T.t.x, src, TILE); r.t.y = __shfl_sync(FULL, T.t.y, src, TILE);
  r.t.z = __shfl_sync(FULL, T.t.z, src, TILE);
  return r;
}

// ------------------------------------------------------------------ per-lane joint constants
struct Joint {
  v3 ot, ax;
  qt oq;
  double lb, ub;
  int type;  // 0 revolute, 1 prismatic, 2 fixed transform (tip lane / padding lane)
};
// tip handling rule (depends on n only, so every tile width gives the same bits): the fixed tip transform rides
// the scan in lane n unless n is exactly a tile width (8, 16, 32), where no spare lane exists.
DEV bool tip_in_scan(int n) { return !(n == 8 || n == 16 || n == 32); }

// lane < n: joint record; lane == n (when the tip rides the scan): origin = tip; lanes beyond: identity
DEV Joint load_joint(const double* s_chain, const se3& tip, int lane, int n) {
  Joint J;
  J.ot = mk3(0, 0, 0); J.ax = mk3(0, 0, 0);
  J.oq.x = J.oq.y = J.oq.z = 0; J.oq.w = 1;
  J.lb = J.ub = 0;
  J.type = 2;
  if (lane < n) {
    const double* j = s_chain + OPTIK_CHAIN_STRIDE * lane;
    J.ot = mk3(j[0], j[1], j[2]);
    J.type = (int)j[3];
    J.oq.x = j[4]; J.oq.y = j[5]; J.oq.z = j[6]; J.oq.w = j[7];
    J.ax = mk3(j[8], j[9], j[10]);
    J.lb = j[12]; J.ub = j[13];
  } else if (lane == n && tip_in_scan(n)) {
    J.oq = tip.q; J.ot = tip.t;
  }
  return J;
}
// ------------------------------------------------------------------ one objective evaluation (all lanes)
// Restates kinematics.rs:123-196 (FK scan, body Jacobian), math.rs:40-203 (so3/se3 log, d-log), objective.rs:7-110
// (residual, weights, task Jacobian).  Outputs: f and r[6] (tile-uniform), Jr[6] (this lane's task-Jacobian column).
// FK runs IN THE TARGET'S FRAME: lane 0's origin was pre-multiplied by T_tgt^-1 when the job started, so the scan
// yields the pose error X = T_tgt^-1 * T_ee (objective.rs:48-49) directly; the body-frame Jacobian does not depend on
// the world frame.  `tgt_q` (the target rotation) is only read on the weighted path.
template <int TILE>
DEV void evaluate(const Joint& J, const se3& O, const double* s_tip, const double* tgt_q, const double* wl,
                  const double* wa, int weighted, int has_prismatic, int n, int lane, double q, double& f, double* r,
                  double* Jr) {
  // lane-local transform origin_j * motion_j(q_j)
  se3 T;
  {
    double s, c;
    dsincos(0.5 * q, s, c);
    qt qa;
    qa.x = J.ax.x * s; qa.y = J.ax.y * s; qa.z = J.ax.z * s; qa.w = c;
    const qt qrev = qmul(O.q, qa);
    T.q = (J.type == 0) ? qrev : O.q;
    T.t = O.t;
    if (has_prismatic) {  // grid-uniform
      const v3 tpri = add3(O.t, qrot(O.q, scale3(J.ax, q)));
      if (J.type == 1) T.t = tpri;
    }
  }
  // Kogge-Stone inclusive scan of SE(3) products over the tile
  const bool fold = tip_in_scan(n);
  const int m = fold ? n + 1 : n;
#pragma unroll
  for (int d = 1; d < TILE; d <<= 1) {
    if (d < m) {  // grid-uniform: n is a kernel constant
      const se3 up = shfl_up_se3<TILE>(T, d);
      const se3 c = se3mul(up, T);
      if (lane >= d) T = c;
    }
  }
  se3 ee = shfl_se3<TILE>(T, m - 1);
  if (!fold) ee = se3mul(ee, load_pose8(s_tip));
  // pose error X == ee (target frame)
  ErrCoef ec;
  v3 elin;
  error_terms(ee.q, ee.t, ec, elin);
  v3 rl = elin, ra = ec.w;
  qt tq;
  if (weighted) {
    tq.x = tgt_q[0]; tq.y = tgt_q[1]; tq.z = tgt_q[2]; tq.w = tgt_q[3];
    rl = weight3(tq, wl, elin); ra = weight3(tq, wa, ec.w);
  }
  r[0] = rl.x; r[1] = rl.y; r[2] = rl.z; r[3] = ra.x; r[4] = ra.y; r[5] = ra.z;
  f = dot6(r, r);
  // this lane's body-Jacobian column (kinematics.rs:171-193) -> task column Jlog6 * col (objective.rs:79-81)
  const v3 axw = qrot(T.q, J.ax);
  const v3 lw = cross3(axw, sub3(ee.t, T.t));
  v3 lin = qrot_inv(ee.q, lw);
  v3 ang = qrot_inv(ee.q, axw);
  if (has_prismatic && J.type == 1) { lin = ang; ang = mk3(0, 0, 0); }
  v3 top, bot;
  task_col(ec, lin, ang, top, bot);
  if (weighted) { top = weight3(tq, wl, top); bot = weight3(tq, wa, bot); }
  const bool live = (J.type != 2);
  Jr[0] = live ? top.x : 0.0; Jr[1] = live ? top.y : 0.0; Jr[2] = live ? top.z : 0.0;
  Jr[3] = live ? bot.x : 0.0; Jr[4] = live ? bot.y : 0.0; Jr[5] = live ? bot.z : 0.0;
}

#include "ldl6.cuh"

// ------------------------------------------------------------------ selection key (lib.rs:397-413)
struct SelKey {
  int has;  // 1 converged, 0 not, -1 nothing scanned
  int evals;
  double score;
  unsigned long long restart;
  unsigned long long idx;  // candidate index (job)
};
DEV bool sel_better(const SelKey& a, const SelKey& b) {  // a better than b
  return (a.has > b.has) || (a.has == b.has && (a.score < b.score || (a.score == b.score && a.restart < b.restart)));
}

// ------------------------------------------------------------------ the solve kernel
template <int TILE>
__global__ void __launch_bounds__(128, OPTIK_SOLVE_MIN_BLOCKS) solve_kernel(const __grid_constant__ SolveParams P) {
  __shared__ alignas(128) double s_chain[OPTIK_MAX_DOF * OPTIK_CHAIN_STRIDE + 8];
  __shared__ alignas(16) double s_tip[8];
  __shared__ alignas(8) uint64_t s_bar;
  stage_chain_tma(s_chain, &s_bar, P.chain, P.chain_bytes);

  const int n = P.n;
  const int lane = threadIdx.x % TILE;
  const unsigned long long njobs = P.T * (unsigned long long)P.C;
  if (threadIdx.x == 0) {  // tip = fixed tip joint * ee_offset, once per block
    const se3 tip = se3mul(load_pose8(s_chain + OPTIK_CHAIN_STRIDE * n), load_pose8(P.ee_offset));
    s_tip[0] = tip.q.x; s_tip[1] = tip.q.y; s_tip[2] = tip.q.z; s_tip[3] = tip.q.w;
    s_tip[4] = tip.t.x; s_tip[5] = tip.t.y; s_tip[6] = tip.t.z; s_tip[7] = 0.0;
  }
  __syncthreads();
  const Joint J = load_joint(s_chain, load_pose8(s_tip), lane, n);
  const bool speed = (P.mode == 2);
  const unsigned long long t_start = P.max_ns ? globaltimer_ns() : 0ull;

  // ---- tile state (uniform across the tile unless marked "lane")
  unsigned long long job = 0;
  // lanes of this tile (shuffles inside the divergent transition code name exactly the tile's lanes)
  const unsigned tile_mask = (TILE == 32) ? 0xffffffffu : (((1u << (TILE & 31)) - 1u) << (((threadIdx.x & 31) / TILE) * TILE));
  unsigned long long tgt_id = 0, src_id = 0, r_idx = 0;
  bool need_job = true, running = false, done = false, best_has = false, rec_any = false;
  se3 O;                          // lane: this lane's origin (lane 0: pre-multiplied by T_tgt^-1 per job)
  O.q = J.oq; O.t = J.ot;
  double x0 = 0.0;                // lane
  double qc = 0.0, qt_ = 0.0;     // lane: current / trial joint value
  double fc = 0.0, rc[6] = {0, 0, 0, 0, 0, 0}, Jc[6] = {0, 0, 0, 0, 0, 0};  // current f, r, (lane) Jr column
  double lambda = P.lambda0, best_score = 0.0;
  int have_cur = 0, slow = 0, evals = 0, job_evals = 0;
  unsigned n_attempts = 0, n_evals = 0, n_conv = 0;

  for (;;) {
    // ---------------- transitions: pick the next attempt for tiles that are idle (no shuffles in here)
    if (!running && !done) {
      for (;;) {
        if (need_job) {
          unsigned long long next = 0;
          if (lane == 0) next = atomicAdd(P.queue, 1ull);  // dynamic job queue, one fetch per tile
          job = __shfl_sync(tile_mask, next, 0, TILE);
          if (job >= njobs) { done = true; break; }
          tgt_id = job / P.C;
          src_id = tgt_id;
          r_idx = P.r_begin + job % P.C;
          x0 = (lane < n) ? P.x0[src_id * n + lane] : 0.0;
          if (lane == 0) {  // FK in the target's frame: O_0 <- T_tgt^-1 * origin_0
            const se3 tgt = load_pose8(P.targets + 8 * src_id);
            se3 ti;
            ti.q = qconj(tgt.q);
            ti.t = neg3(qrot(ti.q, tgt.t));
            se3 o0;
            o0.q = J.oq; o0.t = J.ot;
            O = se3mul(ti, o0);
          }
          best_has = false; rec_any = false; best_score = 0.0; job_evals = 0;
          if (lane < n) P.cand_q[job * n + lane] = x0;  // record of a chunk that runs no attempt
          if (lane == 0) {
            P.cand_f[job] = 0.0; P.cand_score[job] = 0.0; P.cand_restart[job] = r_idx;
            P.cand_status[job] = OPTIK_ST_SKIPPED;
          }
          need_job = false;
        }
        bool skip = r_idx >= P.r_end;
        if (!skip && speed && P.found) skip = *((volatile unsigned long long*)(P.found + tgt_id)) < r_idx;
        if (!skip && P.max_ns) skip = (globaltimer_ns() - t_start) > P.max_ns;
        if (!skip) {  // start restart r_idx: restart 0 = caller's seed, i>=1 = ChaCha8 stream i (lib.rs:360-370)
          double q0 = x0;
          if (r_idx != 0 && lane < n) {
            const double* j = s_chain + OPTIK_CHAIN_STRIDE * lane;
            q0 = uniform_f64(chacha8_u64(P.key, (unsigned long long)(lane >> 3), r_idx, lane), j[14], j[15]);
          }
          qt_ = fmin(fmax(q0, J.lb), J.ub);
          have_cur = 0; slow = 0; evals = 0; lambda = P.lambda0;
          running = true;
          break;
        }
        if (lane == 0) P.cand_evals[job] = job_evals;  // job finished
        need_job = true;
      }
    }
    if (__all_sync(FULL, done)) break;

    // ---------------- evaluate the trial point (every lane of the warp, uniform instruction stream)
    double ft, rt[6], Jt[6];
    evaluate<TILE>(J, O, s_tip, P.targets + 8 * src_id, P.wl, P.wa, P.weighted, P.has_prismatic, n, lane, qt_, ft, rt, Jt);

    // ---------------- bookkeeping (mirrors NLopt's stop tests as the reference configures them, lib.rs:345-347)
    int status = OPTIK_ST_NONE;
    bool accept = false;
    double dxmax = 0.0;
    if (P.tol_dx > 0.0) dxmax = tile_max<TILE>(fabs(qt_ - qc));  // grid-uniform branch
    if (running) {
      evals++;
      if (ft != ft) status = OPTIK_ST_NAN;
      else if (ft < P.tol_f) status = OPTIK_ST_STOPVAL;
      else if (!have_cur) accept = true;
      else if (ft < fc) {
        accept = true;
        const double df = fc - ft;
        if (df < P.tol_df_eff) status = OPTIK_ST_FTOL;
        else if (P.tol_dx > 0.0 && dxmax < P.tol_dx) status = OPTIK_ST_XTOL;
        slow = (df < P.stall_rel * fc) ? slow + 1 : 0;
        if (status == OPTIK_ST_NONE && slow >= P.stall_count) status = OPTIK_ST_STUCK;
        lambda = fmax(lambda * P.lambda_dec, P.lambda_min);
      } else {
        lambda = lambda * P.lambda_inc;
        if (lambda > P.lambda_max) status = OPTIK_ST_STUCK;
      }
      if (status == OPTIK_ST_NONE && evals >= P.max_evals) status = OPTIK_ST_ITERCAP;
      if (status == OPTIK_ST_NONE && P.max_ns && (globaltimer_ns() - t_start) > P.max_ns) status = OPTIK_ST_SKIPPED;
      // should_exit (lib.rs:308): a lower restart index of this target already converged
      if (status == OPTIK_ST_NONE && speed && P.found &&
          *((volatile unsigned long long*)(P.found + tgt_id)) < r_idx)
        status = OPTIK_ST_SKIPPED;
    }
    const bool success = (P.tol_f >= 0.0 && status == OPTIK_ST_STOPVAL) ||
                         (P.tol_df_user >= 0.0 && status == OPTIK_ST_FTOL) ||
                         (P.tol_dx >= 0.0 && status == OPTIK_ST_XTOL);  // lib.rs:376-379
    // Quality score ||q - x0||^2 (lib.rs:402-407); warp-uniform branch so the butterfly stays convergent
    double score = 0.0;
    if (!speed && __any_sync(FULL, success)) {
      const double dq0 = (lane < n) ? (qt_ - x0) : 0.0;
      score = tile_sum<TILE>(dq0 * dq0);
    }
    if (status != OPTIK_ST_NONE) {  // attempt over
      n_attempts++; n_evals += evals; job_evals += evals;
      // r_idx only grows within a chunk, so on score ties the lower index is kept
      const bool record = success ? (!best_has || score < best_score) : !rec_any;  // failures: keep the first
      if (record) {  // the chunk's candidate record lives in HBM, not in registers
        rec_any = true;
        if (lane < n) P.cand_q[job * n + lane] = qt_;
        if (lane == 0) {
          P.cand_f[job] = ft; P.cand_score[job] = score; P.cand_restart[job] = r_idx; P.cand_status[job] = status;
        }
      }
      if (success) {
        n_conv++;
        if (record) { best_has = true; best_score = score; }
        if (speed) {
          if (P.found && lane == 0) atomicMin(P.found + tgt_id, r_idx);
          r_idx = P.r_end;  // first success ends the chunk (lib.rs:381-387, 411)
        }
      }
      if (r_idx < P.r_end) r_idx += P.C;
      running = false;
    }
    if (accept) {
      qc = qt_; fc = ft; have_cur = 1;
#pragma unroll
      for (int i = 0; i < 6; i++) { rc[i] = rt[i]; Jc[i] = Jt[i]; }
    }

    // ---------------- LM step from the current point: y = (Jm Jm^T + lambda I)^-1 r ; dq = -Jm^T y ; project on bounds
    const double g = dot6(rc, Jc);
    const bool pinned = (qc <= J.lb && g > 0.0) || (qc >= J.ub && g < 0.0);
    const double m = (pinned || lane >= n) ? 0.0 : 1.0;
    double Jm[6];
#pragma unroll
    for (int a = 0; a < 6; a++) Jm[a] = m * Jc[a];
    double Ap[21];
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int b = 0; b <= a; b++) Ap[a * (a + 1) / 2 + b] = tile_sum<TILE>(Jm[a] * Jc[b]);
#pragma unroll
    for (int a = 0; a < 6; a++) Ap[a * (a + 1) / 2 + a] = Ap[a * (a + 1) / 2 + a] + lambda;
    double y[6];
    ldl6_solve(Ap, rc, y);
    const double dq = -dot6(Jm, y);
    qt_ = fmin(fmax(qc + dq, J.lb), J.ub);
  }

  if (P.counters && lane == 0) {
    atomicAdd(P.counters + 0, (unsigned long long)n_attempts);
    atomicAdd(P.counters + 1, (unsigned long long)n_evals);
    atomicAdd(P.counters + 2, (unsigned long long)n_conv);
  }

  // ---------------- fused selection for single-target launches (Robot::ik): the LAST block to finish scans the candidate
  // records, writes the packed record -- straight into mapped host memory when the host polls for it --, re-arms the
  // persistent control words for the slot's next call and raises the completion flag.  One launch per ik() wave.
  if (P.fused_record) {
    __shared__ SelKey s_sel[128];
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(P.fused_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      SelKey me;
      me.has = -1; me.evals = 0; me.score = 0.0; me.restart = ~0ull; me.idx = 0;
      for (unsigned long long c = threadIdx.x; c < njobs; c += blockDim.x) {
        const int st = ((volatile int*)P.cand_status)[c];
        SelKey k;
        k.has = (P.tol_f >= 0.0 && st == OPTIK_ST_STOPVAL) || (P.tol_df_user >= 0.0 && st == OPTIK_ST_FTOL) ||
                (P.tol_dx >= 0.0 && st == OPTIK_ST_XTOL);
        k.score = ((volatile double*)P.cand_score)[c];
        k.restart = ((volatile unsigned long long*)P.cand_restart)[c];
        k.idx = c; k.evals = 0;
        if (me.has < 0 || sel_better(k, me)) me = k;
      }
      s_sel[threadIdx.x] = me;
      __syncthreads();
      for (int w = 64; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
          const SelKey o = s_sel[threadIdx.x + w];
          if (o.has >= 0 && (s_sel[threadIdx.x].has < 0 || sel_better(o, s_sel[threadIdx.x]))) s_sel[threadIdx.x] = o;
        }
        __syncthreads();
      }
      const SelKey w = s_sel[0];
      const int len = 8 + n;
      if ((int)threadIdx.x < len) {
        const int t = threadIdx.x;
        double v = 0.0;
        if (t == 0) v = w.has > 0 ? 1.0 : 0.0;
        else if (t == 1) v = speed ? (double)w.restart : w.score;
        else if (t == 2) v = (double)w.restart;
        else if (t == 3) v = ((volatile double*)P.cand_f)[w.idx];
        else if (t == 4) v = (double)((volatile int*)P.cand_status)[w.idx];
        else if (t >= 8) v = ((volatile double*)P.cand_q)[w.idx * n + (t - 8)];
        P.fused_record[t] = v;
      }
      if (P.fused_reset && threadIdx.x == 0) {  // persistent per-slot control words: ready for the next launch
        *P.queue = 0ull;
        *P.fused_done = 0u;
        if (P.found) P.found[0] = ~0ull;
      }
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0 && P.fused_flag) *((volatile unsigned long long*)P.fused_flag) = P.fused_seq;
    }
  }
}

// ------------------------------------------------------------------ selection across chunks (lib.rs:397-413)
// Lexicographic min over the C candidates of a target of (no-solution, score, restart index).  Grid = (T, S): each
// block reduces one slice of the candidates; with S > 1 (one target, tens of thousands of chunks) the slice winners go
// to a partial array and a second launch (P.final_pass) reduces the S partials -- a single block walking 65 536
// records serially costs more than the solve itself.
__global__ void __launch_bounds__(256) select_kernel(const __grid_constant__ SelectParams P) {
  const unsigned long long t = blockIdx.x;
  const unsigned S = gridDim.y, slice = blockIdx.y;
  __shared__ SelKey s_key[256];
  SelKey me;
  me.has = -1; me.evals = 0; me.score = 0.0; me.restart = ~0ull; me.idx = t * P.C;
  if (!P.final_pass) {
    const unsigned long long per = ((unsigned long long)P.C + S - 1) / S;
    const unsigned long long c0 = slice * per, c1 = (c0 + per < P.C) ? c0 + per : P.C;
    for (unsigned long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
      const unsigned long long job = t * P.C + c;
      const int st = P.cand_status[job];
      SelKey k;
      k.has = (P.tol_f >= 0.0 && st == OPTIK_ST_STOPVAL) || (P.tol_df_user >= 0.0 && st == OPTIK_ST_FTOL) ||
              (P.tol_dx >= 0.0 && st == OPTIK_ST_XTOL);
      k.score = P.cand_score[job];
      k.restart = P.cand_restart[job];
      k.idx = job;
      k.evals = 0;
      me.evals += P.cand_evals[job];
      if (me.has < 0 || sel_better(k, me)) { const int ev = me.evals; me = k; me.evals = ev; }
    }
  } else {  // reduce the S partial winners of target t
    for (unsigned s = threadIdx.x; s < P.partials; s += blockDim.x) {
      const SelKey k = P.partial[t * P.partials + s];
      const int ev = me.evals + k.evals;
      if (k.has >= 0 && (me.has < 0 || sel_better(k, me))) me = k;
      me.evals = ev;
    }
  }
  s_key[threadIdx.x] = me;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      const SelKey o = s_key[threadIdx.x + s];
      SelKey m = s_key[threadIdx.x];
      const int ev = m.evals + o.evals;
      if (o.has >= 0 && (m.has < 0 || sel_better(o, m))) m = o;
      m.evals = ev;
      s_key[threadIdx.x] = m;
    }
    __syncthreads();
  }
  const SelKey w = s_key[0];
  if (!P.final_pass && S > 1) {  // slice winner -> partial array
    if (threadIdx.x == 0) P.partial_out[t * S + slice] = w;
    return;
  }
  const unsigned long long win = w.idx;
  const unsigned long long ot = t;
  if (P.q_out)
    for (int j = threadIdx.x; j < P.n; j += blockDim.x) P.q_out[ot * P.n + j] = P.cand_q[win * P.n + j];
  if (P.record_out) {
    double* rec = P.record_out + ot * (8 + P.n);
    for (int j = threadIdx.x; j < P.n; j += blockDim.x) rec[8 + j] = P.cand_q[win * P.n + j];
    if (threadIdx.x == 0) {
      const unsigned long long rr = P.cand_restart[win];
      rec[0] = w.has > 0 ? 1.0 : 0.0;
      rec[1] = (P.mode == 2) ? (double)rr : P.cand_score[win];
      rec[2] = (double)rr;
      rec[3] = P.cand_f[win];
      rec[4] = (double)P.cand_status[win];
      rec[5] = rec[6] = rec[7] = 0.0;
    }
  }
  if (threadIdx.x == 0 && P.f_out) {
    P.f_out[ot] = P.cand_f[win];
    if (P.restart_out) P.restart_out[ot] = P.cand_restart[win];
    P.status_out[ot] = P.cand_status[win];
    if (P.evals_out) P.evals_out[ot] = w.evals;
  }
}

}  // namespace optik

namespace optik {
// Best-pick over gathered candidate records (one warp): converged first, lowest score, lowest restart index.
__global__ void select_records_kernel(const double* __restrict__ rec, unsigned count, int n, double* __restrict__ out) {
  const int len = 8 + n;
  const unsigned lane = threadIdx.x;
  double has = -1.0, score = 0.0, restart = 0.0;
  unsigned idx = 0;
  for (unsigned c = lane; c < count; c += 32) {
    const double h = rec[c * len + 0], s = rec[c * len + 1], r = rec[c * len + 2];
    const bool better = (h > has) || (h == has && (s < score || (s == score && r < restart)));
    if (better) { has = h; score = s; restart = r; idx = c; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double h2 = __shfl_xor_sync(0xffffffffu, has, o), s2 = __shfl_xor_sync(0xffffffffu, score, o),
                 r2 = __shfl_xor_sync(0xffffffffu, restart, o);
    const unsigned i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    const bool better = (h2 > has) || (h2 == has && (s2 < score || (s2 == score && (r2 < restart || (r2 == restart && i2 < idx)))));
    if (better) { has = h2; score = s2; restart = r2; idx = i2; }
  }
  for (int j = lane; j < len; j += 32) out[j] = rec[idx * len + j];
}
}  // namespace optik

namespace optik {
// Device-pointer calls cannot reject a seed outside the joint limits (the reference panics, lib.rs:251-254) without a
// host sync: the kernels clamp it, and this pass marks the target's status (OPTIK_STATUS_FLAG_SEED_CLAMPED = 0x100).
__global__ void flag_clamped_kernel(const double* __restrict__ chain, int n, const double* __restrict__ x0, unsigned long long T,
                                    int* __restrict__ status, unsigned long long status_stride) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  bool out = false;
  for (int j = 0; j < n; j++) {
    const double q = x0[t * n + j];
    out |= !(q >= chain[OPTIK_CHAIN_STRIDE * j + 12] && q <= chain[OPTIK_CHAIN_STRIDE * j + 13]);
  }
  if (out) status[t * status_stride] |= 0x100;
}
}  // namespace optik
extern "C" int optik_launch_flag_clamped(const double* chain, int n, const double* x0, unsigned long long T, int* status,
                                         unsigned long long status_stride, void* stream) {
  optik::flag_clamped_kernel<<<(unsigned)((T + 255) / 256), 256, 0, (cudaStream_t)stream>>>(chain, n, x0, T, status, status_stride);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ host launchers (called from robot.cpp)
extern "C" int optik_launch_select_records(const double* rec, unsigned count, int n, double* out, void* stream) {
  optik::select_records_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(rec, count, n, out);
  return (int)cudaGetLastError();
}
extern "C" int optik_launch_solve(const SolveParams* p, int tile, int blocks, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = 128;
  switch (tile) {
    case 8: optik::solve_kernel<8><<<blocks, threads, 0, s>>>(*p); break;
    case 16: optik::solve_kernel<16><<<blocks, threads, 0, s>>>(*p); break;
    case 32: optik::solve_kernel<32><<<blocks, threads, 0, s>>>(*p); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}
extern "C" int optik_select_partial_bytes(void) { return (int)sizeof(optik::SelKey); }
// slices == 1: one launch.  slices > 1: caller provides partial scratch (T*slices*optik_select_partial_bytes()).
extern "C" int optik_launch_select(const SelectParams* p, unsigned slices, void* partial_scratch, void* stream) {
  SelectParams P = *p;
  P.final_pass = 0; P.partials = 0; P.partial = nullptr; P.partial_out = nullptr;
  if (slices <= 1) {
    optik::select_kernel<<<dim3((unsigned)P.T, 1), 256, 0, (cudaStream_t)stream>>>(P);
    return (int)cudaGetLastError();
  }
  P.partial_out = (optik::SelKey*)partial_scratch;
  optik::select_kernel<<<dim3((unsigned)P.T, slices), 256, 0, (cudaStream_t)stream>>>(P);
  P.final_pass = 1; P.partials = slices; P.partial = (const optik::SelKey*)partial_scratch; P.partial_out = nullptr;
  optik::select_kernel<<<dim3((unsigned)P.T, 1), 256, 0, (cudaStream_t)stream>>>(P);
  return (int)cudaGetLastError();
}
extern "C" int optik_solve_occupancy(int tile, int* blocks_per_sm) {
  const int threads = 128;
  switch (tile) {
    case 8: return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, optik::solve_kernel<8>, threads, 0);
    case 16: return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, optik::solve_kernel<16>, threads, 0);
    case 32: return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, optik::solve_kernel<32>, threads, 0);
  }
  return (int)cudaErrorInvalidValue;
}
