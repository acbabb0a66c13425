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

#include "dmath.cuh"
#include "solver_params.h"

namespace optik {

// ------------------------------------------------------------------ TMA staging
DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEV void stage_chain_tma(double* s_chain, uint64_t* s_bar, const double* g_chain, uint32_t bytes) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(s_bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(s_chain)),
                 "l"(g_chain), "r"(bytes), "r"(smem_u32(s_bar))
                 : "memory");
  }
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(s_bar)), "r"(0u)
        : "memory");
  } while (!ok);
}

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
  double lb, ub, slb, sub;
  int type;  // 0 revolute, 1 prismatic, 2 padding lane
};
DEV Joint load_joint(const double* s_chain, int lane, int n) {
  Joint J;
  if (lane < n) {
    const double* j = s_chain + OPTIK_CHAIN_STRIDE * lane;
    J.ot = mk3(j[0], j[1], j[2]);
    J.type = (int)j[3];
    J.oq.x = j[4]; J.oq.y = j[5]; J.oq.z = j[6]; J.oq.w = j[7];
    J.ax = mk3(j[8], j[9], j[10]);
    J.lb = j[12]; J.ub = j[13]; J.slb = j[14]; J.sub = j[15];
  } else {  // padding lane: identity transform, zero Jacobian column
    J.ot = mk3(0, 0, 0); J.ax = mk3(0, 0, 0);
    J.oq.x = J.oq.y = J.oq.z = 0; J.oq.w = 1;
    J.lb = J.ub = J.slb = J.sub = 0;
    J.type = 2;
  }
  return J;
}
DEV se3 load_pose8(const double* p) {
  se3 r;
  r.q.x = p[0]; r.q.y = p[1]; r.q.z = p[2]; r.q.w = p[3];
  r.t = mk3(p[4], p[5], p[6]);
  return r;
}
DEV v3 weight3(qt tq, const double* w, v3 u) {  // R_tgt^T diag(w) R_tgt u   (objective.rs:13-35)
  v3 a = qrot(tq, u);
  a = mk3(a.x * w[0], a.y * w[1], a.z * w[2]);
  return qrot_inv(tq, a);
}

// ------------------------------------------------------------------ one objective evaluation (all lanes)
// Restates kinematics.rs:123-196 (FK scan, body Jacobian), math.rs:40-203 (so3/se3 log, d-log), objective.rs:7-110
// (residual, weights, task Jacobian).  Outputs: f and r[6] (tile-uniform), Jr[6] (this lane's task-Jacobian column).
template <int TILE>
DEV void evaluate(const Joint& J, const se3& tip, const se3& tgt, const double* wl, const double* wa, int weighted,
                  int n, int lane, double q, double& f, double* r, double* Jr, se3* ee_out = nullptr) {
  // lane-local transform origin_j * motion_j(q_j)
  se3 T;
  {
    double s, c;
    dsincos(0.5 * q, s, c);
    qt qa;
    qa.x = J.ax.x * s; qa.y = J.ax.y * s; qa.z = J.ax.z * s; qa.w = c;
    const qt qrev = qmul(J.oq, qa);
    const v3 tpri = add3(J.ot, qrot(J.oq, scale3(J.ax, q)));
    T.q = (J.type == 0) ? qrev : J.oq;
    T.t = (J.type == 1) ? tpri : J.ot;
  }
  // Kogge-Stone inclusive scan of SE(3) products over the tile
#pragma unroll
  for (int d = 1; d < TILE; d <<= 1) {
    if (d < n) {  // tile-uniform: n is a kernel constant
      const se3 up = shfl_up_se3<TILE>(T, d);
      const se3 c = se3mul(up, T);
      if (lane >= d) T = c;
    }
  }
  const se3 ee = se3mul(shfl_se3<TILE>(T, n - 1), tip);
  if (ee_out) *ee_out = ee;
  // pose error X = T_tgt^-1 * T_ee   (objective.rs:48-49)
  const qt xq = qmul(qconj(tgt.q), ee.q);
  const v3 xt = qrot_inv(tgt.q, sub3(ee.t, tgt.t));
  double qw = xq.w;
  v3 v = mk3(xq.x, xq.y, xq.z);
  if (!(qw >= 0.0)) { qw = -qw; v = neg3(v); }  // double cover, math.rs:43-47
  const double vn2 = dot3(v, v);
  double k, th2, ce, bq;
  if (vn2 > 1e-6) {
    const double vn = sqrt(vn2);
    const double half = datan2_pos(vn, qw);
    const double inv_vn = 1.0 / vn, inv_half = 1.0 / half;
    k = half * inv_vn;
    const double p = k * qw;  // (theta/2)/tan(theta/2)  == 1/2 theta sin/(1-cos), math.rs:112-114
    const double it2 = 0.25 * (inv_half * inv_half);
    th2 = 4.0 * (half * half);
    ce = (1.0 - p) * it2;  // hat(w)^2 coefficient of V^-1 (math.rs:120-121), of J (math.rs:90-93) and a_q (math.rs:150)
    const double a = (vn * qw) * inv_half;  // sin(theta)/theta
    bq = fma((1.0 + a) * it2, 0.25 * (inv_vn * inv_vn), -2.0 * (it2 * it2));  // math.rs:151
  } else {  // Taylor branches, math.rs:55-60, 115-118, 153-158
    const double iw = 1.0 / qw, iw2 = iw * iw;
    k = iw * fma(vn2 * iw2, fma(vn2 * iw2, 0.2, -1.0 / 3.0), 1.0);
    th2 = 4.0 * ((k * k) * vn2);
    ce = fma(th2, fma(th2, 1.0 / 30240.0, 1.0 / 720.0), 1.0 / 12.0);
    bq = fma(th2, 1.0 / 7560.0, 1.0 / 360.0);
  }
  const v3 w = scale3(v, k + k);
  const v3 wxt = cross3(w, xt);
  const v3 elin = axpy3(ce, cross3(w, wxt), axpy3(-0.5, wxt, xt));  // V^-1 t
  v3 rl = elin, ra = w;
  if (weighted) { rl = weight3(tgt.q, wl, elin); ra = weight3(tgt.q, wa, w); }
  r[0] = rl.x; r[1] = rl.y; r[2] = rl.z; r[3] = ra.x; r[4] = ra.y; r[5] = ra.z;
  f = dot6(r, r);
  // Q = C*J scalars (math.rs:160-169)
  const double d = dot3(w, xt);
  const double kc = fma(th2, bq, ce + ce);
  const v3 cv = axpy3(bq * d, w, scale3(xt, -kc));
  const double da = d * ce;
  // this lane's body-Jacobian column (kinematics.rs:171-193) -> task column Jlog6 * col (objective.rs:79-81)
  const v3 axw = qrot(T.q, J.ax);
  const v3 lw = cross3(axw, sub3(ee.t, T.t));
  const v3 lin = qrot_inv(ee.q, (J.type == 0) ? lw : axw);
  v3 ang = qrot_inv(ee.q, axw);
  if (J.type != 0) ang = mk3(0, 0, 0);
  const v3 wxa = cross3(w, ang);
  const v3 ja = axpy3(ce, cross3(w, wxa), axpy3(0.5, wxa, ang));
  const v3 wxl = cross3(w, lin);
  const v3 jl = axpy3(ce, cross3(w, wxl), axpy3(0.5, wxl, lin));
  const double wu = dot3(w, ja), tu = dot3(xt, ja);
  const v3 cu = axpy3(da, ja, axpy3(ce * tu, w, axpy3(wu, cv, scale3(cross3(xt, ja), 0.5))));
  v3 top = add3(jl, cu), bot = ja;
  if (weighted) { top = weight3(tgt.q, wl, top); bot = weight3(tgt.q, wa, bot); }
  const bool live = (J.type != 2);
  Jr[0] = live ? top.x : 0.0; Jr[1] = live ? top.y : 0.0; Jr[2] = live ? top.z : 0.0;
  Jr[3] = live ? bot.x : 0.0; Jr[4] = live ? bot.y : 0.0; Jr[5] = live ? bot.z : 0.0;
}

// LDL^T solve of the SPD 6x6 system (lower triangle of A in packed row order a*(a+1)/2+b)
DEV void ldl6_solve(const double* Ap, const double* r, double* y) {
  double L[6][6], D[6], inv[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double dj = Ap[j * (j + 1) / 2 + j];
#pragma unroll
    for (int k = 0; k < j; k++) dj = fma(-(L[j][k] * L[j][k]), D[k], dj);
    D[j] = dj;
    inv[j] = 1.0 / dj;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double s = Ap[i * (i + 1) / 2 + j];
#pragma unroll
      for (int k = 0; k < j; k++) s = fma(-(L[i][k] * L[j][k]), D[k], s);
      L[i][j] = s * inv[j];
    }
  }
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double s = r[i];
#pragma unroll
    for (int k = 0; k < i; k++) s = fma(-L[i][k], z[k], s);
    z[i] = s;
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double s = z[i] * inv[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s = fma(-L[k][i], y[k], s);
    y[i] = s;
  }
}

DEV unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ------------------------------------------------------------------ the solve kernel
template <int TILE>
__global__ void __launch_bounds__(128) solve_kernel(const __grid_constant__ SolveParams P) {
  __shared__ alignas(128) double s_chain[OPTIK_MAX_DOF * OPTIK_CHAIN_STRIDE + 8];
  __shared__ alignas(8) uint64_t s_bar;
  stage_chain_tma(s_chain, &s_bar, P.chain, P.chain_bytes);

  const int n = P.n;
  const int lane = threadIdx.x % TILE;
  const unsigned long long tiles_per_block = blockDim.x / TILE;
  const unsigned long long total_tiles = tiles_per_block * gridDim.x;
  const unsigned long long njobs = P.T * (unsigned long long)P.C;
  const Joint J = load_joint(s_chain, lane, n);
  const se3 tip = se3mul(load_pose8(s_chain + OPTIK_CHAIN_STRIDE * n), load_pose8(P.ee_offset));
  const bool speed = (P.mode == 2);
  const unsigned long long t_start = P.max_ns ? globaltimer_ns() : 0ull;

  // ---- tile state (uniform across the tile unless marked "lane")
  unsigned long long job = blockIdx.x * tiles_per_block + threadIdx.x / TILE;
  unsigned long long tgt_id = 0, r_idx = 0;
  bool need_job = true, running = false, done = false;
  se3 tgt = tip;
  double x0 = 0.0;                                   // lane
  double qc = 0.0, qt_ = 0.0;                        // lane: current / trial joint value
  double fc = 0.0, rc[6] = {0, 0, 0, 0, 0, 0}, Jc[6] = {0, 0, 0, 0, 0, 0};  // current f, r, (lane) Jr column
  double lambda = P.lambda0;
  int have_cur = 0, slow = 0, evals = 0;
  // per-job best candidate
  bool best_has = false;
  double best_q = 0.0, best_f = 0.0, best_score = 0.0;  // best_q: lane
  unsigned long long best_r = 0;
  int best_status = OPTIK_ST_SKIPPED, job_evals = 0;
  unsigned long long n_attempts = 0, n_evals = 0, n_conv = 0;

  for (;;) {
    // ---------------- transitions: pick the next attempt for tiles that are idle (no shuffles in here)
    if (!running && !done) {
      for (;;) {
        if (need_job) {
          if (job >= njobs) { done = true; break; }
          tgt_id = job / P.C;
          r_idx = P.r_begin + job % P.C;
          tgt = load_pose8(P.targets + 8 * tgt_id);
          x0 = (lane < n) ? P.x0[tgt_id * n + lane] : 0.0;
          best_has = false; best_status = OPTIK_ST_SKIPPED; best_f = 0.0; best_score = 0.0; best_r = r_idx;
          best_q = x0; job_evals = 0;
          need_job = false;
        }
        bool skip = r_idx >= P.r_end;
        if (!skip && speed && P.found) skip = *((volatile unsigned long long*)(P.found + tgt_id)) < r_idx;
        if (!skip && P.max_ns) skip = (globaltimer_ns() - t_start) > P.max_ns;
        if (!skip) {  // start restart r_idx: restart 0 = caller's seed, i>=1 = ChaCha8 stream i (lib.rs:360-370)
          double q0 = x0;
          if (r_idx != 0) q0 = uniform_f64(chacha8_u64(P.key, (unsigned long long)(lane >> 3), r_idx, lane), J.slb, J.sub);
          qt_ = fmin(fmax(q0, J.lb), J.ub);
          have_cur = 0; slow = 0; evals = 0; lambda = P.lambda0;
          running = true;
          break;
        }
        // job finished: write its candidate record
        if (lane < n) P.cand_q[job * n + lane] = best_q;
        if (lane == 0) {
          P.cand_f[job] = best_f;
          P.cand_score[job] = best_score;
          P.cand_restart[job] = best_r;
          P.cand_status[job] = best_status;
          P.cand_evals[job] = job_evals;
        }
        job += total_tiles;
        need_job = true;
      }
    }
    if (__all_sync(FULL, done)) break;

    // ---------------- evaluate the trial point (every lane of the warp, uniform instruction stream)
    double ft, rt[6], Jt[6];
    evaluate<TILE>(J, tip, tgt, P.wl, P.wa, P.weighted, n, lane, qt_, ft, rt, Jt);

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
      if (success) {
        n_conv++;
        if (!best_has || score < best_score) {  // r_idx only grows within a chunk: ties keep the lower index
          best_has = true; best_q = qt_; best_f = ft; best_score = score; best_r = r_idx; best_status = status;
        }
        if (speed) {
          if (P.found && lane == 0) atomicMin(P.found + tgt_id, r_idx);
          r_idx = P.r_end;  // first success ends the chunk (lib.rs:381-387, 411)
        }
      } else if (!best_has) {
        best_q = qt_; best_f = ft; best_r = r_idx; best_status = status;
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
    double Ap[21];
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int b = 0; b <= a; b++) Ap[a * (a + 1) / 2 + b] = tile_sum<TILE>(m * (Jc[a] * Jc[b]));
#pragma unroll
    for (int a = 0; a < 6; a++) Ap[a * (a + 1) / 2 + a] = Ap[a * (a + 1) / 2 + a] + lambda;
    double y[6];
    ldl6_solve(Ap, rc, y);
    const double dq = -(m * dot6(Jc, y));
    qt_ = fmin(fmax(qc + dq, J.lb), J.ub);
  }

  if (P.counters && lane == 0) {
    atomicAdd(P.counters + 0, n_attempts);
    atomicAdd(P.counters + 1, n_evals);
    atomicAdd(P.counters + 2, n_conv);
  }
}

// ------------------------------------------------------------------ selection across chunks (lib.rs:397-413)
// One block per target: lexicographic min over the C candidates of (no-solution, score, restart index).
__global__ void __launch_bounds__(256) select_kernel(const __grid_constant__ SelectParams P) {
  const unsigned long long t = blockIdx.x;
  __shared__ double s_score[256];
  __shared__ unsigned long long s_restart[256];
  __shared__ unsigned int s_idx[256];
  __shared__ int s_has[256];
  __shared__ int s_evals[256];
  int has = 0, ev = 0;
  double score = 0.0;
  unsigned long long restart = ~0ull;
  unsigned int idx = 0;
  for (unsigned int c = threadIdx.x; c < P.C; c += blockDim.x) {
    const unsigned long long job = t * P.C + c;
    const int st = P.cand_status[job];
    const int ok = (P.tol_f >= 0.0 && st == OPTIK_ST_STOPVAL) || (P.tol_df_user >= 0.0 && st == OPTIK_ST_FTOL) ||
                   (P.tol_dx >= 0.0 && st == OPTIK_ST_XTOL);
    const double sc = P.cand_score[job];
    const unsigned long long rr = P.cand_restart[job];
    ev += P.cand_evals[job];
    const bool better = (ok > has) || (ok == has && (sc < score || (sc == score && rr < restart)));
    if (c == threadIdx.x || better) { has = ok; score = sc; restart = rr; idx = c; }
  }
  if (threadIdx.x >= P.C) { has = -1; }  // no candidate scanned
  s_has[threadIdx.x] = has; s_score[threadIdx.x] = score; s_restart[threadIdx.x] = restart; s_idx[threadIdx.x] = idx;
  s_evals[threadIdx.x] = ev;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      const int o = threadIdx.x + s;
      const bool better = (s_has[o] > s_has[threadIdx.x]) ||
                          (s_has[o] == s_has[threadIdx.x] &&
                           (s_score[o] < s_score[threadIdx.x] ||
                            (s_score[o] == s_score[threadIdx.x] && s_restart[o] < s_restart[threadIdx.x])));
      if (better) {
        s_has[threadIdx.x] = s_has[o]; s_score[threadIdx.x] = s_score[o];
        s_restart[threadIdx.x] = s_restart[o]; s_idx[threadIdx.x] = s_idx[o];
      }
      s_evals[threadIdx.x] += s_evals[o];
    }
    __syncthreads();
  }
  const unsigned long long win = t * P.C + s_idx[0];
  if (P.q_out)
    for (int j = threadIdx.x; j < P.n; j += blockDim.x) P.q_out[t * P.n + j] = P.cand_q[win * P.n + j];
  if (P.record_out) {
    double* rec = P.record_out + t * (8 + P.n);
    for (int j = threadIdx.x; j < P.n; j += blockDim.x) rec[8 + j] = P.cand_q[win * P.n + j];
    if (threadIdx.x == 0) {
      const unsigned long long rr = P.cand_restart[win];
      rec[0] = s_has[0] > 0 ? 1.0 : 0.0;
      rec[1] = (P.mode == 2) ? (double)rr : P.cand_score[win];
      rec[2] = (double)rr;
      rec[3] = P.cand_f[win];
      rec[4] = (double)P.cand_status[win];
      rec[5] = rec[6] = rec[7] = 0.0;
    }
  }
  if (threadIdx.x == 0 && P.f_out) {
    P.f_out[t] = P.cand_f[win];
    if (P.restart_out) P.restart_out[t] = P.cand_restart[win];
    P.status_out[t] = P.cand_status[win];
    if (P.evals_out) P.evals_out[t] = s_evals[0];
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
extern "C" int optik_launch_select(const SelectParams* p, void* stream) {
  optik::select_kernel<<<(unsigned)p->T, 256, 0, (cudaStream_t)stream>>>(*p);
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
