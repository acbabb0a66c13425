// eval_kernel.cu -- batched FK / body Jacobian / pose error / gradient ("K_eval"), one thread per configuration.
//
// User-facing batch form of the evaluator the reference exposes one configuration at a time:
//   Robot::fk              crates/optik/src/lib.rs:97-99   -> kinematics.rs:123-164
//   Robot::joint_jacobian  crates/optik/src/lib.rs:93-95   -> kinematics.rs:166-196
//   objective              crates/optik/src/objective.rs:40-57
//   objective_grad         crates/optik/src/objective.rs:60-110
//
// This is the HBM-bound kernel of the design: per configuration it reads q (8n B) + target (64 B) and writes the
// pose (64 B), the 6 x n Jacobian (48n B), f (8 B) and the gradient (8n B) = 8(8n+17) bytes in fp64.
//
// Algorithm: a single BACKWARD recursion over the joints carries C_j = B_j^-1, the inverse of the end-effector pose in
// joint j's frame (B_{j-1} = L_j B_j, B_n = tip, L_j = origin_j * motion_j(q_j)); the body-frame Jacobian column of
// joint j is [ t_C x (R_C axis_j) ; R_C axis_j ] = [ R_Bj^T (axis_j x p_Bj) ; R_Bj^T axis_j ]  (kinematics.rs:171-193)
// and C_0^-1 is the end-effector pose -- no per-joint transforms are stored, so the register footprint is independent
// of n.  Inputs: each warp streams its tiles of 32 joint vectors into shared memory with one TMA bulk load per tile,
// issued one tile ahead; targets are loaded into registers before the recursion that hides their latency.  Outputs:
// every thread stages its 6n-double Jacobian row in shared memory with 128-bit stores (row stride = an odd number of
// 16-byte units, so quarter-warps never bank-conflict), forms the gradient from it, and hands the row to the TMA: one
// `cp.async.bulk.global.shared::cta` per thread writes the 48n contiguous bytes to HBM while the thread already works
// on its next configuration.  No block barrier in the loop.  The chain is staged once per block with a TMA bulk load.
#include <cuda_runtime.h>

#include "dmath.cuh"
#include "solver_params.h"

namespace optik {

DEV uint32_t e_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }


DEV se3 e_load_pose8(const double* p) {
  se3 r;
  r.q.x = p[0]; r.q.y = p[1]; r.q.z = p[2]; r.q.w = p[3];
  r.t = mk3(p[4], p[5], p[6]);
  return r;
}
DEV v3 e_weight3(qt tq, const double* w, v3 u) {
  v3 a = qrot(tq, u);
  a = mk3(a.x * w[0], a.y * w[1], a.z * w[2]);
  return qrot_inv(tq, a);
}

DEV int eval_row_units(int n) { return (3 * n) | 1; }  // 16-byte units per staged row, forced odd

DEV void e_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// one 1-D TMA bulk load global -> shared, completion counted on `bar`
DEV void e_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// N = compile-time joint count (fully unrolled joint loops: static shared-memory offsets, the scheduler overlaps one
// joint's sin/cos with the previous joint's quaternion chain); N = 0 reads n from the parameters.
// TB = threads per block: 128, or 64 for long chains (n > 12) whose 48n-byte rows would leave one 128-thread block per SM
// (n = 20: 4 warps/SM with 128 threads, 6 with 64).
// COLS = Jacobian columns are needed (jac_out or grad_out): per-thread rows in shared memory, ONE tile buffer per warp
// (refilled while the warp forms objective / gradient / stores).  COLS = false is the FK-only launch: no rows, TWO
// tile buffers per warp, the next tile is requested before the current one is consumed.
template <int N, bool COLS, int TB>
__global__ void __launch_bounds__(TB, TB == 128 ? 4 : 6) eval_kernel(const __grid_constant__ EvalParams P) {
  extern __shared__ __align__(128) double smem[];
  // layout: chain blob | block mbarrier (16 B) | two mbarriers per warp (8 B each) | per-thread Jacobian rows (COLS) |
  //         per-warp joint-vector tiles [(TB / 32)][NBUF][32 * n] | per-joint constants origin_q (x) axis [n][4]
  constexpr int NBUF = COLS ? 1 : 2;
  double* s_chain = smem;
  const int n = N ? N : P.n;
  const int chain_doubles = OPTIK_CHAIN_STRIDE * n + 8;
  uint64_t* s_bar = (uint64_t*)(smem + chain_doubles);
  uint64_t* s_wbar = (uint64_t*)(smem + chain_doubles + 2);
  double2* s_rows = (double2*)(smem + chain_doubles + 2 + 2 * (TB / 32));
  const int units = eval_row_units(n);
  double* s_q = (double*)(s_rows + (COLS ? (size_t)TB * units : 0));
  double2* s_oa = (double2*)(s_q + (size_t)(TB / 32) * NBUF * 32 * n);  // [n][2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(e_smem_u32(s_bar)));
#pragma unroll
    for (int w = 0; w < 2 * (TB / 32); w++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(e_smem_u32(s_wbar + w)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) e_bulk_load(e_smem_u32(s_chain), P.chain, P.chain_bytes, e_smem_u32(s_bar));

  // ---- input pipeline: every warp streams ITS tiles of 32 joint vectors (32 * n contiguous doubles of q) into
  // shared memory with one TMA bulk load per tile; the load of tile k+1 is issued as soon as the recursion of tile k
  // has consumed the buffer and lands while the warp forms objective / gradient / stores of tile k.  No thread ever
  // waits on a dependent global load of q inside the joint loop.  Ragged last tile / unaligned q: plain loads.
  double* const qb0 = s_q + (size_t)warp * NBUF * 32 * n;  // buffer b at qb0 + b * 32 * n, barrier b at wbar0 + 8 b
  const uint32_t wbar0 = e_smem_u32(s_wbar + 2 * warp);
  const uint32_t tile_bytes = 256u * (uint32_t)n;
  const bool q_aligned = (((unsigned long long)P.q) & 15ull) == 0ull;
  const unsigned long long stride = (unsigned long long)gridDim.x * TB;
  unsigned long long base = (unsigned long long)blockIdx.x * TB + 32ull * warp;  // first configuration of my tile
  uint32_t parity = 0, pending = 0;  // bit b: phase parity of barrier b / a bulk load into buffer b is in flight
  auto fetch_tile = [&](unsigned long long b, int buf) {
    double* qb = qb0 + buf * 32 * n;
    if (q_aligned && b + 32 <= P.B) {
      if (lane == 0) e_bulk_load(e_smem_u32(qb), P.q + b * n, tile_bytes, wbar0 + 8u * buf);
      pending |= 1u << buf;
      return;
    }
    if (b + lane < P.B)
      for (int j = 0; j < n; j++) qb[lane * n + j] = P.q[(b + lane) * n + j];
  };
  if (base < P.B) fetch_tile(base, 0);
  int buf = 0;
  e_mbar_wait(e_smem_u32(s_bar), 0);  // chain staged

  if (threadIdx.x < n) {  // per-joint constant origin_q (x) (axis, 0)
    const double* jc = s_chain + OPTIK_CHAIN_STRIDE * threadIdx.x;
    qt oq, qa;
    oq.x = jc[4]; oq.y = jc[5]; oq.z = jc[6]; oq.w = jc[7];
    qa.x = jc[8]; qa.y = jc[9]; qa.z = jc[10]; qa.w = 0.0;
    const qt oa = qmul(oq, qa);
    s_oa[2 * threadIdx.x] = make_double2(oa.x, oa.y);
    s_oa[2 * threadIdx.x + 1] = make_double2(oa.z, oa.w);
  }
  __syncthreads();
  const se3 tip = se3mul(e_load_pose8(s_chain + OPTIK_CHAIN_STRIDE * n), e_load_pose8(P.ee_offset));
  se3 tip_inv;
  tip_inv.q = qconj(tip.q);
  tip_inv.t = neg3(qrot(tip_inv.q, tip.t));
  const bool want_obj = (P.f_out != nullptr) || (P.grad_out != nullptr);
  constexpr bool want_cols = COLS;
  double2* my = s_rows + (size_t)threadIdx.x * units;
  const uint32_t my_addr = e_smem_u32(my);
  const uint32_t row_bytes = 48u * (uint32_t)n;

  for (; base < P.B; base += stride) {
    const double* q = qb0 + buf * 32 * n + lane * n;
    if (NBUF == 2 && base + stride < P.B) fetch_tile(base + stride, buf ^ 1);  // consumed one iteration ago
    const unsigned long long i = base + lane;
    const bool active = i < P.B;
    // my previous row must have been read out by the TMA before it is overwritten
    if (P.jac_out) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    // the target is not needed before the recursion ends: its loads fly underneath it
    double2 t0 = make_double2(0, 0), t1 = make_double2(0, 1), t2 = make_double2(0, 0), t3 = make_double2(0, 0);
    if (want_obj && active) {
      const double2* tp = (const double2*)(P.targets + (unsigned long long)P.target_stride * i);
      t0 = tp[0]; t1 = tp[1]; t2 = tp[2]; t3 = tp[3];
    }
    if (pending & (1u << buf)) {
      e_mbar_wait(wbar0 + 8u * buf, (parity >> buf) & 1u);
      parity ^= 1u << buf;
      pending &= ~(1u << buf);
    }
    // Backward recursion on the INVERSE pose C_j = B_j^-1 = (R_B^T, -R_B^T p_B): the column of joint j is then
    //   ang = R_C axis_j ,  lin = t_C x ang      [ = R_B^T (axis_j x p_B) ]
    // (one rotation + one cross product instead of two rotations), and  C_{j-1} = C_j L_j^-1  with
    //   L_j.q = cos * origin_q + sin * (origin_q (x) axis)   (the second factor is a per-joint constant, s_oa)
    //   C.q <- C.q (x) conj(L.q) ,  C.t <- C.t - R_C.q(new) origin_t
    se3 Cinv = tip_inv;
    if (active) {
#pragma unroll
      for (int j = n - 1; j >= 0; j--) {
        const double2* jc = (const double2*)(s_chain + OPTIK_CHAIN_STRIDE * j);
        const double2 c0 = jc[0], c1 = jc[1], c2 = jc[2], c3 = jc[3], c4 = jc[4], c5 = jc[5];
        v3 ot = mk3(c0.x, c0.y, c1.x);
        const int type = (int)c1.y;
        const v3 ax = mk3(c4.x, c4.y, c5.x);
        if (want_cols) {
          v3 ang = qrot(Cinv.q, ax), lin;
          if (type == 0) lin = cross3(Cinv.t, ang);
          else { lin = ang; ang = mk3(0, 0, 0); }
          my[3 * j + 0] = make_double2(lin.x, lin.y);
          my[3 * j + 1] = make_double2(lin.z, ang.x);
          my[3 * j + 2] = make_double2(ang.y, ang.z);
        }
        const double qj = q[j];
        qt lq;  // conj(L.q)
        if (type == 0) {
          double s, c;
          dsincos(0.5 * qj, s, c);
          const double2 a0 = s_oa[2 * j], a1 = s_oa[2 * j + 1];
          lq.x = -fma(c, c2.x, s * a0.x); lq.y = -fma(c, c2.y, s * a0.y); lq.z = -fma(c, c3.x, s * a1.x);
          lq.w = fma(c, c3.y, s * a1.y);
        } else {
          lq.x = -c2.x; lq.y = -c2.y; lq.z = -c3.x; lq.w = c3.y;
          qt oq;
          oq.x = c2.x; oq.y = c2.y; oq.z = c3.x; oq.w = c3.y;
          ot = add3(ot, qrot(oq, scale3(ax, qj)));
        }
        Cinv.q = qmul(Cinv.q, lq);
        Cinv.t = sub3(Cinv.t, qrot(Cinv.q, ot));
      }
    }
    se3 B;  // end-effector pose = C_0^-1
    B.q = qconj(Cinv.q);
    B.t = neg3(qrot(B.q, Cinv.t));
    // the tile buffer is free once every lane has finished its recursion: stream the warp's next tile into it
    __syncwarp();
    if (NBUF == 1) { if (base + stride < P.B) fetch_tile(base + stride, 0); }
    else buf ^= 1;
    if (!active) continue;
    if (P.jac_out) {  // hand the staged row to the TMA: 48n contiguous bytes of HBM per configuration
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(P.jac_out + i * 6ull * n),
                   "r"(my_addr), "r"(row_bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (P.ee_out) {
      double2* o = (double2*)(P.ee_out + 8 * i);
      o[0] = make_double2(B.q.x, B.q.y); o[1] = make_double2(B.q.z, B.q.w);
      o[2] = make_double2(B.t.x, B.t.y); o[3] = make_double2(B.t.z, 0.0);
    }
    if (want_obj) {
      se3 tgt;
      tgt.q.x = t0.x; tgt.q.y = t0.y; tgt.q.z = t1.x; tgt.q.w = t1.y;
      tgt.t = mk3(t2.x, t2.y, t3.x);
      const qt xq = qmul(qconj(tgt.q), B.q);
      const v3 xt = qrot_inv(tgt.q, sub3(B.t, tgt.t));
      double qw = xq.w;
      v3 v = mk3(xq.x, xq.y, xq.z);
      if (!(qw >= 0.0)) { qw = -qw; v = neg3(v); }
      const double vn2 = dot3(v, v);
      double k, th2, ce, bq;
      if (vn2 > 1e-6) {
        const double vn = sqrt(vn2);
        const double half = datan2_pos(vn, qw);
        const double inv_vn = 1.0 / vn, inv_half = 1.0 / half;
        k = half * inv_vn;
        const double p = k * qw;
        const double it2 = 0.25 * (inv_half * inv_half);
        th2 = 4.0 * (half * half);
        ce = (1.0 - p) * it2;
        const double a = (vn * qw) * inv_half;
        bq = fma((1.0 + a) * it2, 0.25 * (inv_vn * inv_vn), -2.0 * (it2 * it2));
      } else {
        const double iw = 1.0 / qw, iw2 = iw * iw;
        k = iw * fma(vn2 * iw2, fma(vn2 * iw2, 0.2, -1.0 / 3.0), 1.0);
        th2 = 4.0 * ((k * k) * vn2);
        ce = fma(th2, fma(th2, 1.0 / 30240.0, 1.0 / 720.0), 1.0 / 12.0);
        bq = fma(th2, 1.0 / 7560.0, 1.0 / 360.0);
      }
      const v3 w = scale3(v, k + k);
      const v3 wxt = cross3(w, xt);
      const v3 elin = axpy3(ce, cross3(w, wxt), axpy3(-0.5, wxt, xt));
      v3 rl = elin, ra = w;      // W e    (objective.rs:52)
      v3 gl = elin, ga = w;      // W^2 e  (objective.rs:102-104)
      if (P.weighted) {
        rl = e_weight3(tgt.q, P.wl, elin); ra = e_weight3(tgt.q, P.wa, w);
        gl = e_weight3(tgt.q, P.wl, rl);   ga = e_weight3(tgt.q, P.wa, ra);
      }
      if (P.f_out) P.f_out[i] = dot3(rl, rl) + dot3(ra, ra);
      if (P.grad_out) {
        // u = Jlog6^T (W^2 e):  Jlog6 = [[J, Q],[0, J]],  J^T x = x - 1/2 w x x + ce w x (w x x),  Q = C J
        // u_lin = J^T gl ; u_ang = Q^T gl + J^T ga = J^T (C^T gl + ga)
        const double d = dot3(w, xt);
        const double kc = fma(th2, bq, ce + ce);
        const v3 cv = axpy3(bq * d, w, scale3(xt, -kc));
        // C^T x = -1/2 t x x + w (cv.x) + t ce (w.x) + d ce x
        const v3 ctg = axpy3(d * ce, gl, axpy3(ce * dot3(w, gl), xt, axpy3(dot3(cv, gl), w, scale3(cross3(xt, gl), -0.5))));
        const v3 s = add3(ctg, ga);
        const v3 wxg = cross3(w, gl);
        const v3 ul = axpy3(ce, cross3(w, wxg), axpy3(-0.5, wxg, gl));
        const v3 wxs = cross3(w, s);
        const v3 ua = axpy3(ce, cross3(w, wxs), axpy3(-0.5, wxs, s));
        double* g = P.grad_out + i * n;
#pragma unroll
        for (int j = 0; j < n; j++) {  // reading my own row while the TMA reads it too is fine
          const double2 a0 = my[3 * j + 0], a1 = my[3 * j + 1], a2 = my[3 * j + 2];
          const double gj = fma(ul.x, a0.x, fma(ul.y, a0.y, fma(ul.z, a1.x, fma(ua.x, a1.y, fma(ua.y, a2.x, ua.z * a2.y)))));
          g[j] = gj + gj;
        }
      }
    }
  }
  // the block's shared memory must stay valid until every bulk store has read it
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

}  // namespace optik

static int eval_threads(int n) { return n > 12 ? 64 : 128; }
static const void* eval_entry(int n, bool cols) {
  switch (n) {
    case 6: return cols ? (const void*)optik::eval_kernel<6, true, 128> : (const void*)optik::eval_kernel<6, false, 128>;
    case 7: return cols ? (const void*)optik::eval_kernel<7, true, 128> : (const void*)optik::eval_kernel<7, false, 128>;
    default:
      if (eval_threads(n) == 64) return cols ? (const void*)optik::eval_kernel<0, true, 64> : (const void*)optik::eval_kernel<0, false, 64>;
      return cols ? (const void*)optik::eval_kernel<0, true, 128> : (const void*)optik::eval_kernel<0, false, 128>;
  }
}
extern "C" int optik_eval_threads(int n) { return eval_threads(n); }
extern "C" int optik_eval_smem_bytes(int n, int cols) {
  const size_t tb = (size_t)eval_threads(n), warps = tb / 32;
  return (int)(sizeof(double) * (OPTIK_CHAIN_STRIDE * n + 8 + 2 + 2 * warps) +
               (cols ? 16ull * tb * (size_t)((3 * n) | 1) : 0ull) +
               sizeof(double) * 32ull * n * warps * (cols ? 1 : 2) + sizeof(double) * 4ull * n);
}
static bool eval_wants_cols(const EvalParams* p) { return p->jac_out != nullptr || p->grad_out != nullptr; }
extern "C" int optik_launch_eval(const EvalParams* p, int blocks, void* stream) {
  const bool cols = eval_wants_cols(p);
  const size_t smem = (size_t)optik_eval_smem_bytes(p->n, cols);
  const void* fn = eval_entry(p->n, cols);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  void* args[] = {(void*)p};
  e = cudaLaunchKernel(fn, dim3(blocks), dim3(eval_threads(p->n)), args, smem, (cudaStream_t)stream);
  return (int)(e != cudaSuccess ? e : cudaGetLastError());
}
// resident blocks per SM: the kernel is persistent (grid-stride over tiles), so the launcher sizes the grid to
// exactly sm_count * this
extern "C" int optik_eval_occupancy(int n, int cols, int* blocks_per_sm) {
  const int smem = optik_eval_smem_bytes(n, cols);
  const void* fn = eval_entry(n, cols != 0);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, eval_threads(n), smem);
}
