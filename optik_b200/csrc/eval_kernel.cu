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
// Algorithm: a single BACKWARD recursion  B_{j-1} = L_j * B_j  (B_n = tip, L_j = origin_j * motion_j(q_j)) carries
// the pose of the end effector in joint j's frame; the body-frame Jacobian column of joint j is then
//   [ R_Bj^T (axis_j x p_Bj) ; R_Bj^T axis_j ]
// and B_0 is the end-effector pose -- no per-joint transforms are stored, so the register footprint is
// independent of n.  Every thread stages its 6n-double Jacobian row in shared memory with 128-bit stores (row stride
// = an odd number of 16-byte units, so quarter-warps never bank-conflict), forms the gradient from it, and then hands
// the row to the TMA: one `cp.async.bulk.global.shared::cta` per thread writes the 48n contiguous bytes to HBM while
// the thread already works on its next configuration (it only waits for the bulk read of its own row before
// overwriting it).  No block barrier in the loop.  The chain is staged once per block with a 1-D TMA bulk load.
#include <cuda_runtime.h>

#include "dmath.cuh"
#include "solver_params.h"

namespace optik {

DEV uint32_t e_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int EVAL_THREADS = 128;

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

__global__ void __launch_bounds__(EVAL_THREADS) eval_kernel(const __grid_constant__ EvalParams P) {
  extern __shared__ __align__(128) double smem[];
  // layout: chain blob | mbarrier (16 B) | per-thread Jacobian rows
  double* s_chain = smem;
  const int n = P.n;
  const int chain_doubles = OPTIK_CHAIN_STRIDE * n + 8;
  uint64_t* s_bar = (uint64_t*)(smem + chain_doubles);
  double2* s_rows = (double2*)(smem + chain_doubles + 2);
  const int units = eval_row_units(n);

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(e_smem_u32(s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(e_smem_u32(s_bar)), "r"(P.chain_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     e_smem_u32(s_chain)),
                 "l"(P.chain), "r"(P.chain_bytes), "r"(e_smem_u32(s_bar))
                 : "memory");
  }
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(e_smem_u32(s_bar)), "r"(0u)
        : "memory");
  } while (!ok);

  const se3 tip = se3mul(e_load_pose8(s_chain + OPTIK_CHAIN_STRIDE * n), e_load_pose8(P.ee_offset));
  const bool want_obj = (P.f_out != nullptr) || (P.grad_out != nullptr);
  const bool want_cols = (P.jac_out != nullptr) || (P.grad_out != nullptr);
  double2* my = s_rows + (size_t)threadIdx.x * units;
  const uint32_t my_addr = e_smem_u32(my);
  const uint32_t row_bytes = 48u * (uint32_t)n;

  for (unsigned long long i = (unsigned long long)blockIdx.x * EVAL_THREADS + threadIdx.x; i < P.B;
       i += (unsigned long long)gridDim.x * EVAL_THREADS) {
    // my previous row must have been read out by the TMA before it is overwritten
    if (P.jac_out) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    se3 B = tip;
    const double* q = P.q + i * n;
    for (int j = n - 1; j >= 0; j--) {
      const double2* jc = (const double2*)(s_chain + OPTIK_CHAIN_STRIDE * j);
      const double2 c0 = jc[0], c1 = jc[1], c2 = jc[2], c3 = jc[3], c4 = jc[4], c5 = jc[5];
      const v3 ot = mk3(c0.x, c0.y, c1.x);
      const int type = (int)c1.y;
      qt oq;
      oq.x = c2.x; oq.y = c2.y; oq.z = c3.x; oq.w = c3.y;
      const v3 ax = mk3(c4.x, c4.y, c5.x);
      if (want_cols) {  // column of joint j from B_j (pose of the EE in frame j)
        const v3 lin = qrot_inv(B.q, (type == 0) ? cross3(ax, B.t) : ax);
        v3 ang = qrot_inv(B.q, ax);
        if (type != 0) ang = mk3(0, 0, 0);
        my[3 * j + 0] = make_double2(lin.x, lin.y);
        my[3 * j + 1] = make_double2(lin.z, ang.x);
        my[3 * j + 2] = make_double2(ang.y, ang.z);
      }
      // B_{j-1} = origin_j * motion_j(q_j) * B_j
      se3 L;
      const double qj = q[j];
      if (type == 0) {
        double s, c;
        dsincos(0.5 * qj, s, c);
        qt qa;
        qa.x = ax.x * s; qa.y = ax.y * s; qa.z = ax.z * s; qa.w = c;
        L.q = qmul(oq, qa);
        L.t = ot;
      } else {
        L.q = oq;
        L.t = add3(ot, qrot(oq, scale3(ax, qj)));
      }
      B = se3mul(L, B);
    }
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
      const se3 tgt = e_load_pose8(P.targets + (unsigned long long)P.target_stride * i);
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

extern "C" int optik_eval_smem_bytes(int n);
extern "C" int optik_launch_eval(const EvalParams* p, int blocks, void* stream) {
  const size_t smem = (size_t)optik_eval_smem_bytes(p->n);
  cudaError_t e = cudaFuncSetAttribute(optik::eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  optik::eval_kernel<<<blocks, optik::EVAL_THREADS, smem, (cudaStream_t)stream>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int optik_eval_smem_bytes(int n) {
  return (int)(sizeof(double) * (OPTIK_CHAIN_STRIDE * n + 8 + 2) + 16ull * optik::EVAL_THREADS * (size_t)((3 * n) | 1));
}
