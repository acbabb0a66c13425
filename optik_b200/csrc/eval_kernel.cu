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
// independent of n.  Columns are staged in shared memory (each thread owns a 6n-double row, padded to avoid bank
// conflicts), the gradient is formed from them, and the block's rows are written back with coalesced 16-byte stores.
// The chain is staged once per block with a 1-D TMA bulk copy.
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

__global__ void __launch_bounds__(EVAL_THREADS) eval_kernel(const __grid_constant__ EvalParams P) {
  extern __shared__ __align__(128) double smem[];
  // layout: chain blob | mbarrier | per-thread Jacobian rows (row stride 6n+1 doubles)
  double* s_chain = smem;
  const int chain_doubles = OPTIK_CHAIN_STRIDE * P.n + 8;
  uint64_t* s_bar = (uint64_t*)(smem + chain_doubles);
  double* s_jac = smem + chain_doubles + 2;
  const int n = P.n;
  const int row = 6 * n + 1;

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
  double* my = s_jac + threadIdx.x * row;

  for (unsigned long long base = (unsigned long long)blockIdx.x * EVAL_THREADS; base < P.B;
       base += (unsigned long long)gridDim.x * EVAL_THREADS) {
    const unsigned long long i = base + threadIdx.x;
    const bool live = i < P.B;
    se3 B = tip;
    if (live) {
      const double* q = P.q + i * n;
      for (int j = n - 1; j >= 0; j--) {
        const double* jc = s_chain + OPTIK_CHAIN_STRIDE * j;
        const int type = (int)jc[3];
        const v3 ax = mk3(jc[8], jc[9], jc[10]);
        // column of joint j from B_j (pose of the EE in frame j)
        const v3 lin = qrot_inv(B.q, (type == 0) ? cross3(ax, B.t) : ax);
        v3 ang = qrot_inv(B.q, ax);
        if (type != 0) ang = mk3(0, 0, 0);
        my[6 * j + 0] = lin.x; my[6 * j + 1] = lin.y; my[6 * j + 2] = lin.z;
        my[6 * j + 3] = ang.x; my[6 * j + 4] = ang.y; my[6 * j + 5] = ang.z;
        // B_{j-1} = origin_j * motion_j(q_j) * B_j
        se3 L;
        qt oq;
        oq.x = jc[4]; oq.y = jc[5]; oq.z = jc[6]; oq.w = jc[7];
        const v3 ot = mk3(jc[0], jc[1], jc[2]);
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
          // u_lin = J^T gl ; u_ang = Q^T gl + J^T ga = J^T (C^T gl) + J^T ga
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
          for (int j = 0; j < n; j++) {
            const double* c = my + 6 * j;
            const double gj = fma(ul.x, c[0], fma(ul.y, c[1], fma(ul.z, c[2], fma(ua.x, c[3], fma(ua.y, c[4], ua.z * c[5])))));
            g[j] = gj + gj;
          }
        }
      }
    }
    if (P.jac_out) {
      // block-cooperative coalesced write-back of the staged rows: the block's rows are contiguous in HBM
      __syncthreads();
      const unsigned long long rows = (P.B - base < (unsigned long long)EVAL_THREADS) ? (P.B - base) : EVAL_THREADS;
      const unsigned long long total = rows * 6ull * n;
      double* out = P.jac_out + base * 6ull * n;
      for (unsigned long long e = threadIdx.x; e < total; e += EVAL_THREADS) {
        const unsigned int rr = (unsigned int)(e / (6u * n)), cc = (unsigned int)(e % (6u * n));
        out[e] = s_jac[rr * row + cc];
      }
      __syncthreads();
    }
  }
}

}  // namespace optik

extern "C" int optik_launch_eval(const EvalParams* p, int blocks, void* stream) {
  const size_t smem = sizeof(double) * (OPTIK_CHAIN_STRIDE * p->n + 8 + 2 + (size_t)optik::EVAL_THREADS * (6 * p->n + 1));
  cudaError_t e = cudaFuncSetAttribute(optik::eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  optik::eval_kernel<<<blocks, optik::EVAL_THREADS, smem, (cudaStream_t)stream>>>(*p);
  return (int)cudaGetLastError();
}
extern "C" int optik_eval_smem_bytes(int n) {
  return (int)(sizeof(double) * (OPTIK_CHAIN_STRIDE * n + 8 + 2 + (size_t)optik::EVAL_THREADS * (6 * n + 1)));
}
