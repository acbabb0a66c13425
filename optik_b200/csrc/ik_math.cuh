// ik_math.cuh -- pose-error and task-Jacobian math shared by the solve kernels (tile and thread-per-seed layouts).
// Restates crates/optik/src/math.rs:40-203 and objective.rs:7-110 (see DESIGN.md "Deviations" for the half-angle
// identities and the Taylor limits); operation order is part of the arithmetic spec mirrored by oracle/solver_twin.c.
#pragma once
#include "dmath.cuh"

namespace optik {

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

struct ErrCoef {
  v3 w, xt, cv;
  double ce, da;
};

// X = (xq, xt) = T_tgt^-1 T_ee  ->  w = so3::log (math.rs:40-63), elin = V^-1 t (math.rs:107-124) and the scalars of
// Jlog6 = [[J, Q],[0, J]] (math.rs:72-94, 135-203)
DEV void error_terms(qt xq, v3 xt, ErrCoef& c, v3& elin) {
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
  elin = axpy3(ce, cross3(w, wxt), axpy3(-0.5, wxt, xt));  // V^-1 t
  const double d = dot3(w, xt);                            // Q = C*J scalars (math.rs:160-169)
  const double kc = fma(th2, bq, ce + ce);
  c.w = w; c.xt = xt; c.ce = ce;
  c.cv = axpy3(bq * d, w, scale3(xt, -kc));
  c.da = d * ce;
}

// task column = Jlog6 * [lin; ang] (objective.rs:79-81):  top = J lin + C (J ang), bot = J ang
DEV void task_col(const ErrCoef& c, v3 lin, v3 ang, v3& top, v3& bot) {
  const v3 w = c.w;
  const v3 wxa = cross3(w, ang);
  const v3 ja = axpy3(c.ce, cross3(w, wxa), axpy3(0.5, wxa, ang));
  const v3 wxl = cross3(w, lin);
  const v3 jl = axpy3(c.ce, cross3(w, wxl), axpy3(0.5, wxl, lin));
  const double wu = dot3(w, ja), tu = dot3(c.xt, ja);
  const v3 cu = axpy3(c.da, ja, axpy3(c.ce * tu, w, axpy3(wu, c.cv, scale3(cross3(c.xt, ja), 0.5))));
  top = add3(jl, cu);
  bot = ja;
}

// Thread-per-seed layout: Jlog6 = [[J, C J],[0, J]] as two explicit 3x3 matrices (row-major), built once per accepted
// point, so that a column costs three matrix-vector products (27 fma) instead of the cross-product form (64):
//   J = (1 - ce |w|^2) I + 1/2 [w]x + ce w w^T ,  C = da I + 1/2 [t]x + cv w^T + ce w t^T   (task_col's operators)
DEV void task_mats(const ErrCoef& c, double* J, double* CJ) {
  const v3 w = c.w, t = c.xt, cv = c.cv;
  const double ce = c.ce;
  const double a = fma(-ce, dot3(w, w), 1.0);
  const double cxy = ce * (w.x * w.y), cxz = ce * (w.x * w.z), cyz = ce * (w.y * w.z);
  J[0] = fma(ce, w.x * w.x, a); J[1] = fma(-0.5, w.z, cxy);   J[2] = fma(0.5, w.y, cxz);
  J[3] = fma(0.5, w.z, cxy);    J[4] = fma(ce, w.y * w.y, a); J[5] = fma(-0.5, w.x, cyz);
  J[6] = fma(-0.5, w.y, cxz);   J[7] = fma(0.5, w.x, cyz);    J[8] = fma(ce, w.z * w.z, a);
  const v3 cw = scale3(w, ce);
  double C[9];
  C[0] = fma(cv.x, w.x, cw.x * t.x) + c.da;          C[1] = fma(cv.x, w.y, fma(cw.x, t.y, -0.5 * t.z)); C[2] = fma(cv.x, w.z, fma(cw.x, t.z, 0.5 * t.y));
  C[3] = fma(cv.y, w.x, fma(cw.y, t.x, 0.5 * t.z));  C[4] = fma(cv.y, w.y, cw.y * t.y) + c.da;          C[5] = fma(cv.y, w.z, fma(cw.y, t.z, -0.5 * t.x));
  C[6] = fma(cv.z, w.x, fma(cw.z, t.x, -0.5 * t.y)); C[7] = fma(cv.z, w.y, fma(cw.z, t.y, 0.5 * t.x));  C[8] = fma(cv.z, w.z, cw.z * t.z) + c.da;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) CJ[3 * i + k] = fma(C[3 * i], J[k], fma(C[3 * i + 1], J[3 + k], C[3 * i + 2] * J[6 + k]));
}
DEV void task_col_m(const double* J, const double* CJ, v3 lin, v3 ang, v3& top, v3& bot) {
  double tp[3], bt[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    bt[i] = fma(J[3 * i], ang.x, fma(J[3 * i + 1], ang.y, J[3 * i + 2] * ang.z));
    tp[i] = fma(J[3 * i], lin.x, fma(J[3 * i + 1], lin.y, fma(J[3 * i + 2], lin.z,
            fma(CJ[3 * i], ang.x, fma(CJ[3 * i + 1], ang.y, CJ[3 * i + 2] * ang.z)))));
  }
  top = mk3(tp[0], tp[1], tp[2]);
  bot = mk3(bt[0], bt[1], bt[2]);
}

DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage the chain blob into shared memory with one 1-D TMA bulk copy (thread 0 issues, everyone waits on the mbarrier).
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

DEV unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

}  // namespace optik
