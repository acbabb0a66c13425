// diffik_kernel.cu -- batched Robot::diff_ik, one thread per configuration.
//
// Reference: crates/optik/src/lib.rs:101-239.  The reference poses, per call, the linear programme
//     max alpha   s.t.  J_W(q) v = alpha V_WE ,  -vmax <= v <= vmax ,  0 <= alpha <= 1
// (alpha bounds :134-151, velocity box :155-174, equality with the body Jacobian rotated into the world frame
// :178-197, objective :204-206) and hands it to Clarabel's interior-point solver.  Here the LP is solved EXACTLY in
// closed form, which is what makes it a batch kernel: every feasible v is alpha (u + t z) with J_W u = V and z
// spanning the null space, so  alpha* = min(1, 1 / min_t max_i |u_i + t z_i| / vmax_i).
//   n = 6: unique solution v = alpha* J_W^-1 V (the only size the reference's constraint assembly accepts, :194-195)
//   n = 7: the convex piecewise-linear max over the one free parameter t is minimised over its breakpoints
// Steps per thread: inverse-pose backward recursion (as eval_kernel) -> body Jacobian columns -> rotate into the world
// frame -> Gauss-Jordan with complete pivoting on the 6 x (n+1) augmented matrix [J_W | V] -> scale.
// The matrix lives in shared memory, interleaved by thread ([entry][thread]) so that the data-dependent row/column
// indices of the pivoting never cause bank conflicts.  The LP is always feasible (alpha = 0, v = 0) and bounded, so the
// reference returns Some((alpha, v)) at singular configurations too (lib.rs:231-238): with a rank-deficient J_W the
// elimination stops at the reduced rank; an unreachable twist direction gives alpha = 0, v = 0, a consistent system
// its basic solution scaled into the velocity box.
#include <cuda_runtime.h>

#include "ik_math.cuh"
#include "solver_params.h"

namespace optik {

constexpr int DIK_THREADS = 128;
constexpr int DIK_COLS = 8;  // n + 1 <= 8

__global__ void __launch_bounds__(DIK_THREADS) diffik_kernel(const __grid_constant__ DiffIkParams P) {
  extern __shared__ __align__(128) double smem[];
  const int n = P.n;
  double* s_chain = smem;
  uint64_t* s_bar = (uint64_t*)(smem + OPTIK_CHAIN_STRIDE * n + 8);
  double* s_M = smem + OPTIK_CHAIN_STRIDE * n + 8 + 2;  // [6 * DIK_COLS][DIK_THREADS]
  stage_chain_tma(s_chain, s_bar, P.chain, P.chain_bytes);
  const se3 tip = se3mul(load_pose8(s_chain + OPTIK_CHAIN_STRIDE * n), load_pose8(P.ee_offset));
  se3 tip_inv;
  tip_inv.q = qconj(tip.q);
  tip_inv.t = neg3(qrot(tip_inv.q, tip.t));
  double* M = s_M + threadIdx.x;
#define MAT(r, c) M[((r) * DIK_COLS + (c)) * DIK_THREADS]

  for (unsigned long long i = (unsigned long long)blockIdx.x * DIK_THREADS + threadIdx.x; i < P.B;
       i += (unsigned long long)gridDim.x * DIK_THREADS) {
    const double* q = P.x0 + i * n;
    const double* V = P.V + (P.shared_V ? 0ull : 6ull * i);
    const double* vmax = P.vmax + (P.shared_vmax ? 0ull : i * (unsigned long long)n);
    // ---- body Jacobian by the inverse-pose recursion (kinematics.rs:166-196)
    se3 Ci = tip_inv;
    for (int j = n - 1; j >= 0; j--) {
      const double* jc = s_chain + OPTIK_CHAIN_STRIDE * j;
      const int type = (int)jc[3];
      const v3 ax = mk3(jc[8], jc[9], jc[10]);
      v3 ot = mk3(jc[0], jc[1], jc[2]);
      qt oq;
      oq.x = jc[4]; oq.y = jc[5]; oq.z = jc[6]; oq.w = jc[7];
      v3 ang = qrot(Ci.q, ax), lin;
      if (type == 0) lin = cross3(Ci.t, ang);
      else { lin = ang; ang = mk3(0, 0, 0); }
      MAT(0, j) = lin.x; MAT(1, j) = lin.y; MAT(2, j) = lin.z;
      MAT(3, j) = ang.x; MAT(4, j) = ang.y; MAT(5, j) = ang.z;
      qt lq;
      if (type == 0) {
        double s, c;
        dsincos(0.5 * q[j], s, c);
        qt qa;
        qa.x = ax.x * s; qa.y = ax.y * s; qa.z = ax.z * s; qa.w = c;
        lq = qconj(qmul(oq, qa));
      } else {
        lq = qconj(oq);
        ot = add3(ot, qrot(oq, scale3(ax, q[j])));
      }
      Ci.q = qmul(Ci.q, lq);
      Ci.t = sub3(Ci.t, qrot(Ci.q, ot));
    }
    // ---- rotate the columns into the world frame (lib.rs:183-190): R_WE = conj(C_0.q); append V
    const qt rwe = qconj(Ci.q);
    double scale = 0.0;
    for (int c = 0; c < n; c++) {
      const v3 lin = qrot(rwe, mk3(MAT(0, c), MAT(1, c), MAT(2, c)));
      const v3 ang = qrot(rwe, mk3(MAT(3, c), MAT(4, c), MAT(5, c)));
      MAT(0, c) = lin.x; MAT(1, c) = lin.y; MAT(2, c) = lin.z;
      MAT(3, c) = ang.x; MAT(4, c) = ang.y; MAT(5, c) = ang.z;
      scale = fmax(scale, fmax(fmax(fabs(lin.x), fabs(lin.y)), fmax(fabs(lin.z), fmax(fabs(ang.x), fmax(fabs(ang.y), fabs(ang.z))))));
    }
    for (int r = 0; r < 6; r++) MAT(r, n) = V[r];
    // ---- Gauss-Jordan, complete pivoting (first maximum in row-major order wins)
    unsigned used = 0, pcs = 0;  // used columns (bit mask), pivot column of row k (3 bits each)
    int rank = 6;
    double vnorm = 0.0;
    for (int r = 0; r < 6; r++) vnorm = fmax(vnorm, fabs(V[r]));
    for (int k = 0; k < 6; k++) {
      int br = -1, bc = -1;
      double best = 0.0;
      for (int r = k; r < 6; r++)
        for (int c = 0; c < n; c++) {
          const double a = fabs(MAT(r, c));
          if (!((used >> c) & 1u) && a > best) { best = a; br = r; bc = c; }
        }
      if (br < 0 || best <= 1e-12 * scale) { rank = k; break; }
      if (br != k)
        for (int j = 0; j <= n; j++) { const double t = MAT(k, j); MAT(k, j) = MAT(br, j); MAT(br, j) = t; }
      used |= 1u << bc;
      pcs |= (unsigned)bc << (3 * k);
      const double inv = 1.0 / MAT(k, bc);
      for (int j = 0; j <= n; j++) MAT(k, j) = MAT(k, j) * inv;
      for (int r = 0; r < 6; r++) {
        if (r == k) continue;
        const double f = MAT(r, bc);
        for (int j = 0; j <= n; j++) MAT(r, j) = fma(-f, MAT(k, j), MAT(r, j));
      }
    }
    if (rank < 6) {
      double resid = 0.0;
      for (int r = rank; r < 6; r++) resid = fmax(resid, fabs(MAT(r, n)));
      if (resid > 1e-9 * vnorm) {  // V is not in the range of J_W: the optimum is alpha = 0
        P.status_out[i] = 1;
        P.alpha_out[i] = 0.0;
        for (int j = 0; j < n; j++) P.v_out[i * n + j] = 0.0;
        continue;
      }
    }
    // ---- v(t) = a + t b ; rows 0/1 of the (now free) matrix storage hold a_i / vmax_i and b_i / vmax_i per joint
    int fcol = -1;
    if (rank == 6) for (int c = 0; c < n; c++) if (!((used >> c) & 1u)) fcol = c;
    double av[6], bv[6];
    for (int k = 0; k < 6; k++) { av[k] = k < rank ? MAT(k, n) : 0.0; bv[k] = (fcol >= 0) ? -MAT(k, fcol) : 0.0; }
#define CS(j) M[(0 * DIK_COLS + (j)) * DIK_THREADS]
#define SS(j) M[(1 * DIK_COLS + (j)) * DIK_THREADS]
#define AA(j) M[(2 * DIK_COLS + (j)) * DIK_THREADS]
#define BB(j) M[(3 * DIK_COLS + (j)) * DIK_THREADS]
    for (int j = 0; j < n; j++) { AA(j) = 0.0; BB(j) = 0.0; }  // joints without a pivot stay at rest
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int c = (pcs >> (3 * k)) & 7;
      if (k < rank) { AA(c) = av[k]; BB(c) = bv[k]; }
    }
    if (fcol >= 0) { AA(fcol) = 0.0; BB(fcol) = 1.0; }
    double gbest = 0.0, tbest = 0.0;
    for (int j = 0; j < n; j++) {
      const double m = vmax[j];
      CS(j) = AA(j) / m; SS(j) = BB(j) / m;
      gbest = fmax(gbest, fabs(CS(j)));
    }
    if (fcol >= 0) {
      for (int p = 0; p < 2 * n; p++)
        for (int qq = p + 1; qq < 2 * n; qq++) {
          const double cp = (p & 1) ? -CS(p >> 1) : CS(p >> 1), sp = (p & 1) ? -SS(p >> 1) : SS(p >> 1);
          const double cq = (qq & 1) ? -CS(qq >> 1) : CS(qq >> 1), sq = (qq & 1) ? -SS(qq >> 1) : SS(qq >> 1);
          if (sp == sq) continue;
          const double t = (cq - cp) / (sp - sq);
          double g = 0.0;
          for (int j = 0; j < n; j++) g = fmax(g, fabs(fma(t, SS(j), CS(j))));
          if (g < gbest) { gbest = g; tbest = t; }
        }
    }
    const double alpha = gbest > 1.0 ? 1.0 / gbest : 1.0;
    P.status_out[i] = 1;
    P.alpha_out[i] = alpha;
    for (int j = 0; j < n; j++) P.v_out[i * n + j] = alpha * fma(tbest, BB(j), AA(j));
  }
#undef MAT
#undef CS
#undef SS
#undef AA
#undef BB
}

}  // namespace optik

extern "C" int optik_diffik_smem_bytes(int n) {
  return (int)(sizeof(double) * (OPTIK_CHAIN_STRIDE * n + 8 + 2) + sizeof(double) * 6 * optik::DIK_COLS * optik::DIK_THREADS);
}
extern "C" int optik_launch_diffik(const DiffIkParams* p, int blocks, void* stream) {
  const int smem = optik_diffik_smem_bytes(p->n);
  cudaError_t e = cudaFuncSetAttribute(optik::diffik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  optik::diffik_kernel<<<blocks, optik::DIK_THREADS, smem, (cudaStream_t)stream>>>(*p);
  return (int)cudaGetLastError();
}
