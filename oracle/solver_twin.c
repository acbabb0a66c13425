/*
 * solver_twin.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * fp64 twin of the CUDA solve kernel (optik_b200/csrc/solve_kernel.cu): the
 * same algorithm, the same seeds, the same sequence of IEEE-754 operations
 * (see twin_math.h), restated in plain C with the tile's lanes written as
 * loops.  It checks GPU<->CPU parity per restart seed (status, iteration
 * count, solution), which is where "identical success set under fixed RNG"
 * is checkable.
 *
 * What it stands in for in the reference: the per-restart NLopt SLSQP solve
 * and its stop rules, crates/optik/src/lib.rs:302-391 (stopval = tol_f :345,
 * ftol_abs = tol_df_eff :283-293,:346, xtol_abs = tol_dx :347, bounds
 * :348-349, success classification :376-379).  The inner optimiser is
 * replaced BY DESIGN (north star) with a bound-projected Levenberg-Marquardt
 * iteration in dual (6x6) form, so per-seed trajectories are "parity
 * unpinned" against NLopt; the evaluator underneath is checked against
 * optik_oracle.c (which is pinned to the reference's golden vectors).
 */
#include <stdlib.h>
#include <string.h>

#include "twin_math.h"

#define CHAIN_STRIDE 16
#define MAX_DOF 32

enum { /* attempt status; mirrored by optik_b200/csrc/solve_kernel.cu */
  ST_NONE = 0,
  ST_STOPVAL = 1,   /* f < tol_f                 (NLopt StopValReached) */
  ST_FTOL = 2,      /* accepted |df| < tol_df    (NLopt FtolReached)    */
  ST_XTOL = 3,      /* accepted max|dx| < tol_dx (NLopt XtolReached)    */
  ST_ITERCAP = 4,   /* evaluation cap (no reference analogue)          */
  ST_STUCK = 5,     /* damping exceeded lambda_max: no descent step     */
  ST_NAN = 6,
  ST_SKIPPED = 7    /* not run: timeout / Speed-mode early exit         */
};

typedef struct twin_params {
  double tol_f, tol_df_eff, tol_df_user, tol_dx;
  double wl[3], wa[3];
  int weighted, max_evals;
  double lambda0, lambda_dec, lambda_inc, lambda_min, lambda_max;
  double stall_rel; /* accepted step with df < stall_rel*f counts as "slow" */
  int stall_count;  /* this many consecutive slow steps => ST_STUCK */
  int layout;       /* 0: tile (lane per joint) kernel, 1: thread-per-seed kernel -- different summation orders */
} twin_params;

typedef struct twin_chain {
  int n;
  int type[MAX_DOF];
  v3 ot[MAX_DOF], ax[MAX_DOF];
  qt oq[MAX_DOF];
  double lb[MAX_DOF], ub[MAX_DOF], slb[MAX_DOF], sub[MAX_DOF]; /* s*: seed sampling range */
  se3t tip;
} twin_chain;

typedef struct {
  double f;
  double r[6];
  double Jr[MAX_DOF][6];
  se3t ee;
} twin_eval;

size_t twin_chain_sizeof(void) { return sizeof(twin_chain); }
int twin_chain_n(const twin_chain* c) { return c->n; }
const double* twin_chain_slb(const twin_chain* c) { return c->slb; }
const double* twin_chain_sub(const twin_chain* c) { return c->sub; }

static se3t se3_from_pose8(const double* p) {
  se3t r;
  r.q.x = p[0]; r.q.y = p[1]; r.q.z = p[2]; r.q.w = p[3];
  r.t = v3_make(p[4], p[5], p[6]);
  return r;
}

/* flat chain (articulated joints followed by at most one fixed tip joint) -> twin_chain */
int twin_chain_init(twin_chain* c, const double* chain, int njoints, const double* ee_offset) {
  const double PI = 3.14159265358979311600e+00;
  static const double ID[8] = {0, 0, 0, 1, 0, 0, 0, 0};
  int n = 0;
  se3t fixed = se3_from_pose8(ID);
  for (int i = 0; i < njoints; i++) {
    const double* j = chain + CHAIN_STRIDE * i;
    int type = (int)j[3];
    if (type == 2) {
      if (i != njoints - 1) return -1;
      double p[8] = {j[4], j[5], j[6], j[7], j[0], j[1], j[2], 0};
      fixed = se3_from_pose8(p);
      continue;
    }
    if (n >= MAX_DOF) return -2;
    c->type[n] = type;
    c->ot[n] = v3_make(j[0], j[1], j[2]);
    c->oq[n].x = j[4]; c->oq[n].y = j[5]; c->oq[n].z = j[6]; c->oq[n].w = j[7];
    c->ax[n] = v3_make(j[8], j[9], j[10]);
    c->lb[n] = j[12]; c->ub[n] = j[13];
    int finite = isfinite(j[12]) && isfinite(j[13]);
    c->slb[n] = finite ? j[12] : -PI;
    c->sub[n] = finite ? j[13] : PI;
    n++;
  }
  c->n = n;
  c->tip = se3_mul(fixed, se3_from_pose8(ee_offset ? ee_offset : ID));
  return 0;
}

static inline double dot6(const double* a, const double* b) {
  return fma(a[0], b[0], fma(a[1], b[1], fma(a[2], b[2], fma(a[3], b[3], fma(a[4], b[4], a[5] * b[5])))));
}
static inline v3 weight3(qt tq, const double w[3], v3 u) {
  v3 a = qt_rot(tq, u);
  a = v3_make(a.x * w[0], a.y * w[1], a.z * w[2]);
  return qt_rot_inv(tq, a);
}

/* Pose-error terms shared by both kernel layouts: X = (xq, xt) -> w = so3 log, elin = V^-1 t, and the scalars of
 * Jlog6 (math.rs:40-203 restated with half-angle identities; Taylor below theta^2 <= 1e-6). */
typedef struct { v3 w, xt, cv; double ce, da; } err_coef;
static void error_terms(qt xq, v3 xt, err_coef* c, v3* elin) {
  double qw = xq.w;
  v3 v = v3_make(xq.x, xq.y, xq.z);
  if (!(qw >= 0.0)) { qw = -qw; v = v3_neg(v); }
  double vn2 = v3_dot(v, v), k, th2, ce, bq;
  if (vn2 > 1e-6) {
    double vn = sqrt(vn2);
    double half = tw_atan2_pos(vn, qw);
    double inv_vn = 1.0 / vn, inv_half = 1.0 / half;
    k = half * inv_vn;                         /* atan2(|v|,w)/|v| */
    double p = k * qw;                         /* (theta/2)/tan(theta/2) */
    double it2 = 0.25 * (inv_half * inv_half); /* 1/theta^2 */
    th2 = 4.0 * (half * half);
    ce = (1.0 - p) * it2;                      /* coefficient of hat(w)^2 in V^-1, J and a_q of Q */
    double a = (vn * qw) * inv_half;           /* sin(theta)/theta */
    bq = fma((1.0 + a) * it2, 0.25 * (inv_vn * inv_vn), -2.0 * (it2 * it2));
  } else {
    double iw = 1.0 / qw, iw2 = iw * iw;
    k = iw * fma(vn2 * iw2, fma(vn2 * iw2, 0.2, -1.0 / 3.0), 1.0);
    th2 = 4.0 * ((k * k) * vn2);
    ce = fma(th2, fma(th2, 1.0 / 30240.0, 1.0 / 720.0), 1.0 / 12.0);
    bq = fma(th2, 1.0 / 7560.0, 1.0 / 360.0);
  }
  v3 w = v3_scale(v, k + k);
  /* e = [V^-1 t ; w],  V^-1 u = u - 1/2 w x u + ce w x (w x u) */
  v3 wxt = v3_cross(w, xt);
  *elin = v3_axpy(ce, v3_cross(w, wxt), v3_axpy(-0.5, wxt, xt));
  /* Q = C*J scalars */
  double d = v3_dot(w, xt);
  double kc = fma(th2, bq, ce + ce);
  c->w = w; c->xt = xt; c->ce = ce;
  c->cv = v3_axpy(bq * d, w, v3_scale(xt, -kc)); /* b d w - (theta^2 b + 2a) t */
  c->da = d * ce;
}
/* task column = Jlog6 * [lin; ang]:  top = J lin + C (J ang), bot = J ang */
static void task_col(const err_coef* c, v3 lin, v3 ang, v3* top, v3* bot) {
  v3 w = c->w;
  /* J u = u + 1/2 w x u + ce w x (w x u) */
  v3 wxa = v3_cross(w, ang);
  v3 ja = v3_axpy(c->ce, v3_cross(w, wxa), v3_axpy(0.5, wxa, ang));
  v3 wxl = v3_cross(w, lin);
  v3 jl = v3_axpy(c->ce, v3_cross(w, wxl), v3_axpy(0.5, wxl, lin));
  /* C u = 1/2 t x u + cv (w.u) + ce w (t.u) + d ce u */
  double wu = v3_dot(w, ja), tu = v3_dot(c->xt, ja);
  v3 cu = v3_axpy(c->da, ja, v3_axpy(c->ce * tu, w, v3_axpy(wu, c->cv, v3_scale(v3_cross(c->xt, ja), 0.5))));
  *top = v3_add(jl, cu);
  *bot = ja;
}

/* Thread-per-seed layout: Jlog6 = [[J, C J],[0, J]] as two explicit 3x3 matrices (row-major), built once per accepted
 * point, so that a column costs three matrix-vector products (27 fma) instead of the cross-product form above (64).
 *   J  = (1 - ce |w|^2) I + 1/2 [w]x + ce w w^T
 *   C  = da I + 1/2 [t]x + cv w^T + ce w t^T         (same operator as task_col's C u) */
static void task_mats(const err_coef* c, double J[9], double CJ[9]) {
  v3 w = c->w, t = c->xt, cv = c->cv;
  double ce = c->ce;
  double a = fma(-ce, v3_dot(w, w), 1.0);
  double cxy = ce * (w.x * w.y), cxz = ce * (w.x * w.z), cyz = ce * (w.y * w.z);
  J[0] = fma(ce, w.x * w.x, a); J[1] = fma(-0.5, w.z, cxy);   J[2] = fma(0.5, w.y, cxz);
  J[3] = fma(0.5, w.z, cxy);    J[4] = fma(ce, w.y * w.y, a); J[5] = fma(-0.5, w.x, cyz);
  J[6] = fma(-0.5, w.y, cxz);   J[7] = fma(0.5, w.x, cyz);    J[8] = fma(ce, w.z * w.z, a);
  v3 cw = v3_scale(w, ce);
  double C[9];
  C[0] = fma(cv.x, w.x, cw.x * t.x) + c->da;        C[1] = fma(cv.x, w.y, fma(cw.x, t.y, -0.5 * t.z)); C[2] = fma(cv.x, w.z, fma(cw.x, t.z, 0.5 * t.y));
  C[3] = fma(cv.y, w.x, fma(cw.y, t.x, 0.5 * t.z)); C[4] = fma(cv.y, w.y, cw.y * t.y) + c->da;         C[5] = fma(cv.y, w.z, fma(cw.y, t.z, -0.5 * t.x));
  C[6] = fma(cv.z, w.x, fma(cw.z, t.x, -0.5 * t.y)); C[7] = fma(cv.z, w.y, fma(cw.z, t.y, 0.5 * t.x)); C[8] = fma(cv.z, w.z, cw.z * t.z) + c->da;
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) CJ[3 * i + k] = fma(C[3 * i], J[k], fma(C[3 * i + 1], J[3 + k], C[3 * i + 2] * J[6 + k]));
}
static void task_col_m(const double J[9], const double CJ[9], v3 lin, v3 ang, v3* top, v3* bot) {
  double tp[3], bt[3];
  for (int i = 0; i < 3; i++) {
    bt[i] = fma(J[3 * i], ang.x, fma(J[3 * i + 1], ang.y, J[3 * i + 2] * ang.z));
    tp[i] = fma(J[3 * i], lin.x, fma(J[3 * i + 1], lin.y, fma(J[3 * i + 2], lin.z,
            fma(CJ[3 * i], ang.x, fma(CJ[3 * i + 1], ang.y, CJ[3 * i + 2] * ang.z)))));
  }
  *top = v3_make(tp[0], tp[1], tp[2]);
  *bot = v3_make(bt[0], bt[1], bt[2]);
}

/* tip handling rule (depends on n only, so every tile width gives the same bits): the fixed tip transform rides
 * the scan as entry n unless n is exactly a tile width (8, 16, 32), where no spare lane exists. */
static int tip_in_scan(int n) { return !(n == 8 || n == 16 || n == 32); }

/* One evaluation: FK scan, body Jacobian, se(3) log error, weighted residual r,
 * weighted task Jacobian Jr = W * Jlog6(X) * J_body.
 * (kinematics.rs:123-196, math.rs:40-203, objective.rs:7-110 -- restated.)
 * FK is evaluated IN THE TARGET'S FRAME: joint 0's origin is pre-multiplied by T_tgt^-1 (tgt_inv), so the scan
 * yields X = T_tgt^-1 * T_ee directly; the body-frame Jacobian is invariant to the choice of world frame. */
void twin_evaluate(const twin_chain* c, const twin_params* P, se3t tgt, const double* q, twin_eval* E) {
  int n = c->n;
  int fold = tip_in_scan(n);
  int m = fold ? n + 1 : n;
  se3t T[MAX_DOF + 1], Tn[MAX_DOF + 1];
  se3t tgt_inv;
  tgt_inv.q = qt_conj(tgt.q);
  tgt_inv.t = v3_neg(qt_rot(tgt_inv.q, tgt.t));
  for (int j = 0; j < n; j++) { /* lane j: local transform origin_j * motion_j(q_j) */
    se3t O;
    O.q = c->oq[j]; O.t = c->ot[j];
    if (j == 0) O = se3_mul(tgt_inv, O);
    if (c->type[j] == 0) {
      double s, cs;
      tw_sincos(0.5 * q[j], &s, &cs);
      qt qa = {c->ax[j].x * s, c->ax[j].y * s, c->ax[j].z * s, cs};
      T[j].q = qt_mul(O.q, qa);
      T[j].t = O.t;
    } else {
      T[j].q = O.q;
      T[j].t = v3_add(O.t, qt_rot(O.q, v3_scale(c->ax[j], q[j])));
    }
  }
  if (fold) T[n] = c->tip;
  for (int d = 1; d < m; d <<= 1) { /* Kogge-Stone inclusive scan over lanes */
    for (int j = 0; j < m; j++) Tn[j] = (j >= d) ? se3_mul(T[j - d], T[j]) : T[j];
    memcpy(T, Tn, sizeof(se3t) * m);
  }
  se3t ee = fold ? T[n] : se3_mul(T[n - 1], c->tip);
  E->ee = ee; /* == X = T_tgt^-1 * T_ee (objective.rs:48-49) */
  qt xq = ee.q;
  v3 xt = ee.t;
  err_coef ec;
  v3 elin;
  error_terms(xq, xt, &ec, &elin);
  v3 rl = elin, ra = ec.w;
  if (P->weighted) { rl = weight3(tgt.q, P->wl, elin); ra = weight3(tgt.q, P->wa, ec.w); }
  E->r[0] = rl.x; E->r[1] = rl.y; E->r[2] = rl.z; E->r[3] = ra.x; E->r[4] = ra.y; E->r[5] = ra.z;
  E->f = dot6(E->r, E->r);
  for (int j = 0; j < n; j++) { /* lane j: body Jacobian column -> task column */
    v3 lin, ang;
    v3 axw = qt_rot(T[j].q, c->ax[j]);
    if (c->type[j] == 0) {
      v3 lw = v3_cross(axw, v3_sub(ee.t, T[j].t));
      lin = qt_rot_inv(ee.q, lw);
      ang = qt_rot_inv(ee.q, axw);
    } else {
      lin = qt_rot_inv(ee.q, axw);
      ang = v3_make(0, 0, 0);
    }
    v3 top, bot;
    task_col(&ec, lin, ang, &top, &bot);
    if (P->weighted) { top = weight3(tgt.q, P->wl, top); bot = weight3(tgt.q, P->wa, bot); }
    E->Jr[j][0] = top.x; E->Jr[j][1] = top.y; E->Jr[j][2] = top.z;
    E->Jr[j][3] = bot.x; E->Jr[j][4] = bot.y; E->Jr[j][5] = bot.z;
  }
}

/* xor-butterfly all-reduce over the tile (lanes >= n hold 0): pairwise tree, low strides first */
static double tree_sum(const double* x, int n) {
  double b[MAX_DOF];
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = 0; i < m; i++) b[i] = i < n ? x[i] : 0.0;
  for (int s = 1; s < m; s <<= 1)
    for (int i = 0; i < m; i += 2 * s) b[i] = b[i] + b[i + s];
  return b[0];
}

/* LDL^T solve of the SPD 6x6 system A y = r (A symmetric, lower triangle used) */
static void ldl6_solve(double A[6][6], const double* r, double* y) {
  double L[6][6], D[6], inv[6];
  for (int j = 0; j < 6; j++) {
    double dj = A[j][j];
    for (int k = 0; k < j; k++) dj = fma(-(L[j][k] * L[j][k]), D[k], dj);
    D[j] = dj;
    inv[j] = 1.0 / dj;
    for (int i = j + 1; i < 6; i++) {
      double s = A[i][j];
      for (int k = 0; k < j; k++) s = fma(-(L[i][k] * L[j][k]), D[k], s);
      L[i][j] = s * inv[j];
    }
  }
  double z[6];
  for (int i = 0; i < 6; i++) {
    double s = r[i];
    for (int k = 0; k < i; k++) s = fma(-L[i][k], z[k], s);
    z[i] = s;
  }
  for (int i = 5; i >= 0; i--) {
    double s = z[i] * inv[i];
    for (int k = i + 1; k < 6; k++) s = fma(-L[k][i], y[k], s);
    y[i] = s;
  }
}

double* twin_trace = 0; /* debug: if set, receives (f_trial, lambda, accept) per evaluation */
int twin_trace_cap = 0;
void twin_set_trace(double* buf, int cap) { twin_trace = buf; twin_trace_cap = cap; }

/* One restart attempt from q_init. Returns status; q_out/f_out = last trial point on exit. */
static int twin_attempt_t1(const twin_chain* c, const twin_params* P, se3t tgt, const double* q_init, double* q_out,
                           double* f_out, int* evals_out);
int twin_attempt(const twin_chain* c, const twin_params* P, se3t tgt, const double* q_init, double* q_out,
                 double* f_out, int* evals_out) {
  if (P->layout == 1) return twin_attempt_t1(c, P, tgt, q_init, q_out, f_out, evals_out);
  int n = c->n, evals = 0, have_cur = 0, status = ST_NONE, slow = 0;
  double qc[MAX_DOF], qt_[MAX_DOF], lambda = P->lambda0;
  twin_eval Ec, Et;
  Ec.f = 0;
  for (int j = 0; j < n; j++) qt_[j] = fmin(fmax(q_init[j], c->lb[j]), c->ub[j]);
  for (;;) {
    twin_evaluate(c, P, tgt, qt_, &Et);
    evals++;
    int accept = 0;
    if (Et.f != Et.f) status = ST_NAN;
    else if (Et.f < P->tol_f) status = ST_STOPVAL;
    else if (!have_cur) accept = 1;
    else if (Et.f < Ec.f) {
      accept = 1;
      double df = Ec.f - Et.f, dx = 0;
      for (int j = 0; j < n; j++) dx = fmax(dx, fabs(qt_[j] - qc[j]));
      if (df < P->tol_df_eff) status = ST_FTOL;
      else if (P->tol_dx > 0.0 && dx < P->tol_dx) status = ST_XTOL;
      slow = (df < P->stall_rel * Ec.f) ? slow + 1 : 0;
      if (status == ST_NONE && slow >= P->stall_count) status = ST_STUCK;
      lambda = fmax(lambda * P->lambda_dec, P->lambda_min);
    } else {
      lambda = lambda * P->lambda_inc;
      if (lambda > P->lambda_max) status = ST_STUCK;
    }
    if (twin_trace && evals <= twin_trace_cap) {
      twin_trace[3 * (evals - 1)] = Et.f; twin_trace[3 * (evals - 1) + 1] = lambda; twin_trace[3 * (evals - 1) + 2] = accept;
    }
    if (status == ST_NONE && evals >= P->max_evals) status = ST_ITERCAP;
    if (status != ST_NONE) break;
    if (accept) { memcpy(qc, qt_, sizeof(double) * n); Ec = Et; have_cur = 1; }
    /* step from the current point */
    double m[MAX_DOF], A[6][6], y[6], tmp[MAX_DOF];
    for (int j = 0; j < n; j++) {
      double g = dot6(Ec.r, Ec.Jr[j]);
      int pinned = (qc[j] <= c->lb[j] && g > 0.0) || (qc[j] >= c->ub[j] && g < 0.0);
      m[j] = pinned ? 0.0 : 1.0;
    }
    for (int a = 0; a < 6; a++)
      for (int b = 0; b <= a; b++) {
        for (int j = 0; j < n; j++) tmp[j] = (m[j] * Ec.Jr[j][a]) * Ec.Jr[j][b];
        A[a][b] = tree_sum(tmp, n);
      }
    for (int a = 0; a < 6; a++) A[a][a] = A[a][a] + lambda;
    ldl6_solve(A, Ec.r, y);
    for (int j = 0; j < n; j++) {
      double Jm[6];
      for (int a = 0; a < 6; a++) Jm[a] = m[j] * Ec.Jr[j][a];
      double dq = -dot6(Jm, y);
      qt_[j] = fmin(fmax(qc[j] + dq, c->lb[j]), c->ub[j]);
    }
  }
  memcpy(q_out, qt_, sizeof(double) * n);
  *f_out = Et.f;
  *evals_out = evals;
  return status;
}

/* ======================= layout 1: thread-per-seed kernel (solve_t1_kernel.cu) =======================
 * Same objective, stop rules and LM step; sequential instead of lane-parallel evaluation order:
 *   - one BACKWARD recursion in the BASE frame on the INVERSE pose C_j = B_j^-1 (B_{j-1} = L_j B_j, B_n = tip):
 *       column of joint j = [ t_C x (R_C a_j) ; R_C a_j ]  ( = [R_Bj^T (a_j x p_Bj) ; R_Bj^T a_j] ),
 *       conj(L_j.q) = conj(cos * origin_q + sin * (origin_q (x) (a_j, 0))),
 *       C.q <- C.q (x) conj(L_j.q),  C.t <- C.t - R_C.q(new) L_j.t ;   X = (C_0 T_tgt)^-1   (objective.rs:48-49)
 *   - Gram matrix and score are accumulated joint by joint with fma (no tree). */
static void twin_eval_t1(const twin_chain* c, const twin_params* P, se3t tgt, const double* q, double* f, double* r,
                         double body[][6], err_coef* ec) {
  int n = c->n;
  se3t C; /* tip^-1 */
  C.q = qt_conj(c->tip.q);
  C.t = v3_neg(qt_rot(C.q, c->tip.t));
  for (int j = n - 1; j >= 0; j--) {
    v3 ax = c->ax[j];
    v3 ang = qt_rot(C.q, ax);
    if (c->type[j] == 0) {
      v3 lin = v3_cross(C.t, ang);
      body[j][0] = lin.x; body[j][1] = lin.y; body[j][2] = lin.z; body[j][3] = ang.x; body[j][4] = ang.y; body[j][5] = ang.z;
    } else {
      body[j][0] = ang.x; body[j][1] = ang.y; body[j][2] = ang.z; body[j][3] = 0; body[j][4] = 0; body[j][5] = 0;
    }
    se3t O;
    O.q = c->oq[j]; O.t = c->ot[j];
    qt lq; /* conj(L.q) */
    v3 lt = O.t;
    if (c->type[j] == 0) {
      double s, cs;
      tw_sincos(0.5 * q[j], &s, &cs);
      qt axq = {ax.x, ax.y, ax.z, 0.0};
      qt oa = qt_mul(O.q, axq);
      lq.x = -fma(cs, O.q.x, s * oa.x); lq.y = -fma(cs, O.q.y, s * oa.y); lq.z = -fma(cs, O.q.z, s * oa.z);
      lq.w = fma(cs, O.q.w, s * oa.w);
    } else {
      lq = qt_conj(O.q);
      lt = v3_add(O.t, qt_rot(O.q, v3_scale(ax, q[j])));
    }
    C.q = qt_mul(C.q, lq);
    C.t = v3_sub(C.t, qt_rot(C.q, lt));
  }
  qt xiq = qt_mul(C.q, tgt.q); /* X^-1 = C_0 T_tgt */
  v3 xit = v3_add(C.t, qt_rot(C.q, tgt.t));
  qt xq = qt_conj(xiq);
  v3 xt = v3_neg(qt_rot(xq, xit));
  v3 elin;
  error_terms(xq, xt, ec, &elin);
  v3 rl = elin, ra = ec->w;
  if (P->weighted) { rl = weight3(tgt.q, P->wl, elin); ra = weight3(tgt.q, P->wa, ec->w); }
  r[0] = rl.x; r[1] = rl.y; r[2] = rl.z; r[3] = ra.x; r[4] = ra.y; r[5] = ra.z;
  *f = dot6(r, r);
}

static int twin_attempt_t1(const twin_chain* c, const twin_params* P, se3t tgt, const double* q_init, double* q_out,
                           double* f_out, int* evals_out) {
  int n = c->n, evals = 0, have_cur = 0, status = ST_NONE, slow = 0;
  double qc[MAX_DOF], qt_[MAX_DOF], lambda = P->lambda0, fc = 0, ft, rc[6], rt[6];
  double body[MAX_DOF][6], C[MAX_DOF][6];
  err_coef ec;
  for (int j = 0; j < n; j++) qt_[j] = fmin(fmax(q_init[j], c->lb[j]), c->ub[j]);
  for (;;) {
    twin_eval_t1(c, P, tgt, qt_, &ft, rt, body, &ec);
    evals++;
    int accept = 0;
    if (ft != ft) status = ST_NAN;
    else if (ft < P->tol_f) status = ST_STOPVAL;
    else if (!have_cur) accept = 1;
    else if (ft < fc) {
      accept = 1;
      double df = fc - ft, dx = 0;
      for (int j = 0; j < n; j++) dx = fmax(dx, fabs(qt_[j] - qc[j]));
      if (df < P->tol_df_eff) status = ST_FTOL;
      else if (P->tol_dx > 0.0 && dx < P->tol_dx) status = ST_XTOL;
      slow = (df < P->stall_rel * fc) ? slow + 1 : 0;
      if (status == ST_NONE && slow >= P->stall_count) status = ST_STUCK;
      lambda = fmax(lambda * P->lambda_dec, P->lambda_min);
    } else {
      lambda = lambda * P->lambda_inc;
      if (lambda > P->lambda_max) status = ST_STUCK;
    }
    if (twin_trace && evals <= twin_trace_cap) {
      twin_trace[3 * (evals - 1)] = ft; twin_trace[3 * (evals - 1) + 1] = lambda; twin_trace[3 * (evals - 1) + 2] = accept;
    }
    if (status == ST_NONE && evals >= P->max_evals) status = ST_ITERCAP;
    if (status != ST_NONE) break;
    if (accept) { /* current point <- trial: task columns from the trial's body columns */
      memcpy(qc, qt_, sizeof(double) * n);
      memcpy(rc, rt, sizeof(rc));
      fc = ft; have_cur = 1;
      double Jm[9], CJ[9];
      task_mats(&ec, Jm, CJ);
      for (int j = 0; j < n; j++) {
        v3 top, bot;
        task_col_m(Jm, CJ, v3_make(body[j][0], body[j][1], body[j][2]), v3_make(body[j][3], body[j][4], body[j][5]), &top, &bot);
        if (P->weighted) { top = weight3(tgt.q, P->wl, top); bot = weight3(tgt.q, P->wa, bot); }
        C[j][0] = top.x; C[j][1] = top.y; C[j][2] = top.z; C[j][3] = bot.x; C[j][4] = bot.y; C[j][5] = bot.z;
      }
    }
    double A[6][6], y[6], m[MAX_DOF];
    for (int a = 0; a < 6; a++) for (int b = 0; b <= a; b++) A[a][b] = 0.0;
    for (int j = 0; j < n; j++) {
      double g = dot6(rc, C[j]);
      int pinned = (qc[j] <= c->lb[j] && g > 0.0) || (qc[j] >= c->ub[j] && g < 0.0);
      m[j] = pinned ? 0.0 : 1.0;
      for (int a = 0; a < 6; a++) {
        double jm = m[j] * C[j][a];
        for (int b = 0; b <= a; b++) A[a][b] = fma(jm, C[j][b], A[a][b]);
      }
    }
    for (int a = 0; a < 6; a++) A[a][a] = A[a][a] + lambda;
    ldl6_solve(A, rc, y);
    for (int j = 0; j < n; j++) {
      double x = qc[j] - m[j] * dot6(C[j], y);
      x = x < c->lb[j] ? c->lb[j] : x;
      x = x > c->ub[j] ? c->ub[j] : x;
      qt_[j] = x;
    }
  }
  memcpy(q_out, qt_, sizeof(double) * n);
  *f_out = ft;
  *evals_out = evals;
  return status;
}

int twin_status_success(const twin_params* P, int status) { /* lib.rs:376-379 */
  return (P->tol_f >= 0.0 && status == ST_STOPVAL) || (P->tol_df_user >= 0.0 && status == ST_FTOL) ||
         (P->tol_dx >= 0.0 && status == ST_XTOL);
}

/* -------- entry points for ctypes -------- */
void oracle_restart_seed(uint64_t restart, const double* lb, const double* ub, int n, double* q);

/* evaluate with the twin arithmetic: outputs ee pose8, f, r[6], Jr (n x 6 row-major), grad[n] = 2 r^T Jr */
int twin_eval_c(const double* chain, int njoints, const double* ee_offset, const twin_params* P, const double* target,
                const double* q, double* ee8, double* f, double* r6, double* Jr, double* grad) {
  twin_chain c;
  int rc = twin_chain_init(&c, chain, njoints, ee_offset);
  if (rc) return rc;
  twin_eval E;
  twin_evaluate(&c, P, se3_from_pose8(target), q, &E);
  E.ee = se3_mul(se3_from_pose8(target), E.ee); /* back to the world frame for reporting */
  ee8[0] = E.ee.q.x; ee8[1] = E.ee.q.y; ee8[2] = E.ee.q.z; ee8[3] = E.ee.q.w;
  ee8[4] = E.ee.t.x; ee8[5] = E.ee.t.y; ee8[6] = E.ee.t.z; ee8[7] = 0;
  *f = E.f;
  memcpy(r6, E.r, sizeof(E.r));
  for (int j = 0; j < c.n; j++) {
    memcpy(Jr + 6 * j, E.Jr[j], 6 * sizeof(double));
    double g = dot6(E.r, E.Jr[j]);
    grad[j] = g + g;
  }
  return 0;
}

/* Run restarts [r_begin, r_end) (restart 0 = x0, i>=1 = ChaCha8 stream i) for one target and record each attempt.
 * q_all: (r_end-r_begin) x n, f_all, status_all, evals_all: per attempt. */
int twin_attempts_c(const double* chain, int njoints, const double* ee_offset, const twin_params* P,
                    const double* target, const double* x0, uint64_t r_begin, uint64_t r_end, double* q_all,
                    double* f_all, int* status_all, int* evals_all) {
  twin_chain c;
  int rc = twin_chain_init(&c, chain, njoints, ee_offset);
  if (rc) return rc;
  se3t tgt = se3_from_pose8(target);
  for (uint64_t r = r_begin; r < r_end; r++) {
    double qi[MAX_DOF];
    if (r == 0) memcpy(qi, x0, sizeof(double) * c.n);
    else oracle_restart_seed(r, c.slb, c.sub, c.n, qi);
    size_t o = (size_t)(r - r_begin);
    status_all[o] = twin_attempt(&c, P, tgt, qi, q_all + o * c.n, f_all + o, evals_all + o);
  }
  return 0;
}

/* Robot::ik selection semantics in single-thread order (lib.rs:397-413) over restarts [r_begin, r_end).
 * mode 2 = Speed: first converged restart.  mode 1 = Quality: arg-min ||q-x0||^2, ties -> lower index.
 * Returns 1 if a restart converged (outputs describe it), else 0 (outputs describe the first attempt). */
int twin_ik_c(const double* chain, int njoints, const double* ee_offset, const twin_params* P, const double* target,
              const double* x0, uint64_t r_begin, uint64_t r_end, int mode, double* q_out, double* f_out,
              int* status_out, uint64_t* restart_out) {
  twin_chain c;
  int rc = twin_chain_init(&c, chain, njoints, ee_offset);
  if (rc) return rc;
  se3t tgt = se3_from_pose8(target);
  int have = 0, first = 1;
  double best_score = 0;
  for (uint64_t r = r_begin; r < r_end; r++) {
    double qi[MAX_DOF], q[MAX_DOF], f, d2[MAX_DOF];
    int evals;
    if (r == 0) memcpy(qi, x0, sizeof(double) * c.n);
    else oracle_restart_seed(r, c.slb, c.sub, c.n, qi);
    int st = twin_attempt(&c, P, tgt, qi, q, &f, &evals);
    int ok = twin_status_success(P, st);
    if (ok) {
      double score = 0.0;
      if (mode == 1) {
        if (P->layout == 1) {
          for (int j = 0; j < c.n; j++) { double d = q[j] - x0[j]; score = fma(d, d, score); }
        } else {
          for (int j = 0; j < c.n; j++) { double d = q[j] - x0[j]; d2[j] = d * d; }
          score = tree_sum(d2, c.n);
        }
      }
      if (!have || score < best_score) {
        have = 1; best_score = score;
        memcpy(q_out, q, sizeof(double) * c.n); *f_out = f; *status_out = st; *restart_out = r;
      }
      if (mode == 2) break;
    } else if (!have && first) {
      memcpy(q_out, q, sizeof(double) * c.n); *f_out = f; *status_out = st; *restart_out = r;
    }
    first = 0;
  }
  return have;
}
