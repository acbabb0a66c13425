/*
 * optik_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C fp64 restatement of the reference's (kylc/optik @ 355e463) IK hot
 * path, written from the math, not translated.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may link or
 * call this file; the product (optik_b200/csrc) never does.
 *
 * Parity status:
 *   - evaluator (FK, body Jacobian, so3/se3 log, right Jacobians, objective,
 *     gradient): PINNED against the reference's own golden vectors
 *     (crates/optik/tests/data/*.json, tests/test_fk.rs, tests/test_math.rs)
 *     and the reference's FD-gradient property (tests/test_gradient.rs).
 *   - seed generator (ChaCha8 / PCG32 seed expansion / f64 uniform): restated
 *     from the published algorithms; the reference holds no golden seed values
 *     and rand/rand_chacha sources are not vendored => "parity unpinned"
 *     against Rust; pinned only by the RFC 7539 quarter-round vector.
 *   - inner optimiser: the reference calls NLopt SLSQP (un-vendored git
 *     dependency kylc/rust-nlopt@8e731e3, crate nlopt 0.8.1).  It is REPLACED,
 *     by design, with a projected Levenberg-Marquardt solve; its fp64 twin
 *     lives in solver_twin.c.  "parity unpinned" for per-seed trajectories.
 *
 * Conventions: pose8 = {qx,qy,qz,qw, tx,ty,tz, pad}; matrices column-major
 * like nalgebra; chain = 16 doubles per joint:
 *   [0..2] origin xyz  [3] type (0 revolute, 1 prismatic, 2 fixed)
 *   [4..7] origin quaternion xyzw  [8..10] axis  [12] lower  [13] upper
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define CHAIN_STRIDE 16
#define EPSILON 1e-6 /* crates/optik/src/math.rs:7 */

/* ------------------------------------------------------------------ basics */
static void q_mul(const double a[4], const double b[4], double o[4]) {
  double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by - ax * bz + ay * bw + az * bx;
  o[2] = aw * bz + ax * by - ay * bx + az * bw;
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
}
static void q_to_mat(const double q[4], double R[9] /* column-major */) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[3] = 2 * (x * y - z * w);     R[6] = 2 * (x * z + y * w);
  R[1] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[7] = 2 * (y * z - x * w);
  R[2] = 2 * (x * z - y * w);     R[5] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
static void m_vec(const double R[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
static void mt_vec(const double R[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
static void q_rot(const double q[4], const double v[3], double o[3]) {
  double R[9]; q_to_mat(q, R); m_vec(R, v, o);
}
static void q_rot_inv(const double q[4], const double v[3], double o[3]) {
  double R[9]; q_to_mat(q, R); mt_vec(R, v, o);
}
static void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
/* pose composition A*B  (Isometry3 product) */
static void pose_mul(const double a[8], const double b[8], double o[8]) {
  double r[3], q[4];
  q_rot(a, b + 4, r);
  q_mul(a, b, q);
  o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3];
  o[4] = a[4] + r[0]; o[5] = a[5] + r[1]; o[6] = a[6] + r[2]; o[7] = 0;
}
/* A^-1 * B  (Isometry3::inv_mul, objective.rs:49) */
static void pose_inv_mul(const double a[8], const double b[8], double o[8]) {
  double ac[4] = {-a[0], -a[1], -a[2], a[3]}, d[3] = {b[4] - a[4], b[5] - a[5], b[6] - a[6]}, r[3], q[4];
  q_mul(ac, b, q);
  q_rot_inv(a, d, r);
  o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3];
  o[4] = r[0]; o[5] = r[1]; o[6] = r[2]; o[7] = 0;
}
static const double POSE_ID[8] = {0, 0, 0, 1, 0, 0, 0, 0};

/* 3x3 column-major helpers */
static void m3_hat(const double w[3], double M[9]) { /* math.rs:13-15 */
  M[0] = 0;     M[3] = -w[2]; M[6] = w[1];
  M[1] = w[2];  M[4] = 0;     M[7] = -w[0];
  M[2] = -w[1]; M[5] = w[0];  M[8] = 0;
}
static void m3_hat2(const double w[3], double M[9]) { /* math.rs:18-31 */
  double w11 = w[0] * w[0], w12 = w[0] * w[1], w13 = w[0] * w[2];
  double w22 = w[1] * w[1], w23 = w[1] * w[2], w33 = w[2] * w[2];
  M[0] = -w22 - w33; M[3] = w12;        M[6] = w13;
  M[1] = w12;        M[4] = -w11 - w33; M[7] = w23;
  M[2] = w13;        M[5] = w23;        M[8] = -w11 - w22;
}
static void m3_mul(const double A[9], const double B[9], double C[9]) {
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += A[3 * k + r] * B[3 * c + k];
      C[3 * c + r] = s;
    }
}

/* ------------------------------------------------- math.rs restatement */
/* so3::log, math.rs:40-63 */
void oracle_so3_log(const double q[4], double w_out[3]) {
  double w = q[3], v[3] = {q[0], q[1], q[2]};
  if (!(w >= 0.0)) { w = -w; v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }
  double vn2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], k;
  if (vn2 > EPSILON) {
    double vn = sqrt(vn2);
    k = atan2(vn, w) / vn;
  } else {
    k = 1. / w - 1. / (3. * w * w * w) * vn2 + 1. / (5. * w * w * w * w * w) * vn2 * vn2;
  }
  for (int i = 0; i < 3; i++) w_out[i] = 2.0 * v[i] * k;
}

/* so3::right_jacobian, math.rs:72-94.  Deviation: at
 * theta^2 == 0 exactly the reference evaluates 0/0; we return the limit. */
void oracle_so3_right_jacobian(const double w[3], double J[9]) {
  double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], t4 = t2 * t2, th = sqrt(t2);
  double s = sin(th), c = cos(th);
  double a = (t2 > EPSILON) ? s / th : 1. - 1. / 6. * t2 + 1. / 120. * t4;
  double b = (t2 > EPSILON) ? (1. - c) / t2 : 1. / 2. - 1. / 24. * t2 + 1. / 720. * t4;
  double cc = (t2 > 0.0) ? (1. - a) / t2 : 1. / 6.;
  double e = (b - 2. * cc) / (2. * a);
  double H[9], H2[9];
  m3_hat(w, H); m3_hat2(w, H2);
  for (int i = 0; i < 9; i++) J[i] = 0.5 * H[i] + e * H2[i];
  J[0] += 1; J[4] += 1; J[8] += 1;
}

/* se3::log, math.rs:107-124 (same exact-zero guard) */
void oracle_se3_log(const double X[8], double e[6]) {
  double w[3];
  oracle_so3_log(X, w);
  double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(t2), p;
  if (th > EPSILON) {
    p = 0.5 * (th * sin(th)) / (1. - cos(th));
  } else {
    p = 1. - t2 / 12. - t2 * t2 / 720.;
  }
  double k = (t2 > 0.0) ? 1. / t2 * (1. - p) : 1. / 12.;
  double H[9], H2[9], Vi[9];
  m3_hat(w, H); m3_hat2(w, H2);
  for (int i = 0; i < 9; i++) Vi[i] = -0.5 * H[i] + k * H2[i];
  Vi[0] += 1; Vi[4] += 1; Vi[8] += 1;
  m_vec(Vi, X + 4, e);
  e[3] = w[0]; e[4] = w[1]; e[5] = w[2];
}

/* se3::right_jacobian_q_matrix, math.rs:135-170 */
static void se3_q_matrix(const double v[3], const double w[3], double Q[9]) {
  double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(t2), t4 = t2 * t2, a, b;
  if (t2 > EPSILON) {
    double s = sin(th), c = cos(th), s_t = s / th, inv_1mc = 1. / (2. * (1. - c));
    a = 1. / t2 - s_t * inv_1mc;
    b = -2. / t4 + (1. + s_t) * inv_1mc / t2;
  } else {
    a = 1. / 12. + t2 / 720.;
    b = 1. / 360.;
  }
  double d = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
  double cv[3];
  for (int i = 0; i < 3; i++) cv[i] = b * d * w[i] - (t2 * b + 2. * a) * v[i];
  double C[9], H[9];
  m3_hat(v, H);
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++)
      C[3 * c + r] = 0.5 * H[3 * c + r] + cv[r] * w[c] + a * w[r] * v[c] + (r == c ? d * a : 0.0);
  double E[9];
  oracle_so3_right_jacobian(w, E);
  m3_mul(C, E, Q);
}

/* se3::right_jacobian, math.rs:191-203 -> 6x6 column-major */
void oracle_se3_right_jacobian(const double X[8], double U[36]) {
  double w[3], J[9], Q[9];
  oracle_so3_log(X, w);
  oracle_so3_right_jacobian(w, J);
  se3_q_matrix(X + 4, w, Q);
  memset(U, 0, 36 * sizeof(double));
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) {
      U[6 * c + r] = J[3 * c + r];
      U[6 * (c + 3) + r] = Q[3 * c + r];
      U[6 * (c + 3) + (r + 3)] = J[3 * c + r];
    }
}

/* ------------------------------------------- kinematics.rs restatement */
/* JointType::local_transform, kinematics.rs:243-255 */
static void local_transform(const double* j, double qi, double L[8]) {
  int type = (int)j[3];
  memcpy(L, POSE_ID, sizeof(POSE_ID));
  if (type == 0) {
    double s = sin(0.5 * qi), c = cos(0.5 * qi);
    L[0] = j[8] * s; L[1] = j[9] * s; L[2] = j[10] * s; L[3] = c;
  } else if (type == 1) {
    L[4] = j[8] * qi; L[5] = j[9] * qi; L[6] = j[10] * qi;
  }
}
int oracle_num_positions(const double* chain, int njoints) {
  int n = 0;
  for (int i = 0; i < njoints; i++) n += ((int)chain[CHAIN_STRIDE * i + 3] != 2);
  return n;
}
/* forward_kinematics_mut, kinematics.rs:123-164.  joint_tfms: njoints x pose8 */
void oracle_fk(const double* chain, int njoints, const double* q, const double* ee_offset /* pose8 or NULL */,
               double* joint_tfms, double* ee) {
  double T[8];
  memcpy(T, POSE_ID, sizeof(T));
  int qi = 0;
  for (int i = 0; i < njoints; i++) {
    const double* j = chain + CHAIN_STRIDE * i;
    double origin[8] = {j[4], j[5], j[6], j[7], j[0], j[1], j[2], 0}, L[8], OL[8], Tn[8];
    int nq = ((int)j[3] != 2);
    local_transform(j, nq ? q[qi] : 0.0, L);
    pose_mul(origin, L, OL);
    pose_mul(T, OL, Tn);
    memcpy(T, Tn, sizeof(T));
    memcpy(joint_tfms + 8 * i, T, sizeof(T));
    qi += nq;
  }
  pose_mul(T, ee_offset ? ee_offset : POSE_ID, ee);
}
/* joint_jacobian, kinematics.rs:166-196 -> 6 x n column-major, rows [lin; ang].
 * Prismatic columns: the reference panics (todo!, :185); we give the
 * geometric column [R_ee^T (R_i axis); 0] and say so in DESIGN.md. */
void oracle_joint_jacobian(const double* chain, int njoints, const double* joint_tfms, const double* ee, double* J) {
  int col = 0;
  for (int i = 0; i < njoints; i++) {
    const double* j = chain + CHAIN_STRIDE * i;
    const double* T = joint_tfms + 8 * i;
    int type = (int)j[3];
    if (type == 2) continue;
    double ang[3], lin[3], d[3] = {ee[4] - T[4], ee[5] - T[5], ee[6] - T[6]};
    q_rot(T, j + 8, ang);
    if (type == 0) {
      cross3(ang, d, lin);
      q_rot_inv(ee, lin, J + 6 * col);
      q_rot_inv(ee, ang, J + 6 * col + 3);
    } else {
      q_rot_inv(ee, ang, J + 6 * col);
      J[6 * col + 3] = J[6 * col + 4] = J[6 * col + 5] = 0;
    }
    col++;
  }
}

/* -------------------------------------------- objective.rs restatement */
/* apply_weighting, objective.rs:7-38 (always applied; numerically a no-op for unit weights) */
static void apply_weighting(double e[6], const double tgt[8], const double wl[3], const double wa[3]) {
  for (int blk = 0; blk < 2; blk++) {
    const double* w = blk ? wa : wl;
    double a[3], b[3];
    q_rot(tgt, e + 3 * blk, a);
    a[0] *= w[0]; a[1] *= w[1]; a[2] *= w[2];
    q_rot_inv(tgt, a, b);
    memcpy(e + 3 * blk, b, sizeof(b));
  }
}
/* objective, objective.rs:40-57 */
double oracle_objective(const double* chain, int njoints, const double* q, const double* target,
                        const double* ee_offset, const double wl[3], const double wa[3]) {
  double tf[8 * 64], ee[8], X[8], e[6];
  double* tfms = njoints <= 64 ? tf : (double*)malloc(sizeof(double) * 8 * njoints);
  oracle_fk(chain, njoints, q, ee_offset, tfms, ee);
  pose_inv_mul(target, ee, X);
  oracle_se3_log(X, e);
  apply_weighting(e, target, wl, wa);
  if (tfms != tf) free(tfms);
  return e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3] + e[4] * e[4] + e[5] * e[5];
}
/* objective_grad, objective.rs:60-110; also returns the 6-vector e (unweighted) and ee pose if wanted */
void oracle_objective_grad(const double* chain, int njoints, const double* q, const double* target,
                           const double* ee_offset, const double wl[3], const double wa[3], double* g) {
  int n = oracle_num_positions(chain, njoints);
  double* tfms = (double*)malloc(sizeof(double) * 8 * njoints);
  double* J = (double*)malloc(sizeof(double) * 6 * n);
  double ee[8], X[8], U[36], e[6];
  oracle_fk(chain, njoints, q, ee_offset, tfms, ee);
  pose_inv_mul(target, ee, X);
  oracle_joint_jacobian(chain, njoints, tfms, ee, J);
  oracle_se3_right_jacobian(X, U);
  oracle_se3_log(X, e);
  double wl2[3] = {wl[0] * wl[0], wl[1] * wl[1], wl[2] * wl[2]};
  double wa2[3] = {wa[0] * wa[0], wa[1] * wa[1], wa[2] * wa[2]};
  apply_weighting(e, target, wl2, wa2);
  for (int c = 0; c < n; c++) {
    double s = 0;
    for (int r = 0; r < 6; r++) {
      double jt = 0; /* (Jlog6 * J)[r][c] */
      for (int k = 0; k < 6; k++) jt += U[6 * k + r] * J[6 * c + k];
      s += 2.0 * e[r] * jt;
    }
    g[c] = s;
  }
  free(tfms); free(J);
}

/* ------------------------------------------------- seeds (lib.rs:86-91, 360-370) */
static uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
#define QR(a, b, c, d) \
  a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
  a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);
void oracle_chacha_quarter_round(uint32_t v[4]) { QR(v[0], v[1], v[2], v[3]); }
/* ChaCha with 8 rounds; state per rand_chacha: 4 consts | 8 key | 64-bit block counter | 64-bit stream */
void oracle_chacha8_block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint32_t out[16]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                    key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t x[16];
  memcpy(x, s, sizeof(x));
  for (int r = 0; r < 4; r++) {
    QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
    QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}
/* rand_core SeedableRng::seed_from_u64: PCG32 expansion of the u64 into the 32-byte key */
void oracle_seed_key(uint64_t state, uint32_t key[8]) {
  for (int i = 0; i < 8; i++) {
    state = state * 6364136223846793005ULL + 11634580027462260723ULL;
    uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
  }
}
/* k-th u64 of ChaCha8Rng(seed).set_stream(stream) */
uint64_t oracle_rng_u64(uint64_t seed, uint64_t stream, uint32_t k) {
  uint32_t key[8], blk[16];
  oracle_seed_key(seed, key);
  oracle_chacha8_block(key, k / 8, stream, blk);
  return (uint64_t)blk[2 * (k % 8)] | ((uint64_t)blk[2 * (k % 8) + 1] << 32);
}
/* rand::Rng::random_range(lb..=ub) for f64: 52 random mantissa bits -> [0,1), scaled */
double oracle_uniform(uint64_t u, double lb, double ub) {
  uint64_t bits = (u >> 12) | 0x3FF0000000000000ULL;
  double x12, x01;
  memcpy(&x12, &bits, 8);
  x01 = x12 - 1.0;
  double scale = (ub - lb) / (1.0 - 2.220446049250313e-16);
  double v = x01 * scale + lb;
  return v > ub ? ub : v;
}
/* Robot::random_configuration with ChaCha8Rng::seed_from_u64(42).set_stream(restart), lib.rs:360-370 */
void oracle_restart_seed(uint64_t restart, const double* lb, const double* ub, int n, double* q) {
  for (int i = 0; i < n; i++) q[i] = oracle_uniform(oracle_rng_u64(42, restart, (uint32_t)i), lb[i], ub[i]);
}

/* ------------------------------------------------------------------ evaluator batch (bench.py: the CPU number beside the
 * evaluator kernel's roofline).  Per configuration the work of ONE objective callback of the reference with a gradient
 * request (lib.rs:305-337): forward_kinematics_mut once, objective_grad (joint_jacobian + se3::right_jacobian + se3::log +
 * weighting + the 6 x n product, objective.rs:60-110), then objective (se3::log again, objective.rs:40-57); plus the
 * outputs the CUDA evaluator writes (ee pose, the body Jacobian itself).  `threads` pthread workers over contiguous
 * ranges, no allocation inside the loop. */
#include <pthread.h>
typedef struct {
  const double* chain; int njoints, n;
  const double *q, *targets;
  double *ee, *jac, *f, *grad;
  uint64_t begin, end;
} evjob_t;

static void* eval_worker(void* arg) {
  evjob_t* W = (evjob_t*)arg;
  const double one[3] = {1.0, 1.0, 1.0};
  double* tf = (double*)malloc(sizeof(double) * 8 * (size_t)W->njoints);
  for (uint64_t i = W->begin; i < W->end; i++) {
    const double* q = W->q + i * (uint64_t)W->n;
    const double* tg = W->targets + 8 * i;
    double* ee = W->ee + 8 * i;
    double* J = W->jac + i * 6ull * (uint64_t)W->n;
    double X[8], U[36], e[6], e2[6];
    oracle_fk(W->chain, W->njoints, q, NULL, tf, ee);
    pose_inv_mul(tg, ee, X);
    oracle_joint_jacobian(W->chain, W->njoints, tf, ee, J);
    oracle_se3_right_jacobian(X, U);
    oracle_se3_log(X, e);
    apply_weighting(e, tg, one, one); /* squared weights, objective.rs:102-104 */
    for (int c = 0; c < W->n; c++) {
      double s = 0;
      for (int r = 0; r < 6; r++) {
        double jt = 0;
        for (int k = 0; k < 6; k++) jt += U[6 * k + r] * J[6 * c + k];
        s += 2.0 * e[r] * jt;
      }
      W->grad[i * (uint64_t)W->n + c] = s;
    }
    oracle_se3_log(X, e2); /* objective(): the log once more */
    apply_weighting(e2, tg, one, one);
    W->f[i] = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2] + e2[3] * e2[3] + e2[4] * e2[4] + e2[5] * e2[5];
  }
  free(tf);
  return NULL;
}

int oracle_eval_batch_threaded(const double* chain, int njoints, const double* q, const double* targets, uint64_t B,
                               int threads, double* ee, double* jac, double* f, double* grad) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  const int n = oracle_num_positions(chain, njoints);
  pthread_t th[256];
  evjob_t jobs[256];
  for (int t = 0; t < threads; t++) {
    jobs[t] = (evjob_t){chain, njoints, n, q, targets, ee, jac, f, grad, B * (uint64_t)t / threads, B * (uint64_t)(t + 1) / threads};
    if (pthread_create(&th[t], NULL, eval_worker, &jobs[t]) != 0) return -1;
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  return 0;
}
