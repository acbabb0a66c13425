/*
 * twin_math.h -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * Deterministic fp64 primitives of the "kernel arithmetic spec" (DESIGN.md
 * section "Arithmetic spec").  The CUDA kernel (optik_b200/csrc/dmath.cuh)
 * implements the same sequence of IEEE-754 operations -- every fused
 * multiply-add is an explicit fma(), everything else is a separately rounded
 * +,-,*,/,sqrt -- so that GPU and CPU results are bit-identical.  This file is
 * a restatement of that spec in plain C, not shared source: the kernel never
 * includes it.
 *
 * sin/cos: Cody-Waite reduction by pi/2 (3 constants) + the classic fdlibm
 * minimax polynomials on [-pi/4, pi/4]; atan: fdlibm breakpoints/polynomial.
 * (Published algorithms; coefficients verified against libm in tests.)
 */
#ifndef TWIN_MATH_H
#define TWIN_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { double x, y, z; } v3;
typedef struct { double x, y, z, w; } qt;
typedef struct { qt q; v3 t; } se3t;

static inline v3 v3_make(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, double s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
/* a + s*b */
static inline v3 v3_axpy(double s, v3 b, v3 a) { return v3_make(fma(s, b.x, a.x), fma(s, b.y, a.y), fma(s, b.z, a.z)); }
static inline double v3_dot(v3 a, v3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
static inline v3 v3_cross(v3 a, v3 b) {
  return v3_make(fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x)));
}
static inline qt qt_mul(qt a, qt b) {
  qt r;
  r.w = fma(a.w, b.w, -fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)));
  r.x = fma(a.w, b.x, fma(a.x, b.w, fma(a.y, b.z, -(a.z * b.y))));
  r.y = fma(a.w, b.y, fma(a.y, b.w, fma(a.z, b.x, -(a.x * b.z))));
  r.z = fma(a.w, b.z, fma(a.z, b.w, fma(a.x, b.y, -(a.y * b.x))));
  return r;
}
static inline qt qt_conj(qt a) { qt r = {-a.x, -a.y, -a.z, a.w}; return r; }
/* v + w*(2 u x v) + u x (2 u x v) */
static inline v3 qt_rot(qt q, v3 v) {
  v3 u = v3_make(q.x, q.y, q.z);
  v3 t = v3_cross(u, v);
  t = v3_add(t, t);
  v3 c = v3_cross(u, t);
  return v3_make(fma(q.w, t.x, v.x) + c.x, fma(q.w, t.y, v.y) + c.y, fma(q.w, t.z, v.z) + c.z);
}
static inline v3 qt_rot_inv(qt q, v3 v) { return qt_rot(qt_conj(q), v); }
/* A*B */
static inline se3t se3_mul(se3t a, se3t b) {
  se3t r;
  r.t = v3_add(a.t, qt_rot(a.q, b.t));
  r.q = qt_mul(a.q, b.q);
  return r;
}

/* ---- sin/cos ---- */
static inline void tw_sincos(double x, double* sn, double* cs) {
  const double TWO_OVER_PI = 6.36619772367581382433e-01;
  const double PIO2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
  const double PIO2_2 = 6.07710050630396597660e-11;  /* next 33 bits */
  const double PIO2_3 = 2.02226624871116645580e-21;  /* remainder */
  /* rint by the 1.5*2^52 shift, exactly as the kernels do it (csrc/dmath.cuh) */
  const double SHIFT = 6755399441055744.0;
  double kk = x * TWO_OVER_PI + SHIFT;
  double k = kk - SHIFT;
  double r = fma(-k, PIO2_1, x);
  r = fma(-k, PIO2_2, r);
  r = fma(-k, PIO2_3, r);
  double z = r * r;
  /* sin(r) = r + r*z*(S1 + z*(S2 + ... )) */
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(z, ps, 2.75573137070700676789e-06);
  ps = fma(z, ps, -1.98412698298579493134e-04);
  ps = fma(z, ps, 8.33333333332248946124e-03);
  ps = fma(z, ps, -1.66666666666666324348e-01);
  double s = fma(r * z, ps, r);
  /* cos(r) = 1 - z/2 + z*z*(C1 + z*(C2 + ...)) */
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(z, pc, -2.75573143513906633035e-07);
  pc = fma(z, pc, 2.48015872894767294178e-05);
  pc = fma(z, pc, -1.38888888888741095749e-03);
  pc = fma(z, pc, 4.16666666666666019037e-02);
  double c = fma(z * z, pc, fma(z, -0.5, 1.0));
  uint64_t kbits;
  memcpy(&kbits, &kk, sizeof kbits);
  int q = (int)(kbits & 3u);
  double ss = (q & 1) ? c : s;
  double cc = (q & 1) ? s : c;
  if (q == 1 || q == 2) cc = -cc;
  if (q >= 2) ss = -ss;
  *sn = ss;
  *cs = cc;
}

/* ---- atan for t in [0,1] and atan2(y,x) for y>=0, x>=0 (not both 0) ---- */
static inline double tw_atan_poly(double x) { /* |x| <= 7/16, fdlibm kernel */
  double z = x * x, w = z * z;
  double s1 = fma(w, 1.62858201153657823623e-02, 4.97687799461593236017e-02);
  s1 = fma(w, s1, 6.66107313738753120669e-02);
  s1 = fma(w, s1, 9.09088713343650656196e-02);
  s1 = fma(w, s1, 1.42857142725034663711e-01);
  s1 = fma(w, s1, 3.33333333333329318027e-01);
  s1 = z * s1;
  double s2 = fma(w, -3.65315727442169155270e-02, -5.83357013379057348645e-02);
  s2 = fma(w, s2, -7.69187620504482999495e-02);
  s2 = fma(w, s2, -1.11111104054623557880e-01);
  s2 = fma(w, s2, -1.99999999998764832476e-01);
  s2 = w * s2;
  return x - x * (s1 + s2);
}
static inline double tw_atan01(double t) { /* 0 <= t <= 1 */
  if (t < 0.4375) return tw_atan_poly(t);
  if (t < 0.6875) { /* atan(0.5) + atan((2t-1)/(2+t)) */
    double u = (t + t - 1.0) / (2.0 + t);
    return 4.63647609000806093515e-01 + (tw_atan_poly(u) + 2.26987774529616870924e-17);
  }
  { /* atan(1) + atan((t-1)/(t+1)) */
    double u = (t - 1.0) / (t + 1.0);
    return 7.85398163397448278999e-01 + (tw_atan_poly(u) + 3.06161699786838301793e-17);
  }
}
static inline double tw_atan2_pos(double y, double x) {
  if (y <= x) return tw_atan01(y / x);
  return 1.57079632679489655800e+00 - (tw_atan01(x / y) - 6.12323399573676603587e-17);
}
#endif
