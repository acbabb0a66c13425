// dmath.cuh -- fp64 device primitives for the IK kernels (sm_100a).
//
// Arithmetic spec (DESIGN.md "Arithmetic spec"): this translation unit is
// compiled with -fmad=false, so the compiler never contracts a*b+c; every
// fused multiply-add below is an explicit fma().  Together with our own
// sin/cos/atan (no libdevice transcendentals) every result is a fixed sequence
// of IEEE-754 operations, which makes a solve bit-reproducible across GPUs and
// against the CPU twin in oracle/ (a separate restatement, used only by tests).
//
// What the functions restate from the reference (kylc/optik @ 355e463):
//   quaternion/isometry algebra      nalgebra Isometry3<f64> (kinematics.rs:149-163)
//   so3/se3 log and d-log scalars    crates/optik/src/math.rs:40-203
#pragma once
#include <cstdint>

namespace optik {

struct v3 { double x, y, z; };
struct qt { double x, y, z, w; };
struct se3 { qt q; v3 t; };

#define DEV __device__ __forceinline__

DEV v3 mk3(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
DEV v3 add3(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
DEV v3 sub3(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
DEV v3 scale3(v3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
DEV v3 neg3(v3 a) { return mk3(-a.x, -a.y, -a.z); }
// a + s*b
DEV v3 axpy3(double s, v3 b, v3 a) { return mk3(fma(s, b.x, a.x), fma(s, b.y, a.y), fma(s, b.z, a.z)); }
DEV double dot3(v3 a, v3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
DEV v3 cross3(v3 a, v3 b) {
  return mk3(fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x)));
}
DEV qt qmul(qt a, qt b) {
  qt r;
  r.w = fma(a.w, b.w, -fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)));
  r.x = fma(a.w, b.x, fma(a.x, b.w, fma(a.y, b.z, -(a.z * b.y))));
  r.y = fma(a.w, b.y, fma(a.y, b.w, fma(a.z, b.x, -(a.x * b.z))));
  r.z = fma(a.w, b.z, fma(a.z, b.w, fma(a.x, b.y, -(a.y * b.x))));
  return r;
}
DEV qt qconj(qt a) { qt r; r.x = -a.x; r.y = -a.y; r.z = -a.z; r.w = a.w; return r; }
// v + w (2 u x v) + u x (2 u x v)
DEV v3 qrot(qt q, v3 v) {
  v3 u = mk3(q.x, q.y, q.z);
  v3 t = cross3(u, v);
  t = add3(t, t);
  v3 c = cross3(u, t);
  return mk3(fma(q.w, t.x, v.x) + c.x, fma(q.w, t.y, v.y) + c.y, fma(q.w, t.z, v.z) + c.z);
}
DEV v3 qrot_inv(qt q, v3 v) { return qrot(qconj(q), v); }
// Isometry product A*B
DEV se3 se3mul(se3 a, se3 b) {
  se3 r;
  r.t = add3(a.t, qrot(a.q, b.t));
  r.q = qmul(a.q, b.q);
  return r;
}
DEV double dot6(const double* a, const double* b) {
  return fma(a[0], b[0], fma(a[1], b[1], fma(a[2], b[2], fma(a[3], b[3], fma(a[4], b[4], a[5] * b[5])))));
}

// Polynomial coefficients live in constant memory: an fp64 instruction takes a constant-bank operand directly,
// whereas a 64-bit literal costs two extra move instructions every time it is used (a third of dsincos was moves).
static __constant__ double K_SC[20] = {
    6.36619772367581382433e-01,   // 0  2/pi
    6755399441055744.0,           // 1  1.5 * 2^52
    1.57079632673412561417e+00,   // 2  pi/2, first 33 bits
    6.07710050630396597660e-11,   // 3  next 33 bits
    2.02226624871116645580e-21,   // 4  remainder
    1.58969099521155010221e-10,   // 5  sin: S6
    -2.50507602534068634195e-08,  // 6  S5
    2.75573137070700676789e-06,   // 7  S4
    -1.98412698298579493134e-04,  // 8  S3
    8.33333333332248946124e-03,   // 9  S2
    -1.66666666666666324348e-01,  // 10 S1
    -1.13596475577881948265e-11,  // 11 cos: C6
    2.08757232129817482790e-09,   // 12 C5
    -2.75573143513906633035e-07,  // 13 C4
    2.48015872894767294178e-05,   // 14 C3
    -1.38888888888741095749e-03,  // 15 C2
    4.16666666666666019037e-02,   // 16 C1
    -0.5, 1.0, 0.0};
static __constant__ double K_AT[18] = {
    1.62858201153657823623e-02,   // 0  aT[10]
    4.97687799461593236017e-02,   // 1  aT[8]
    6.66107313738753120669e-02,   // 2  aT[6]
    9.09088713343650656196e-02,   // 3  aT[4]
    1.42857142725034663711e-01,   // 4  aT[2]
    3.33333333333329318027e-01,   // 5  aT[0]
    -3.65315727442169155270e-02,  // 6  aT[9]
    -5.83357013379057348645e-02,  // 7  aT[7]
    -7.69187620504482999495e-02,  // 8  aT[5]
    -1.11111104054623557880e-01,  // 9  aT[3]
    -1.99999999998764832476e-01,  // 10 aT[1]
    4.63647609000806093515e-01,   // 11 atan(1/2) hi
    2.26987774529616870924e-17,   // 12 atan(1/2) lo
    7.85398163397448278999e-01,   // 13 pi/4 hi
    3.06161699786838301793e-17,   // 14 pi/4 lo
    1.57079632679489655800e+00,   // 15 pi/2 hi
    6.12323399573676603587e-17,   // 16 pi/2 lo
    0.0};

// x with its sign bit XORed by `mask` (0 or 0x80000000): the same bits as a conditional negation, on the integer pipe
DEV double dflip(double x, unsigned mask) { return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x)); }

// sin/cos: Cody-Waite reduction by pi/2 + minimax polynomials on [-pi/4, pi/4].
DEV void dsincos(double x, double& sn, double& cs) {
  // k = rint(x * 2/pi) by the 1.5*2^52 shift (two fp64 adds instead of a round + a 64-bit convert, which run at a
  // quarter of the fp64 rate); the quadrant is the low two bits of the shifted value's mantissa.  Same result as
  // rint for |x * 2/pi| < 2^51; the CPU twin uses the identical sequence.
  const double kk = x * K_SC[0] + K_SC[1];
  const double k = kk - K_SC[1];
  double r = fma(-k, K_SC[2], x);
  r = fma(-k, K_SC[3], r);
  r = fma(-k, K_SC[4], r);
  const double z = r * r;
  double ps = fma(z, K_SC[5], K_SC[6]);
  ps = fma(z, ps, K_SC[7]);
  ps = fma(z, ps, K_SC[8]);
  ps = fma(z, ps, K_SC[9]);
  ps = fma(z, ps, K_SC[10]);
  const double s = fma(r * z, ps, r);
  double pc = fma(z, K_SC[11], K_SC[12]);
  pc = fma(z, pc, K_SC[13]);
  pc = fma(z, pc, K_SC[14]);
  pc = fma(z, pc, K_SC[15]);
  pc = fma(z, pc, K_SC[16]);
  const double c = fma(z * z, pc, fma(z, K_SC[17], K_SC[18]));
  const unsigned q = (unsigned)__double2loint(kk);
  const bool swap = (q & 1u) != 0u;
  const double ss = swap ? c : s;
  const double cc = swap ? s : c;
  sn = dflip(ss, (q & 2u) << 30);         // quadrants 2, 3: -sin
  cs = dflip(cc, ((q + 1u) & 2u) << 30);  // quadrants 1, 2: -cos
}

DEV double datan_poly(double x) {  // |x| <= 7/16
  const double z = x * x, w = z * z;
  double s1 = fma(w, K_AT[0], K_AT[1]);
  s1 = fma(w, s1, K_AT[2]);
  s1 = fma(w, s1, K_AT[3]);
  s1 = fma(w, s1, K_AT[4]);
  s1 = fma(w, s1, K_AT[5]);
  s1 = z * s1;
  double s2 = fma(w, K_AT[6], K_AT[7]);
  s2 = fma(w, s2, K_AT[8]);
  s2 = fma(w, s2, K_AT[9]);
  s2 = fma(w, s2, K_AT[10]);
  s2 = w * s2;
  return x - x * (s1 + s2);
}
// atan on [0,1], fdlibm breakpoints: t < 7/16: poly(t); t < 11/16: atan(1/2) + atan((2t-1)/(2+t)); else
// pi/4 + atan((t-1)/(t+1)).  Written branch-free (selected constants) so that threads with different arguments do
// not diverge; every select reproduces the branchy formulation bit for bit ((1*t-0)/(1+0*t) == t, 2*t == t+t,
// 0 + (p + 0) == p for p >= 0), which is what the CPU twin evaluates.
DEV double datan01(double t) {  // 0 <= t <= 1
  const bool b1 = t < 0.4375, b2 = t < 0.6875;
  const double a = (b1 || !b2) ? 1.0 : 2.0;   // numerator  a*t - b
  const double b = b1 ? 0.0 : 1.0;
  const double c = (b1 || !b2) ? 1.0 : 2.0;   // denominator c + d*t
  const double d = b1 ? 0.0 : 1.0;
  const double hi = b1 ? 0.0 : (b2 ? K_AT[11] : K_AT[13]);
  const double lo = b1 ? 0.0 : (b2 ? K_AT[12] : K_AT[14]);
  const double u = (a * t - b) / (c + d * t);
  return hi + (datan_poly(u) + lo);
}
// atan2(y, x) for y >= 0, x >= 0 (not both zero)
DEV double datan2_pos(double y, double x) {
  const bool small = y <= x;
  const double a = datan01(small ? y / x : x / y);
  return small ? a : K_AT[15] - (a - K_AT[16]);
}

// ---- ChaCha8 block (integer; restates rand_chacha's ChaCha8Rng state layout:
// 4 constants | 8 key words | 64-bit block counter | 64-bit stream id) --------
DEV uint32_t rotl32(uint32_t x, int k) { return __funnelshift_l(x, x, k); }
#define OPTIK_QR(a, b, c, d)                                              \
  a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12);   \
  a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);
// returns the (idx&7)-th u64 of block `counter` of stream `stream`
DEV uint64_t chacha8_u64(const uint32_t* key, uint64_t counter, uint64_t stream, int idx) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                    key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    OPTIK_QR(x[0], x[4], x[8], x[12]) OPTIK_QR(x[1], x[5], x[9], x[13])
    OPTIK_QR(x[2], x[6], x[10], x[14]) OPTIK_QR(x[3], x[7], x[11], x[15])
    OPTIK_QR(x[0], x[5], x[10], x[15]) OPTIK_QR(x[1], x[6], x[11], x[12])
    OPTIK_QR(x[2], x[7], x[8], x[13]) OPTIK_QR(x[3], x[4], x[9], x[14])
  }
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int k = 0; k < 8; k++)
    if ((idx & 7) == k) { lo = x[2 * k] + s[2 * k]; hi = x[2 * k + 1] + s[2 * k + 1]; }
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
// rand's f64 `random_range(lb..=ub)`: 52 mantissa bits -> [0,1), scale, shift (lib.rs:86-91)
DEV double uniform_f64(uint64_t u, double lb, double ub) {
  const double x01 = __longlong_as_double((long long)((u >> 12) | 0x3FF0000000000000ULL)) - 1.0;
  const double scale = (ub - lb) / (1.0 - 2.220446049250313e-16);
  const double v = x01 * scale + lb;
  return v > ub ? ub : v;
}

}  // namespace optik
