// dojo_math.cuh -- device-side fp64 small-vector / quaternion algebra for the Dojo step kernels.
//
// Everything here lives in registers: fixed-size structs, fully unrolled loops, static indexing only
// (a dynamically indexed register array would be demoted to local memory).  Quaternions are Hamilton,
// scalar-first, as in the reference (src/orientation/quaternion.jl:13-32).  Instead of the reference's
// 4x4 / 3x4 matrix forms (L, R, T, V ...) the kernels use closed forms written directly in terms of
// rotation matrices and the *attitude* (body-frame) perturbation q (x) (1, d): see DESIGN.md §"Closed forms".
#pragma once
#include <math.h>
#include <string.h>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define DJ_DEV __device__ __forceinline__
#else
// host compilation of the lane-independent helpers: tests/hostcheck builds the per-environment kinematics code
// (dojo_kin.cuh, dojo_kinjac.cuh) with g++ so that its arithmetic can be checked on a machine without a GPU.  The product
// library is always built by nvcc; nothing in it runs on the CPU.
#define DJ_DEV inline
#endif

namespace dj {

struct V3 {
  double x, y, z;
};
struct Quat {
  double s, x, y, z;
};
struct M33 {
  double m[3][3];
};

DJ_DEV V3 v3(double x, double y, double z) { return V3{x, y, z}; }
DJ_DEV V3 v3zero() { return V3{0.0, 0.0, 0.0}; }
DJ_DEV V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
DJ_DEV V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
DJ_DEV V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
DJ_DEV V3 operator*(double k, V3 a) { return V3{k * a.x, k * a.y, k * a.z}; }
DJ_DEV V3 operator*(V3 a, double k) { return V3{k * a.x, k * a.y, k * a.z}; }
DJ_DEV V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
DJ_DEV V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
DJ_DEV double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DJ_DEV V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
DJ_DEV double comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }  // i must be a compile-time constant after unrolling
DJ_DEV double maxabs(V3 a) { return fmax(fabs(a.x), fmax(fabs(a.y), fabs(a.z))); }

DJ_DEV M33 m33zero() {
  M33 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = 0.0;
  return r;
}
DJ_DEV M33 m33ident(double k = 1.0) {
  M33 r = m33zero();
  r.m[0][0] = k; r.m[1][1] = k; r.m[2][2] = k;
  return r;
}
DJ_DEV M33 skew(V3 p) {  // skew(p) q = p x q
  M33 r;
  r.m[0][0] = 0.0; r.m[0][1] = -p.z; r.m[0][2] = p.y;
  r.m[1][0] = p.z; r.m[1][1] = 0.0; r.m[1][2] = -p.x;
  r.m[2][0] = -p.y; r.m[2][1] = p.x; r.m[2][2] = 0.0;
  return r;
}
DJ_DEV M33 outer(V3 a, V3 b) {
  M33 r;
  r.m[0][0] = a.x * b.x; r.m[0][1] = a.x * b.y; r.m[0][2] = a.x * b.z;
  r.m[1][0] = a.y * b.x; r.m[1][1] = a.y * b.y; r.m[1][2] = a.y * b.z;
  r.m[2][0] = a.z * b.x; r.m[2][1] = a.z * b.y; r.m[2][2] = a.z * b.z;
  return r;
}
DJ_DEV M33 operator+(const M33& a, const M33& b) {
  M33 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
DJ_DEV M33 operator-(const M33& a, const M33& b) {
  M33 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r;
}
DJ_DEV M33 operator*(double k, const M33& a) {
  M33 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = k * a.m[i][j];
  return r;
}
DJ_DEV M33 operator*(const M33& a, const M33& b) {
  M33 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
DJ_DEV V3 operator*(const M33& a, V3 v) {
  return V3{a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
DJ_DEV M33 transpose(const M33& a) {
  M33 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
DJ_DEV V3 tmul(const M33& a, V3 v) {  // a' * v
  return V3{a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z, a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z,
            a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z};
}
DJ_DEV V3 row(const M33& a, int i) { return V3{a.m[i][0], a.m[i][1], a.m[i][2]}; }
DJ_DEV V3 vtmul(V3 v, const M33& a) { return tmul(a, v); }  // (v' * a)' = a' v

// ----------------------------------------------------------------------------------------- quaternions
DJ_DEV Quat qmul(Quat a, Quat b) {
  return Quat{a.s * b.s - a.x * b.x - a.y * b.y - a.z * b.z, a.s * b.x + a.x * b.s + a.y * b.z - a.z * b.y,
              a.s * b.y - a.x * b.z + a.y * b.s + a.z * b.x, a.s * b.z + a.x * b.y - a.y * b.x + a.z * b.s};
}
DJ_DEV Quat qconj(Quat q) { return Quat{q.s, -q.x, -q.y, -q.z}; }
DJ_DEV Quat qinv(Quat q) {  // conj(q) / |q|^2, as Quaternions.jl
  double n2 = q.s * q.s + q.x * q.x + q.y * q.y + q.z * q.z;
  double k = 1.0 / n2;
  return Quat{q.s * k, -q.x * k, -q.y * k, -q.z * k};
}
DJ_DEV V3 qvec(Quat q) { return V3{q.x, q.y, q.z}; }
// rotation_matrix(q) = VR'(q) LV'(q)  (src/orientation/rotate.jl:23); equals the SO(3) matrix for unit q
DJ_DEV M33 rotmat(Quat q) {
  M33 r;
  double ss = q.s * q.s, xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  r.m[0][0] = ss + xx - yy - zz; r.m[0][1] = 2.0 * (q.x * q.y - q.s * q.z); r.m[0][2] = 2.0 * (q.x * q.z + q.s * q.y);
  r.m[1][0] = 2.0 * (q.x * q.y + q.s * q.z); r.m[1][1] = ss - xx + yy - zz; r.m[1][2] = 2.0 * (q.y * q.z - q.s * q.x);
  r.m[2][0] = 2.0 * (q.x * q.z - q.s * q.y); r.m[2][1] = 2.0 * (q.y * q.z + q.s * q.x); r.m[2][2] = ss - xx - yy + zz;
  return r;
}

// Integrator: q3 = q2 (x) m(w),  m(w) = (sqrt(4/h^2 - w.w), w) h/2 = (m0, (h/2) w)   (integrators/integrator.jl:15, mapping.jl:1-3)
DJ_DEV Quat qmap(V3 w, double h) {
  double m0 = 0.5 * h * sqrt(4.0 / (h * h) - dot(w, w));
  return Quat{m0, 0.5 * h * w.x, 0.5 * h * w.y, 0.5 * h * w.z};
}
// E(w): d q3 = LV'(q3) E dw, i.e. the body-frame attitude perturbation of q3 per unit change of w25.
// E = (h/2) [ m0 I + (h^2/4) w w' / m0 - (h/2) skew(w) ]      (closed form of V L(m)' dm/dw)
DJ_DEV M33 attitude_velocity_jacobian(V3 w, double h) {
  double m0 = 0.5 * h * sqrt(4.0 / (h * h) - dot(w, w));
  M33 r = m33ident(m0) + ((0.25 * h * h) / m0) * outer(w, w) - (0.5 * h) * skew(w);
  return (0.5 * h) * r;
}

// rotation_vector(q) = 4 atan(|mrp|) mrp/|mrp|, mrp = v/(1+s)   (src/orientation/mrp.jl:1-64)
DJ_DEV V3 rotation_vector(Quat q) {
  double k = 1.0 / (q.s + 1.0);
  V3 m = v3(q.x * k, q.y * k, q.z * k);
  double mag = sqrt(dot(m, m));
  if (mag > 0.0) return (4.0 * atan(mag) / mag) * m;
  return v3zero();
}
// d rotation_vector / dq (3x4) as 3 rows of (ds, dv) pairs          (src/orientation/mrp.jl:10-80)
struct M34 {
  double m[3][4];
};
DJ_DEV M34 drotation_vector_dq(Quat q) {
  M34 r;
  double k = 1.0 / (q.s + 1.0);
  V3 m = v3(q.x * k, q.y * k, q.z * k);
  double n2 = dot(m, m);
  double n = sqrt(n2);
  if (!(n > 0.0)) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) r.m[i][j] = (j == i + 1) ? 2.0 : 0.0;
    return r;
  }
  // dmrp/dq = [ -v/(1+s)^2 , I/(1+s) ]
  double th = 4.0 * atan(n);
  // rv = th * m / n ;  d rv = (m/n) dth + th d(m/n),  dth = 4/(1+n^2) (m/n)' dm,  d(m/n) = (I/n - m m'/n^3) dm
  // => drv/dm = (4/(1+n^2)) u u' + (th/n) (I - u u'),  u = m/n
  V3 u = (1.0 / n) * m;
  double a = 4.0 / (1.0 + n2), b = th / n;
  M33 dm = (a - b) * outer(u, u) + m33ident(b);
  V3 dms = (-k) * m;  // dm/ds = -v/(1+s)^2 = -m/(1+s)
  V3 c0 = dm * dms;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r.m[i][0] = comp(c0, i);
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i][1 + j] = dm.m[i][j] * k;
  }
  return r;
}

#if defined(__CUDACC__) || defined(DJ_HOSTEMU)
// warp reductions
DJ_DEV double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
DJ_DEV double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
DJ_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// max that propagates NaN like Julia's max (so a non-finite iterate is detected, not silently dropped).  Every caller passes
// magnitudes (|x|, or the result of an earlier nanmax starting from +0): for non-negative doubles the IEEE-754 bit patterns are
// ordered like the values, and a NaN (sign cleared by fabs) has a larger pattern than +Inf -- so the 64-bit INTEGER maximum of the
// two patterns is the NaN-propagating maximum: four integer instructions instead of two DSETP + DADD + the fp64 min/max emulation
// (the violations are reduced with this in every residual evaluation; 4 % of the forward kernel's instructions before).
DJ_DEV double nanmax(double a, double b) {
#ifdef __CUDA_ARCH__
  const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
  return __longlong_as_double(ia > ib ? ia : ib);
#else
  long long ia, ib;
  memcpy(&ia, &a, 8); memcpy(&ib, &b, 8);
  const long long im = ia > ib ? ia : ib;
  double r; memcpy(&r, &im, 8);
  return r;
#endif
}
DJ_DEV double warp_nanmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

}  // namespace dj
