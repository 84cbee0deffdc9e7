// oracle/dojo_math.hpp -- TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// Fixed-size fp64 matrices and the quaternion algebra of the reference, restated literally so
// that the oracle reads like the Julia it follows:
//   /root/reference/src/orientation/quaternion.jl:13-223  (L, R, T, V, VL, VR, LV', ... and d/dq forms)
//   /root/reference/src/orientation/rotate.jl:2-39        (vector_rotate, rotation_matrix, d/dq)
//   /root/reference/src/orientation/mapping.jl:1-8        (quaternion_map, its Jacobian)
//   /root/reference/src/orientation/mrp.jl:1-80           (mrp, axis, angle, rotation_vector, d/dq)
//   Quaternions.jl (external, compat 0.5.2..0.7.6): Hamilton product, inv(q) = conj(q)/|q|^2.
// Parity status: see oracle/README.md ("parity unpinned": no Julia in the build image).
#pragma once
#include <cmath>
#include <cstring>

namespace dojo_oracle {

template <int R, int C>
struct Mat {
  double a[R * C];
  Mat() { std::memset(a, 0, sizeof(a)); }
  double& operator()(int i, int j) { return a[i * C + j]; }
  double operator()(int i, int j) const { return a[i * C + j]; }
  double& operator[](int i) { return a[i]; }
  double operator[](int i) const { return a[i]; }
  static Mat identity() {
    Mat m;
    for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
    return m;
  }
};
using V3 = Mat<3, 1>;
using V4 = Mat<4, 1>;
using V6 = Mat<6, 1>;
using M33 = Mat<3, 3>;
using M34 = Mat<3, 4>;
using M43 = Mat<4, 3>;
using M44 = Mat<4, 4>;
using M66 = Mat<6, 6>;

template <int R, int C>
Mat<R, C> operator+(const Mat<R, C>& x, const Mat<R, C>& y) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = x.a[i] + y.a[i];
  return z;
}
template <int R, int C>
Mat<R, C> operator-(const Mat<R, C>& x, const Mat<R, C>& y) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = x.a[i] - y.a[i];
  return z;
}
template <int R, int C>
Mat<R, C> operator-(const Mat<R, C>& x) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = -x.a[i];
  return z;
}
template <int R, int C>
Mat<R, C>& operator+=(Mat<R, C>& x, const Mat<R, C>& y) {
  for (int i = 0; i < R * C; ++i) x.a[i] += y.a[i];
  return x;
}
template <int R, int C>
Mat<R, C>& operator-=(Mat<R, C>& x, const Mat<R, C>& y) {
  for (int i = 0; i < R * C; ++i) x.a[i] -= y.a[i];
  return x;
}
template <int R, int C>
Mat<R, C> operator*(double s, const Mat<R, C>& x) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = s * x.a[i];
  return z;
}
template <int R, int C>
Mat<R, C> operator*(const Mat<R, C>& x, double s) { return s * x; }
template <int R, int C>
Mat<R, C> operator/(const Mat<R, C>& x, double s) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = x.a[i] / s;
  return z;
}
template <int R, int K, int C>
Mat<R, C> operator*(const Mat<R, K>& x, const Mat<K, C>& y) {
  Mat<R, C> z;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
      for (int k = 0; k < K; ++k) s += x(i, k) * y(k, j);
      z(i, j) = s;
    }
  return z;
}
template <int R, int C>
Mat<C, R> tr(const Mat<R, C>& x) {
  Mat<C, R> z;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) z(j, i) = x(i, j);
  return z;
}
template <int R>
double dot(const Mat<R, 1>& x, const Mat<R, 1>& y) {
  double s = 0.0;
  for (int i = 0; i < R; ++i) s += x[i] * y[i];
  return s;
}
template <int R>
double norm(const Mat<R, 1>& x) { return std::sqrt(dot(x, x)); }
template <int R, int C1, int C2>
Mat<R, C1 + C2> hcat(const Mat<R, C1>& x, const Mat<R, C2>& y) {
  Mat<R, C1 + C2> z;
  for (int i = 0; i < R; ++i) {
    for (int j = 0; j < C1; ++j) z(i, j) = x(i, j);
    for (int j = 0; j < C2; ++j) z(i, C1 + j) = y(i, j);
  }
  return z;
}
template <int R1, int R2, int C>
Mat<R1 + R2, C> vcat(const Mat<R1, C>& x, const Mat<R2, C>& y) {
  Mat<R1 + R2, C> z;
  for (int j = 0; j < C; ++j) {
    for (int i = 0; i < R1; ++i) z(i, j) = x(i, j);
    for (int i = 0; i < R2; ++i) z(R1 + i, j) = y(i, j);
  }
  return z;
}
template <int R0, int C0, int R, int C>
Mat<R0, C0> block(const Mat<R, C>& x, int r, int c) {
  Mat<R0, C0> z;
  for (int i = 0; i < R0; ++i)
    for (int j = 0; j < C0; ++j) z(i, j) = x(r + i, c + j);
  return z;
}
template <int R0, int C0, int R, int C>
void set_block(Mat<R, C>& x, int r, int c, const Mat<R0, C0>& y) {
  for (int i = 0; i < R0; ++i)
    for (int j = 0; j < C0; ++j) x(r + i, c + j) = y(i, j);
}
inline V3 vec3(double x, double y, double z) {
  V3 v;
  v[0] = x; v[1] = y; v[2] = z;
  return v;
}
inline V3 vec3(const double* p) { return vec3(p[0], p[1], p[2]); }
inline M33 mat33(const double* p) {
  M33 m;
  for (int i = 0; i < 9; ++i) m.a[i] = p[i];
  return m;
}

// ---------------------------------------------------------------- quaternion (scalar first)
struct Quat {
  double s, v1, v2, v3;
  Quat() : s(1), v1(0), v2(0), v3(0) {}
  Quat(double s_, double a, double b, double c) : s(s_), v1(a), v2(b), v3(c) {}
};
inline Quat qvec(const V3& v) { return Quat(0.0, v[0], v[1], v[2]); }  // Quaternion(v) quaternion.jl:1
inline V4 vector(const Quat& q) { V4 v; v[0] = q.s; v[1] = q.v1; v[2] = q.v2; v[3] = q.v3; return v; }
inline Quat quat(const V4& v) { return Quat(v[0], v[1], v[2], v[3]); }
inline Quat operator*(const Quat& a, const Quat& b) {  // Hamilton product (Quaternions.jl)
  return Quat(a.s * b.s - a.v1 * b.v1 - a.v2 * b.v2 - a.v3 * b.v3,
              a.s * b.v1 + a.v1 * b.s + a.v2 * b.v3 - a.v3 * b.v2,
              a.s * b.v2 - a.v1 * b.v3 + a.v2 * b.s + a.v3 * b.v1,
              a.s * b.v3 + a.v1 * b.v2 - a.v2 * b.v1 + a.v3 * b.s);
}
inline Quat operator*(const Quat& a, double k) { return Quat(a.s * k, a.v1 * k, a.v2 * k, a.v3 * k); }
inline Quat operator/(const Quat& a, double k) { return Quat(a.s / k, a.v1 / k, a.v2 / k, a.v3 / k); }
inline Quat inv(const Quat& q) {  // conj(q) / abs2(q)
  double n2 = q.s * q.s + q.v1 * q.v1 + q.v2 * q.v2 + q.v3 * q.v3;
  return Quat(q.s / n2, -q.v1 / n2, -q.v2 / n2, -q.v3 / n2);
}

inline M44 Lmat(const Quat& q) {  // quaternion.jl:16-23
  M44 m;
  const double r[16] = {q.s, -q.v1, -q.v2, -q.v3, q.v1, q.s, -q.v3, q.v2,
                        q.v2, q.v3, q.s, -q.v1, q.v3, -q.v2, q.v1, q.s};
  std::memcpy(m.a, r, sizeof(r));
  return m;
}
inline M44 Rmat(const Quat& q) {  // quaternion.jl:25-32
  M44 m;
  const double r[16] = {q.s, -q.v1, -q.v2, -q.v3, q.v1, q.s, q.v3, -q.v2,
                        q.v2, -q.v3, q.s, q.v1, q.v3, q.v2, -q.v1, q.s};
  std::memcpy(m.a, r, sizeof(r));
  return m;
}
inline M44 Ltmat(const Quat& q) { return tr(Lmat(q)); }
inline M44 Rtmat(const Quat& q) { return tr(Rmat(q)); }
inline M44 Tmat() {  // quaternion.jl:37-44
  M44 m;
  m(0, 0) = 1; m(1, 1) = -1; m(2, 2) = -1; m(3, 3) = -1;
  return m;
}
inline M34 Vmat() {  // quaternion.jl:46-52
  M34 m;
  m(0, 1) = 1; m(1, 2) = 1; m(2, 3) = 1;
  return m;
}
inline M43 Vtmat() { return tr(Vmat()); }
inline V3 Vmat(const Quat& q) { return vec3(q.v1, q.v2, q.v3); }
inline M34 VLmat(const Quat& q) { return Vmat() * Lmat(q); }      // quaternion.jl:65-71
inline M34 VLtmat(const Quat& q) { return Vmat() * Ltmat(q); }    // :73-79
inline M34 VRmat(const Quat& q) { return Vmat() * Rmat(q); }      // :81-87
inline M34 VRtmat(const Quat& q) { return Vmat() * Rtmat(q); }    // :89-95
inline M43 LVtmat(const Quat& q) { return Lmat(q) * Vtmat(); }    // :97-104
inline M43 LtVtmat(const Quat& q) { return Ltmat(q) * Vtmat(); }  // :106-113
inline M43 RVtmat(const Quat& q) { return Rmat(q) * Vtmat(); }    // :115-122
inline M43 RtVtmat(const Quat& q) { return Rtmat(q) * Vtmat(); }  // :124-131

// Matrix-vector product Jacobians, quaternion.jl:136-211 (constant sign patterns of p)
inline M44 dLVtmat_dq(const V3& p) {  // ∂LVᵀmat∂q :145-152  (4x4)
  M44 m; const double r[16] = {0, -p[0], -p[1], -p[2], p[0], 0, p[2], -p[1],
                               p[1], -p[2], 0, p[0], p[2], p[1], -p[0], 0};
  std::memcpy(m.a, r, sizeof(r));
  return m;
}
inline M34 dVLtmat_dq(const V4& p) {  // ∂VLᵀmat∂q :154-160
  M34 m; const double r[12] = {p[1], -p[0], -p[3], p[2], p[2], p[3], -p[0], -p[1], p[3], -p[2], p[1], -p[0]};
  std::memcpy(m.a, r, sizeof(r));
  return m;
}
inline M34 dVRtmat_dq(const V4& p) {  // ∂VRᵀmat∂q :187-193
  M34 m; const double r[12] = {p[1], -p[0], p[3], -p[2], p[2], -p[3], -p[0], p[1], p[3], p[2], -p[1], -p[0]};
  std::memcpy(m.a, r, sizeof(r));
  return m;
}
inline M44 dRtmat_dq(const V4& p) {  // ∂Rᵀmat∂q :195-202
  M44 m; const double r[16] = {p[0], p[1], p[2], p[3], p[1], -p[0], p[3], -p[2],
                               p[2], -p[3], -p[0], p[1], p[3], p[2], -p[1], -p[0]};
  std::memcpy(m.a, r, sizeof(r));
  return m;
}
inline M44 dLmat_dq(const V4& p) {  // ∂Lmat∂q :204-211
  M44 m; const double r[16] = {p[0], -p[1], -p[2], -p[3], p[1], p[0], p[3], -p[2],
                               p[2], -p[3], p[0], p[1], p[3], p[2], -p[1], p[0]};
  std::memcpy(m.a, r, sizeof(r));
  return m;
}
inline M33 skew(const V3& p) {  // :213-219
  M33 m; const double r[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
  std::memcpy(m.a, r, sizeof(r));
  return m;
}
inline M33 dskew_dp(const V3& l) { return skew(-l); }  // ∂skew∂p :221-223

// rotate.jl
inline V3 vector_rotate(const V3& v, const Quat& q) { return Vmat(q * qvec(v) * inv(q)); }  // :2-5
inline M34 dvector_rotate_dq(const V3& p, const Quat& q) {                                    // :6
  return VLmat(q) * Lmat(qvec(p)) * Tmat() + VRtmat(q) * Rmat(qvec(p));
}
inline M33 rotation_matrix(const Quat& q) { return VRtmat(q) * LVtmat(q); }  // :23
inline M34 drotation_matrix_dq(const Quat& q, const V3& p) {                 // :24-25
  return dVRtmat_dq(LVtmat(q) * p) + VRtmat(q) * dLVtmat_dq(p);
}
inline M34 drotation_matrix_inv_dq(const Quat& q, const V3& p) {  // :32-33
  return drotation_matrix_dq(inv(q), p) * Tmat();
}

// mapping.jl
inline Quat quaternion_map(const V3& w, double h) { return Quat(std::sqrt(4.0 / (h * h) - dot(w, w)), w[0], w[1], w[2]); }
inline M43 quaternion_map_jacobian(const V3& w, double h) {
  double msq = -std::sqrt(4.0 / (h * h) - dot(w, w));
  M43 m;
  for (int j = 0; j < 3; ++j) m(0, j) = w[j] / msq;
  m(1, 0) = 1; m(2, 1) = 1; m(3, 2) = 1;
  return m;
}

// mrp.jl
inline V3 mrp(const V4& q) { return vec3(q[1] / (q[0] + 1.0), q[2] / (q[0] + 1.0), q[3] / (q[0] + 1.0)); }
inline M34 dmrp_dq(const V4& q) {
  double s = q[0];
  double d1 = 1.0 / ((s + 1) * (s + 1)), di = 1.0 / (s + 1);
  M34 m;
  for (int i = 0; i < 3; ++i) { m(i, 0) = -q[1 + i] * d1; m(i, 1 + i) = di; }
  return m;
}
inline V3 axis(const V4& q) {
  V3 m = mrp(q);
  double mag = norm(m);
  if (mag > 0) return m / mag;
  return vec3(1.0, 0.0, 0.0);
}
inline double angle(const V4& q) {
  V3 m = mrp(q);
  double mag = norm(m);
  return mag > 0 ? 4.0 * std::atan(mag) : 0.0;
}
inline M34 daxis_dq(const V4& q) {
  V3 m = mrp(q);
  double n = norm(m);
  V3 mh = m / n;
  M34 D = dmrp_dq(q);
  return D / n - (m / (n * n)) * (tr(mh) * D);
}
inline Mat<1, 4> dangle_dq(const V4& q) {
  V3 m = mrp(q);
  double n = norm(m);
  return (4.0 / (1.0 + n * n)) * (tr(m / n) * dmrp_dq(q));
}
inline V3 rotation_vector(const Quat& q) { V4 v = vector(q); return angle(v) * axis(v); }
inline M34 drotation_vector_dq(const V4& q) {
  double th = angle(q);
  if (th != 0.0) return axis(q) * dangle_dq(q) + th * daxis_dq(q);
  M34 m;
  m(0, 1) = 2; m(1, 2) = 2; m(2, 3) = 2;
  return m;
}

}  // namespace dojo_oracle
