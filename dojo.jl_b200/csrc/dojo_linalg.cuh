// dojo_linalg.cuh -- warp-cooperative block linear algebra on the per-environment shared-memory arena.
//
// One warp owns one environment's block-sparse KKT system.  The factorisation is the block LDU of
// GraphBasedSystems.jl (external dependency of the reference; call sites src/solver/mehrotra.jl:36-37,49):
// nodes are eliminated leaves -> root, diagonal blocks are inverted explicitly (Gauss-Jordan with partial
// pivoting inside the block, the whole block living in registers, one column per lane), off-diagonal
// blocks are updated with warp-wide small GEMMs whose output elements are spread over the lanes.
//
//   for c in elimination order, N(c) = later-eliminated neighbours (<= 2 for tree mechanisms):
//       Dinv_c  = inv(D_c)                                  (in place)
//       L~_ic   = M_ic * Dinv_c            for i in N(c)     (in place; the column blocks of c)
//       M_ij   -= L~_ic * M_cj             for i, j in N(c)  (the row blocks M_cj of c are left untouched)
//   solve:  forward  z_i -= L~_ic z_c      backward  x_c = Dinv_c (z_c - sum_j M_cj x_j)
#pragma once
#include "dojo_math.cuh"

#ifndef DJ_NOINLINE_LA
#define DJ_LA __device__ __forceinline__
#else
#define DJ_LA __device__ __noinline__
#endif

namespace dj {

// In-place inverse of an n x n row-major block (leading dimension ld) held in shared memory.
// Lanes 0..n-1 own the columns of A, lanes n..2n-1 the columns of the identity that becomes A^{-1}.
// Returns false (warp-uniform) if a zero / non-finite pivot is met.
template <int N>
DJ_LA bool block_inverse_t(double* A, int ld, int lane) {
  static_assert(2 * N <= 32, "block too large for the one-column-per-lane inverse");
  double a[N];
  const bool left = lane < N;
  const int col = left ? lane : lane - N;
  const bool active = lane < 2 * N;
#pragma unroll
  for (int r = 0; r < N; ++r) a[r] = active ? (left ? A[r * ld + col] : (r == col ? 1.0 : 0.0)) : 0.0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    // pivot search in column k (owned by lane k)
    int p = k;
    double best = fabs(a[k]);
#pragma unroll
    for (int r = k + 1; r < N; ++r) {
      double v = fabs(a[r]);
      if (v > best) { best = v; p = r; }
    }
    p = __shfl_sync(0xffffffffu, p, k);
    best = __shfl_sync(0xffffffffu, best, k);
    if (!(best > 0.0) || !(best < 1e300)) ok = false;
    // swap rows k and p (predicated, static indices)
    double akp = a[k];
#pragma unroll
    for (int r = k + 1; r < N; ++r)
      if (r == p) { double t = a[r]; a[r] = akp; akp = t; }
    a[k] = akp;
    double piv = __shfl_sync(0xffffffffu, a[k], k);
    double inv = 1.0 / piv;
    a[k] *= inv;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (r == k) continue;
      double f = __shfl_sync(0xffffffffu, a[r], k);
      a[r] -= f * a[k];
    }
  }
  __syncwarp();
  if (active && !left) {
#pragma unroll
    for (int r = 0; r < N; ++r) A[r * ld + col] = a[r];
  }
  __syncwarp();
  return ok;
}

// Gauss-Jordan inverse WITHOUT pivoting for the blocks of the condensed KKT system: body blocks (mass matrix plus
// contact / limit / damper terms) and joint blocks (REG I + U D_b^-1 G, a J M^-1 J' form) have safely non-zero
// diagonals in elimination order; the blocks that needed pivoting (contact and limit complementarity rows) are
// condensed out analytically.  A vanishing / non-finite pivot is reported (status 3), never silently used.
template <int N>
DJ_LA bool block_inverse_nopivot_t(double* A, int ld, int lane) {
  static_assert(2 * N <= 32, "block too large for the one-column-per-lane inverse");
  double a[N];
  const bool left = lane < N;
  const int col = left ? lane : lane - N;
  const bool active = lane < 2 * N;
#pragma unroll
  for (int r = 0; r < N; ++r) a[r] = active ? (left ? A[r * ld + col] : (r == col ? 1.0 : 0.0)) : 0.0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double piv = __shfl_sync(0xffffffffu, a[k], k);
    if (!(fabs(piv) > 1e-300) || !(fabs(piv) < 1e300)) ok = false;
    double akk = a[k] * (1.0 / piv);
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (r == k) continue;
      double f = __shfl_sync(0xffffffffu, a[r], k);
      a[r] -= f * akk;
    }
    a[k] = akk;
  }
  __syncwarp();
  if (active && !left) {
#pragma unroll
    for (int r = 0; r < N; ++r) A[r * ld + col] = a[r];
  }
  __syncwarp();
  return ok;
}

// n <= 16 fallback: embed the block into a 16 x 16 matrix padded with the identity
DJ_DEV bool block_inverse_pad16(double* A, int n, int ld, int lane) {
  double a[16];
  const bool left = lane < 16;
  const int col = lane & 15;
#pragma unroll
  for (int r = 0; r < 16; ++r) a[r] = left ? ((r < n && col < n) ? A[r * ld + col] : (r == col ? 1.0 : 0.0)) : (r == col ? 1.0 : 0.0);
  bool ok = true;
#pragma unroll 1
  for (int k = 0; k < 16; ++k) {
    int p = 0;
    double best = -1.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double v = fabs(a[r]);
      if (r >= k && v > best) { best = v; p = r; }
    }
    p = __shfl_sync(0xffffffffu, p, k);
    best = __shfl_sync(0xffffffffu, best, k);
    if (!(best > 0.0) || !(best < 1e300)) ok = false;
    double ak = 0.0, ap = 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) { if (r == k) ak = a[r]; if (r == p) ap = a[r]; }
#pragma unroll
    for (int r = 0; r < 16; ++r) { if (r == k) a[r] = ap; else if (r == p) a[r] = ak; }
    double piv = __shfl_sync(0xffffffffu, ap, k);
    double akk = ap * (1.0 / piv);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double f = __shfl_sync(0xffffffffu, a[r], k);
      a[r] = (r == k) ? akk : a[r] - f * akk;
    }
  }
  __syncwarp();
  if (!left && col < n) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r < n) A[r * ld + col] = a[r];
  }
  __syncwarp();
  return ok;
}

// Runtime-size fallback: the augmented matrix lives in registers of a 16-column template; only used for block
// sizes outside the specialised set below.
DJ_DEV bool block_inverse(double* A, int n, int ld, int lane) {
#ifdef DJ_PIVOT
  switch (n) {
    case 1: return block_inverse_t<1>(A, ld, lane);
    case 2: return block_inverse_t<2>(A, ld, lane);
    case 3: return block_inverse_t<3>(A, ld, lane);
    case 4: return block_inverse_t<4>(A, ld, lane);
    case 5: return block_inverse_t<5>(A, ld, lane);
    case 6: return block_inverse_t<6>(A, ld, lane);
    default: return false;
  }
#endif
  switch (n) {
    case 1: return block_inverse_nopivot_t<1>(A, ld, lane);
    case 2: return block_inverse_nopivot_t<2>(A, ld, lane);
    case 3: return block_inverse_nopivot_t<3>(A, ld, lane);
    case 4: return block_inverse_nopivot_t<4>(A, ld, lane);
    case 5: return block_inverse_nopivot_t<5>(A, ld, lane);
    case 6: return block_inverse_nopivot_t<6>(A, ld, lane);
    default: return false;  // joint equality blocks are at most 6 x 6
  }
}

// ---------------------------------------------------------------------------------------------------------
// Generic (runtime-size) fallbacks
// ---------------------------------------------------------------------------------------------------------
// L (m x n, row-major) <- L * Dinv (n x n), in place.  Output elements are spread over the lanes; everything is
// computed into registers before anything is written back (a row of L is both input and output).
__device__ __noinline__ void right_multiply_generic(double* L, const double* Dinv, int m, int n, int lane) {
  const int total = m * n;  // <= 256
  double out[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    int e = lane + 32 * p;
    double acc = 0.0;
    if (e < total) {
      int i = e / n, j = e - i * n;
      for (int k = 0; k < n; ++k) acc += L[i * n + k] * Dinv[k * n + j];
    }
    out[p] = acc;
  }
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    int e = lane + 32 * p;
    if (e < total) L[e] = out[p];
  }
  __syncwarp();
}

// C (ni x nj, ld = nj) -= A (ni x k, ld = lda) * B (k x nj, ld = nj)
__device__ __noinline__ void schur_generic(double* C, const double* A, int lda, const double* B, int ni, int k, int nj, int lane) {
  const int total = ni * nj;
  for (int e = lane; e < total; e += 32) {
    int i = e / nj, j = e - i * nj;
    double acc = 0.0;
    for (int t = 0; t < k; ++t) acc += A[i * lda + t] * B[t * nj + j];
    C[e] -= acc;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------
// Fixed-size kernels: fully unrolled, immediate shared-memory offsets, each lane owns a 1 x 3 strip of the output
// (one load of A feeds three FMAs with independent accumulators).
// ---------------------------------------------------------------------------------------------------------
constexpr int kStrip = 3;

template <int M, int N>
DJ_LA void right_multiply_t(double* L, const double* Dinv, int lane) {
  constexpr int SPR = (N + kStrip - 1) / kStrip;
  constexpr int NS = M * SPR;
  static_assert(NS <= 32, "single pass only");
  const int i = lane / SPR, js = (lane - i * SPR) * kStrip;
  double acc[kStrip] = {0.0, 0.0, 0.0};
  if (lane < NS) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double a = L[i * N + k];
#pragma unroll
      for (int c = 0; c < kStrip; ++c)
        if (js + c < N) acc[c] += a * Dinv[k * N + js + c];
    }
  }
  __syncwarp();
  if (lane < NS) {
#pragma unroll
    for (int c = 0; c < kStrip; ++c)
      if (js + c < N) L[i * N + js + c] = acc[c];
  }
  __syncwarp();
}

template <int NI, int K, int NJ>
DJ_LA void schur_t(double* C, const double* A, int lda, const double* B, int lane) {
  constexpr int SPR = (NJ + kStrip - 1) / kStrip;
  constexpr int NS = NI * SPR;
  static_assert(NS <= 32, "single pass only");
  if (lane < NS) {
    const int i = lane / SPR, js = (lane - i * SPR) * kStrip;
    double acc[kStrip] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < K; ++t) {
      double a = A[i * lda + t];
#pragma unroll
      for (int c = 0; c < kStrip; ++c)
        if (js + c < NJ) acc[c] += a * B[t * NJ + js + c];
    }
#pragma unroll
    for (int c = 0; c < kStrip; ++c)
      if (js + c < NJ) C[i * NJ + js + c] -= acc[c];
  }
  __syncwarp();
}

#define DJ_RM_CASE(M, N) case (M) * 32 + (N): right_multiply_t<M, N>(L, Dinv, lane); return;
DJ_DEV void right_multiply_inplace(double* L, const double* Dinv, int m, int n, int lane) {
  switch (m * 32 + n) {
    DJ_RM_CASE(6, 8) DJ_RM_CASE(6, 6) DJ_RM_CASE(5, 6) DJ_RM_CASE(9, 6) DJ_RM_CASE(6, 5) DJ_RM_CASE(6, 9)
    DJ_RM_CASE(3, 6) DJ_RM_CASE(6, 3)
    default: right_multiply_generic(L, Dinv, m, n, lane);
  }
}
#define DJ_SC_CASE(NI, K, NJ) case ((NI) * 32 + (K)) * 32 + (NJ): schur_t<NI, K, NJ>(C, A, lda, B, lane); return;
DJ_DEV void schur_update(double* C, const double* A, int lda, const double* B, int ni, int k, int nj, int lane) {
  switch ((ni * 32 + k) * 32 + nj) {
    DJ_SC_CASE(6, 4, 6)
    DJ_SC_CASE(6, 6, 6)
    DJ_SC_CASE(5, 6, 5) DJ_SC_CASE(5, 6, 6) DJ_SC_CASE(6, 6, 5) DJ_SC_CASE(6, 5, 6)
    DJ_SC_CASE(9, 6, 9) DJ_SC_CASE(9, 6, 6) DJ_SC_CASE(6, 6, 9) DJ_SC_CASE(6, 9, 6)
    DJ_SC_CASE(3, 6, 3) DJ_SC_CASE(3, 6, 6) DJ_SC_CASE(6, 6, 3) DJ_SC_CASE(6, 3, 6)
    default: schur_generic(C, A, lda, B, ni, k, nj, lane);
  }
}

}  // namespace dj
