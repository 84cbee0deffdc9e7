// dojo_linalg.cuh -- half-warp-cooperative block linear algebra on the per-environment shared-memory arena.
//
// The factorisation is the block LDU of GraphBasedSystems.jl (external dependency of the reference; call sites
// src/solver/mehrotra.jl:36-37,49): nodes are eliminated leaves -> root, diagonal blocks are inverted explicitly
// (Gauss-Jordan, the whole block living in registers, one column per lane), off-diagonal blocks are updated with small
// GEMMs whose output elements are spread over the lanes.  Blocks are at most 6 x 6, so every routine works on a GROUP of
// 16 lanes (l = lane & 15, mask = the group's lanes): the two halves of a warp run two independent elimination steps
// at the same time with the same instruction stream.
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

// Gauss-Jordan inverse WITHOUT pivoting for the blocks of the condensed KKT system: body blocks (mass matrix plus
// contact / limit / damper terms) and joint blocks (REG I + U D_b^-1 G, a J M^-1 J' form) have safely non-zero
// diagonals in elimination order; the blocks that needed pivoting (contact and limit complementarity rows) are
// condensed out analytically.  A vanishing / non-finite pivot is reported (status 3), never silently used.
// In-place on an n x n row-major block (leading dimension ld) in shared memory: lanes 0..n-1 of the group own the
// columns of A, lanes n..2n-1 the columns of the identity that becomes A^{-1}.
template <int N>
DJ_LA bool block_inverse_nopivot_t(double* A, int ld, int l, unsigned mask) {
  static_assert(2 * N <= 16, "block too large for the one-column-per-lane inverse");
  double a[N];
  const bool left = l < N;
  const int col = left ? l : l - N;
  const bool active = l < 2 * N;
#pragma unroll
  for (int r = 0; r < N; ++r) a[r] = active ? (left ? A[r * ld + col] : (r == col ? 1.0 : 0.0)) : 0.0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
#ifdef DJ_BLOCK_PIVOT  // opt-in build variant; measured (128 ant environments in hard contact, rtol = btol = 1e-10 / 1e-12): no gain
    // in convergence over the unpivoted elimination once the joint limits are no longer condensed through 1 / s (dojo_plan.h joint_nq)
    // partial (row) pivoting inside the block: the lane that owns column k picks the largest remaining entry of its column, every
    // lane of the group swaps the two rows of its own column.  Row operations on [A | I] leave A^-1 in the right half as before.
    if (k + 1 < N) {
      int p = k;
      double best = fabs(a[k]);
#pragma unroll
      for (int r = k + 1; r < N; ++r) {
        const double v = fabs(a[r]);
        if (v > best) { best = v; p = r; }
      }
      p = __shfl_sync(mask, p, k, 16);
#pragma unroll
      for (int r = k + 1; r < N; ++r)
        if (r == p) { const double t = a[k]; a[k] = a[r]; a[r] = t; }
    }
#endif
    double piv = __shfl_sync(mask, a[k], k, 16);
    if (!(fabs(piv) > 1e-300) || !(fabs(piv) < 1e300)) ok = false;
    double akk = a[k] * (1.0 / piv);
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (r == k) continue;
      double f = __shfl_sync(mask, a[r], k, 16);
      a[r] -= f * akk;
    }
    a[k] = akk;
  }
  __syncwarp(mask);
  if (active && !left) {
#pragma unroll
    for (int r = 0; r < N; ++r) A[r * ld + col] = a[r];
  }
  __syncwarp(mask);
  return ok;
}

DJ_DEV bool block_inverse(double* A, int n, int ld, int l, unsigned mask) {
  switch (n) {
    case 1: return block_inverse_nopivot_t<1>(A, ld, l, mask);
    case 2: return block_inverse_nopivot_t<2>(A, ld, l, mask);
    case 3: return block_inverse_nopivot_t<3>(A, ld, l, mask);
    case 4: return block_inverse_nopivot_t<4>(A, ld, l, mask);
    case 5: return block_inverse_nopivot_t<5>(A, ld, l, mask);
    case 6: return block_inverse_nopivot_t<6>(A, ld, l, mask);
    default: return false;  // body blocks are 6 x 6, joint equality blocks at most 6 x 6
  }
}

// ---------------------------------------------------------------------------------------------------------
// Generic (runtime-size) fallbacks, m, n <= 6
// ---------------------------------------------------------------------------------------------------------
// L (m x n, row-major) <- L * Dinv (n x n), in place.  Output elements are spread over the lanes; everything is
// computed into registers before anything is written back (a row of L is both input and output).
__device__ __noinline__ void right_multiply_generic(double* L, const double* Dinv, int m, int n, int l, unsigned mask) {
  const int total = m * n;  // <= 36
  double out[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    int e = l + 16 * p;
    double acc = 0.0;
    if (e < total) {
      int i = e / n, j = e - i * n;
      for (int k = 0; k < n; ++k) acc += L[i * n + k] * Dinv[k * n + j];
    }
    out[p] = acc;
  }
  __syncwarp(mask);
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    int e = l + 16 * p;
    if (e < total) L[e] = out[p];
  }
  __syncwarp(mask);
}

// C (ni x nj, ld = ldc) -= A (ni x k, ld = lda) * B (k x nj, ld = nj)
__device__ __noinline__ void schur_generic(double* C, int ldc, const double* A, int lda, const double* B, int ni, int k, int nj, int l, unsigned mask) {
  const int total = ni * nj;
  for (int e = l; e < total; e += 16) {
    int i = e / nj, j = e - i * nj;
    double acc = 0.0;
    for (int t = 0; t < k; ++t) acc += A[i * lda + t] * B[t * nj + j];
    C[i * ldc + j] -= acc;
  }
  __syncwarp(mask);
}

// ---------------------------------------------------------------------------------------------------------
// Fixed-size kernels: fully unrolled, immediate shared-memory offsets, each lane owns a 1 x 3 strip of the output
// (one load of A feeds three FMAs with independent accumulators).
// ---------------------------------------------------------------------------------------------------------
constexpr int kStrip = 3;

template <int M, int N>
DJ_LA void right_multiply_t(double* L, const double* Dinv, int l, unsigned mask) {
  constexpr int SPR = (N + kStrip - 1) / kStrip;
  constexpr int NS = M * SPR;
  static_assert(NS <= 16, "single pass only");
  const int i = l / SPR, js = (l - i * SPR) * kStrip;
  double acc[kStrip] = {0.0, 0.0, 0.0};
  if (l < NS) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double a = L[i * N + k];
#pragma unroll
      for (int c = 0; c < kStrip; ++c)
        if (js + c < N) acc[c] += a * Dinv[k * N + js + c];
    }
  }
  __syncwarp(mask);
  if (l < NS) {
#pragma unroll
    for (int c = 0; c < kStrip; ++c)
      if (js + c < N) L[i * N + js + c] = acc[c];
  }
  __syncwarp(mask);
}

template <int NI, int K, int NJ, int LDC = NJ>
DJ_LA void schur_t(double* C, const double* A, int lda, const double* B, int l, unsigned mask) {
  constexpr int SPR = (NJ + kStrip - 1) / kStrip;
  constexpr int NS = NI * SPR;
  static_assert(NS <= 16, "single pass only");
  if (l < NS) {
    const int i = l / SPR, js = (l - i * SPR) * kStrip;
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
      if (js + c < NJ) C[i * LDC + js + c] -= acc[c];
  }
  __syncwarp(mask);
}

#define DJ_RM_CASE(M, N) case (M) * 32 + (N): right_multiply_t<M, N>(L, Dinv, l, mask); return;
DJ_DEV void right_multiply_inplace(double* L, const double* Dinv, int m, int n, int l, unsigned mask) {
  switch (m * 32 + n) {
    DJ_RM_CASE(6, 6) DJ_RM_CASE(5, 6) DJ_RM_CASE(6, 5) DJ_RM_CASE(3, 6) DJ_RM_CASE(6, 3)
    default: right_multiply_generic(L, Dinv, m, n, l, mask);
  }
}
// ldc = nj (a whole block) or 6 (three columns of a block that belongs to a body: the angular part, dojo_plan.h ElimNb::ld)
#define DJ_SC_CASE(NI, K, NJ) case (((NI) * 32 + (K)) * 32 + (NJ)) * 8 + (NJ): schur_t<NI, K, NJ>(C, A, lda, B, l, mask); return;
#define DJ_SC_CASE_LD6(NI, K, NJ) case (((NI) * 32 + (K)) * 32 + (NJ)) * 8 + 6: schur_t<NI, K, NJ, 6>(C, A, lda, B, l, mask); return;
DJ_DEV void schur_update(double* C, int ldc, const double* A, int lda, const double* B, int ni, int k, int nj, int l, unsigned mask) {
  switch (((ni * 32 + k) * 32 + nj) * 8 + ldc) {
    DJ_SC_CASE(6, 6, 6)
    DJ_SC_CASE(5, 6, 5) DJ_SC_CASE(5, 6, 6) DJ_SC_CASE(6, 6, 5) DJ_SC_CASE(6, 5, 6)
    DJ_SC_CASE(3, 6, 3) DJ_SC_CASE(3, 6, 6) DJ_SC_CASE(6, 6, 3) DJ_SC_CASE(6, 3, 6)
    // body-body coupling through the angular rows (limits / dampers): (joint, parent) columns 3..5, (parent, joint) rows 3..5, parent diagonal
    DJ_SC_CASE_LD6(6, 3, 3) DJ_SC_CASE_LD6(5, 3, 3) DJ_SC_CASE_LD6(3, 3, 3) DJ_SC_CASE(3, 6, 5)
    default: schur_generic(C, ldc, A, lda, B, ni, k, nj, l, mask);
  }
}

}  // namespace dj
