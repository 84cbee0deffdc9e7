// dojo_plan.h -- flattened, device-resident description of one mechanism ("plan").
// Built once on the host in dojo_create() from the DojoMechanismDesc (include/dojo_b200.h), read-only on
// the device.  All *_off fields are offsets (in doubles) into the per-environment shared-memory arena.
#pragma once
#include <stdint.h>

namespace dj {

constexpr double kReg = 1.0e-10;  // REG, src/Dojo.jl:4

struct BodyDev {
  double mass;
  double J[9];
  int sol_off;  // [v25(3); w25(3)] inside the solution vector
  int st_off;   // x2(3), q2(4)
  int cst_off;  // constant part of the dynamics residual for this step (6)
  int D_off;    // 6x6 diagonal block
};

struct JointDev {
  int parent, child;  // body indices, parent = -1 for the origin
  int n, sol_off;     // impulse dimension / offset inside the solution vector
  int nl_t, nl_r, nb2_r, nb_r;  // constrained axes (tra, rot), rotational limits Nb/2 and Nb
  int row_r;          // first row of the rotational element inside the joint vector (= nl_t)
  int nfree_t, nfree_r, u_off;
  double pa[3], pb[3], qoff[4];
  double Ct[9], At[9], Cr[9], Ar[9];  // constraint / nullspace masks, zero-padded to 3 rows (joints/joint.jl:56-64)
  double spring_r, damper_r, spring_off_r[3], lo[3], hi[3];
  int D_off;                    // n x n
  int Uc_off, Lc_off;           // (joint,child) n x 6, rewritten every assembly ; (child,joint) 6 x n = -G_c, constant over the
                                // solve and never written by the factorisation (lives in the constant region of the arena)
  int Up_off, Lp_off, Gp_off;   // same for the parent body (-1 when the parent is the origin); Lp is consumed by the
                                // factorisation and refreshed from the pristine impulse map Gp at every assembly
  int BBpc_off, BBcp_off;       // (parent,child) / (child,parent) 6x6 blocks, -1 without dampers
  int color_parent;             // index of this joint among its parent's child joints
};

struct ContactDev {
  int body, sol_off;  // [s(4); gamma(4)] inside the solution vector
  double mu, radius;
  double n[3], t[6], o[3], off[3];
  int D_off, U_off, L_off;  // 8x8 ; (contact,body) rows 4..7: 4x6 ; (body,contact) 6x8
  int color;                // index among the contacts of the same body
};

// One elimination step of the block LDU (GraphBasedSystems ldu_factorization!)
struct ElimNb {
  int n, vec_off;   // neighbour dimension / offset of its entry in the solution-ordered vectors
  int L_off;        // M_{nb,c}: n_nb x n_c   (overwritten by M_{nb,c} * Dinv_c)
  int U_off, U_k;   // M_{c,nb}: rows [U_row, U_row + U_k) of c, U_k x n_nb (never written by the factorisation)
  int U_row;
};
struct ElimStep {
  int d_off, n, vec_off;
  int nnb;
  ElimNb nb[2];
  int tgt[2][2];  // M_{nb_i, nb_j}
};

struct Plan {
  int Nb, Ne, Ni, nres, nu, nz;
  int nsteps;
  double h, input_scaling, g[3];
  // arena layout (doubles)
  int sol_off, rhs_off, sav_off, mat_off, mat_len, arena_len;  // [mat_off, mat_off + mat_len) is re-zeroed at every assembly
  const BodyDev* bodies;
  const JointDev* joints;
  const ContactDev* contacts;
  const ElimStep* steps;
};

struct Options {
  double rtol, btol, undercut, no_progress_undercut;
  int max_iter, max_ls, no_progress_max;
};

}  // namespace dj
