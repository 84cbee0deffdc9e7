// dojo_storage.cuh -- trajectory recording on the device: what save_to_storage! computes next to the
// state after every solve of simulate!, plus the momentum / energy diagnostics the reference derives from a Storage.
//
//   save_to_storage!            simulation/storage.jl:50-67   per body: px, pq (world frame), vl = px / m, wl = J \ R(q2)' pq
//   momentum(mechanism, body)   mechanics/momentum.jl:17-52   D2x, D2q at (x2, q2, v25, w25) minus half of the input, joint,
//                                                             spring and damper impulses (joint_impulses :44-52)
//   momentum(mech, storage, t)  mechanics/momentum.jl:54-74   total linear / angular momentum about the centre of mass
//   kinetic_energy              mechanics/energy.jl:32-41     sum 1/2 m vl'vl + 1/2 wl' J wl
//   potential_energy            mechanics/energy.jl:60-93     -sum m g'x + sum 1/2 |spring force|^2 / k
//
// Inputs per environment: z (state before the solve: x2, q2), z' (v25, w25 in its velocity slots), u, and the solver
// solution in the reference ordering (joint impulses).  Closed forms: D2q = m0 J w + (h/2) w x J w;
// joint impulse transforms T = [X'; (Q LV')' / 2] written with rotation matrices:
//   translational  parent [-Ra; -[e + pa]x]   child [Ra; [pb]x Rb' Ra]        e = Ra'(xb + Rb pb - xa) - pa
//   rotational     parent -Roff (s I + [v]x) / 2   child (s I - [v]x) / 2     (s, v) = qoff^-1 qa^-1 qb
// One THREAD per environment: recording is off the hot path; traffic 8 (26 Nb + nu + nres + 12 Nb + 8) bytes.
#pragma once
#include "dojo_kin.cuh"

namespace dj {

struct StorageArgs {
  const BodyDev* bodies;
  const JointDev* joints;
  int Ne, Nb, nu, nres, B;
  double h, input_scaling, g[3];
  const double* Z;     // [13 Nb x B] state before the solve
  const double* Zn;    // [13 Nb x B] state after the solve (v25, w25)
  const double* U;     // [nu x B] nullable
  const double* sol;   // [nres x B] reference ordering
  double* body_out;    // [12 Nb x B]: per body px(3) pq(3) vl(3) wl(3)
  double* diag;        // [8 x B]: p_linear(3) p_angular(3) kinetic potential
};

DJ_DEV M33 inverse33(const M33& a) {
  const double c00 = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1], c01 = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2],
               c02 = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
  const double det = a.m[0][0] * c00 + a.m[0][1] * c01 + a.m[0][2] * c02, id = 1.0 / det;
  M33 r;
  r.m[0][0] = c00 * id; r.m[1][0] = c01 * id; r.m[2][0] = c02 * id;
  r.m[0][1] = (a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2]) * id;
  r.m[1][1] = (a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0]) * id;
  r.m[2][1] = (a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1]) * id;
  r.m[0][2] = (a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1]) * id;
  r.m[1][2] = (a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2]) * id;
  r.m[2][2] = (a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]) * id;
  return r;
}
DJ_DEV V3 mask_row_s(const double* A, int i) { return v3(A[3 * i], A[3 * i + 1], A[3 * i + 2]); }
DJ_DEV M33 sI_plus_s(double s, V3 v) { return m33ident(s) + skew(v); }
DJ_DEV M33 sI_minus_s(double s, V3 v) { return m33ident(s) - skew(v); }
DJ_DEV M33 m33_from(const double* p) {
  M33 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = p[3 * i + j];
  return r;
}

// all impulses a joint applies to one of its bodies during the step, [F(3); tau(3)] in the convention of the body residual:
// impulse map * lambda + spring + damper (joint_impulses, momentum.jl:44-52) and, separately, the input impulse JF2 / Jtau2
// (set_input!, mechanism/set.jl:40-53; translational/input.jl:5-27, rotational/input.jl:5-17).
struct JointWrench { V3 F, tau, JF, Jtau; };

DJ_DEV JointWrench joint_wrench(const StorageArgs& a, const JointDev& jd, bool parent, const BodyState& A, const BodyState& Bc, const double* u,
                                const double* sol) {
  JointWrench w;
  w.F = w.tau = w.JF = w.Jtau = v3zero();
  const double h = a.h;
  const V3 pa = v3(jd.pa[0], jd.pa[1], jd.pa[2]), pb = v3(jd.pb[0], jd.pb[1], jd.pb[2]);
  const Quat qoff = Quat{jd.qoff[0], jd.qoff[1], jd.qoff[2], jd.qoff[3]};
  const M33 Ra = rotmat(A.q), Rb = rotmat(Bc.q);
  // reference ordering of the joint's entry: [tra eq (nl_t) | s (nb_r) | gamma (nb_r) | rot eq (nl_r)] with rotational limits,
  // [s (nb_r) | gamma (nb_r) | tra eq (nl_t) | rot eq (nl_r)] with translational limits (JF_LIM_TRA: the limits belong to the first half)
  const bool lim_tra = (jd.flags & JF_LIM_TRA) != 0;
  const double* lam = sol + jd.sol_off + (lim_tra ? 2 * jd.nb_r : 0);
  // ---- translational element
  {
    V3 f3 = v3zero();  // C' lambda
    for (int i = 0; i < jd.nl_t; ++i) f3 += lam[i] * v3(jd.Ct[3 * i], jd.Ct[3 * i + 1], jd.Ct[3 * i + 2]);
    V3 in3 = v3zero();
    if (u) in3 = a.input_scaling * masked_sum(jd.At, jd.nfree_t, u + jd.u_off);
    const V3 e = tra_displacement(jd, A.x, A.q, Bc.x, Bc.q);
    if (lim_tra) {  // limit duals: A' (gamma_lower - gamma_upper)   (joints/joint.jl impulse_projector)
      const double* gam = sol + jd.sol_off + jd.nb_r;
      for (int i = 0; i < jd.nb2_r; ++i) f3 += (gam[jd.nb2_r + i] - gam[i]) * mask_row_s(jd.At, i);
    }
    if (jd.flags & (JF_TRA_SPRING | JF_TRA_DAMPER)) {  // translational/springs.jl:5-29, dampers.jl:5-35 (velocities v25, w25)
      const double* tp = joint_tra_params(jd);
      V3 fs = v3zero();
      if (jd.flags & JF_TRA_SPRING)
        for (int i = 0; i < jd.nfree_t; ++i) fs += (h * tp[0] * (tp[2 + i] - dot(mask_row_s(jd.At, i), e))) * mask_row_s(jd.At, i);
      if (jd.flags & JF_TRA_DAMPER) {
        const Quat qa1 = next_orientation(A.q, -A.w, h), qb1 = next_orientation(Bc.q, -Bc.w, h);
        const V3 e1 = tra_displacement(jd, A.x - h * A.v, qa1, Bc.x - h * Bc.v, qb1);
        for (int i = 0; i < jd.nfree_t; ++i) fs += (-tp[1] * dot(mask_row_s(jd.At, i), e - e1)) * mask_row_s(jd.At, i);
      }
      f3 += fs;
    }
    if (parent) {
      w.F += -1.0 * (Ra * f3);  w.tau += -1.0 * cross(e + pa, f3);
      w.JF += -1.0 * (Ra * in3); w.Jtau += (-0.5) * cross(e + pa, in3);
    } else {
      const M33 Tq = skew(pb) * transpose(Rb) * Ra;
      w.F += Ra * f3;  w.tau += Tq * f3;
      w.JF += Ra * in3; w.Jtau += 0.5 * (Tq * in3);
    }
  }
  // ---- rotational element
  {
    const Quat q = qmul(qmul(qinv(qoff), qinv(A.q)), Bc.q);
    const M33 Roff = rotmat(qoff);
    const M33 Rrel = rotmat(qmul(qmul(qinv(Bc.q), A.q), qoff));  // rotation_matrix(inv(qb) * qa * qoff)
    V3 f3 = v3zero();
    const double* gam = lam + jd.nl_t + jd.nb_r;
    if (!lim_tra) for (int i = 0; i < jd.nb2_r; ++i) f3 += (gam[jd.nb2_r + i] - gam[i]) * mask_row_s(jd.Ar, i);
    const double* lr = lam + jd.nl_t + (lim_tra ? 0 : 2 * jd.nb_r);
    for (int i = 0; i < jd.nl_r; ++i) f3 += lr[i] * mask_row_s(jd.Cr, i);
    if (parent) w.tau += (-0.5) * (Roff * (sI_plus_s(q.s, qvec(q)) * f3));
    else w.tau += 0.5 * (sI_minus_s(q.s, qvec(q)) * f3);
    if (u) {
      const V3 tq = a.input_scaling * masked_sum(jd.Ar, jd.nfree_r, u + jd.u_off + jd.nfree_t);
      if (parent) w.Jtau += -1.0 * (Roff * tq);
      else w.Jtau += Rrel * tq;
    }
    if (jd.nfree_r > 0 && jd.spring_r != 0.0) {  // rotational/springs.jl:5-38
      const V3 th = rotation_vector(q);
      V3 dist = v3zero();
      for (int i = 0; i < jd.nfree_r; ++i) dist += (jd.spring_off_r[i] - dot(mask_row_s(jd.Ar, i), th)) * mask_row_s(jd.Ar, i);
      const V3 force = (-jd.spring_r) * dist;
      w.tau += parent ? h * (Roff * force) : h * (Rrel * (-1.0 * force));
    }
    if (jd.nfree_r > 0 && jd.damper_r != 0.0) {  // rotational/dampers.jl:4-27 (velocities v25, w25)
      const Quat qa1 = next_orientation(A.q, -A.w, h), qb1 = next_orientation(Bc.q, -Bc.w, h);
      const Quat q1 = qmul(qmul(qinv(qoff), qinv(qa1)), qb1);
      const V3 rv = (1.0 / h) * rotation_vector(qmul(qinv(q1), q));
      V3 vel = v3zero();
      for (int i = 0; i < jd.nfree_r; ++i) vel += dot(mask_row_s(jd.Ar, i), rv) * mask_row_s(jd.Ar, i);
      w.tau += parent ? h * (Roff * (jd.damper_r * vel)) : h * (Rrel * ((-jd.damper_r) * vel));
    }
  }
  return w;
}

DJ_DEV void storage_env(const StorageArgs& a, int e) {
  const double* z = a.Z + (size_t)e * 13 * a.Nb;
  const double* zn = a.Zn + (size_t)e * 13 * a.Nb;
  const double* u = a.U ? a.U + (size_t)e * a.nu : nullptr;
  const double* sol = a.sol + (size_t)e * a.nres;
  double* out = a.body_out + (size_t)e * 12 * a.Nb;
  const double h = a.h;
  const V3 g = v3(a.g[0], a.g[1], a.g[2]);
  double mass = 0.0, kinetic = 0.0, potential = 0.0;
  V3 com = v3zero(), P = v3zero();
  for (int b = 0; b < a.Nb; ++b) {
    const BodyDev& bd = a.bodies[b];
    BodyState s = kin_load(z, b);
    const BodyState sn = kin_load(zn, b);
    s.v = sn.v; s.w = sn.w;  // (x2, v25, q2, w25): current_configuration_velocity after the solve
    const M33 J = m33_from(bd.J);
    const V3 x3 = s.x + h * s.v;
    const double m0 = 0.5 * h * sqrt(4.0 / (h * h) - dot(s.w, s.w));
    const V3 Jw = J * s.w;
    V3 p_lin = (1.0 / h * bd.mass) * (x3 - s.x) - (0.5 * h * bd.mass) * g;
    V3 p_ang = m0 * Jw + (0.5 * h) * cross(s.w, Jw);
    for (int j = 0; j < a.Ne; ++j) {
      const JointDev& jd = a.joints[j];
      if (jd.parent != b && jd.child != b) continue;
      BodyState A = kin_load(z, jd.parent), Bc = kin_load(z, jd.child);
      if (jd.parent >= 0) { const BodyState an = kin_load(zn, jd.parent); A.v = an.v; A.w = an.w; }
      { const BodyState bn = kin_load(zn, jd.child); Bc.v = bn.v; Bc.w = bn.w; }
      const JointWrench w = joint_wrench(a, jd, jd.parent == b, A, Bc, u, sol);
      p_lin -= 0.5 * (w.F + w.JF);
      p_ang -= 0.5 * (w.tau + w.Jtau);
    }
    const V3 pq = rotmat(s.q) * p_ang;  // vector_rotate(p_angular_body, q2)
    const V3 vl = (1.0 / bd.mass) * p_lin;
    const V3 wl = inverse33(J) * tmul(rotmat(s.q), pq);
    double* o = out + 12 * b;
    o[0] = p_lin.x; o[1] = p_lin.y; o[2] = p_lin.z; o[3] = pq.x; o[4] = pq.y; o[5] = pq.z;
    o[6] = vl.x; o[7] = vl.y; o[8] = vl.z; o[9] = wl.x; o[10] = wl.y; o[11] = wl.z;
    mass += bd.mass; com += bd.mass * s.x; P += p_lin;
    kinetic += 0.5 * bd.mass * dot(vl, vl) + 0.5 * dot(wl, J * wl);
    potential -= bd.mass * dot(g, s.x);
  }
  com = (1.0 / mass) * com;
  const V3 vcom = (1.0 / mass) * P;
  V3 L = v3zero();
  for (int b = 0; b < a.Nb; ++b) {
    const BodyDev& bd = a.bodies[b];
    const double* o = out + 12 * b;
    const V3 r = v3(z[13 * b], z[13 * b + 1], z[13 * b + 2]) - com;
    L += v3(o[3], o[4], o[5]) + cross(r, bd.mass * (v3(o[6], o[7], o[8]) - vcom));
  }
  for (int j = 0; j < a.Ne; ++j) {  // spring potential (energy.jl:69-90): 1/2 |force|^2 / k
    const JointDev& jd = a.joints[j];
    if ((jd.flags & JF_TRA_SPRING) && joint_tra_params(jd)[0] > 0.0) {
      const BodyState A = kin_load(z, jd.parent), Bc = kin_load(z, jd.child);
      const double* tp = joint_tra_params(jd);
      const V3 e = tra_displacement(jd, A.x, A.q, Bc.x, Bc.q);
      double d2 = 0.0;
      for (int i = 0; i < jd.nfree_t; ++i) { const double di = tp[2 + i] - dot(mask_row_s(jd.At, i), e); d2 += di * di; }
      potential += 0.5 * tp[0] * d2;
    }
    if (jd.nfree_r == 0 || !(jd.spring_r > 0.0)) continue;
    const BodyState A = kin_load(z, jd.parent), Bc = kin_load(z, jd.child);
    const Quat q = qmul(qmul(qinv(Quat{jd.qoff[0], jd.qoff[1], jd.qoff[2], jd.qoff[3]}), qinv(A.q)), Bc.q);
    const V3 th = rotation_vector(q);
    V3 dist = v3zero();
    for (int i = 0; i < jd.nfree_r; ++i) dist += (jd.spring_off_r[i] - dot(mask_row_s(jd.Ar, i), th)) * mask_row_s(jd.Ar, i);
    potential += 0.5 * jd.spring_r * dot(dist, dist);
  }
  double* d = a.diag + (size_t)e * 8;
  d[0] = P.x; d[1] = P.y; d[2] = P.z; d[3] = L.x; d[4] = L.y; d[5] = L.z; d[6] = kinetic; d[7] = potential;
}

#ifdef __CUDACC__
__global__ void dojo_storage_kernel(const StorageArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.B) storage_env(a, e);
}
#endif

}  // namespace dj
