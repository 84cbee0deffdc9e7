// dojo_joint_tra.cuh -- translational springs, dampers and limits of a joint:
//
//   spring_impulses      joints/translational/springs.jl:5-29     explicit, at (x2, q2): force k A'(offset - A e) in the parent frame
//   damper_impulses      joints/translational/dampers.jl:5-35     implicit: force -d A'A (e(x2, q2) - e(x1, q1)) / h with the backward
//                        (+ minimal.jl:93-113 minimal_velocities)   step x1 = x2 - h v25, q1 = q2 (x) m(-w25) of the CANDIDATE velocities
//   damper_jacobian_velocity  dampers.jl:100-123 (+ minimal.jl:155-193)  6 x 6 blocks on and between the two bodies
//   limits               joints/limits.jl:1-29 on the translational coordinates A e(x3, q3)
//
// Every force f (parent frame) reaches the bodies through the transposed displacement Jacobian at the CURRENT configuration
// (impulse_transform, joints/translational/impulses.jl):  body force = X' f, body torque = 1/2 Qt' f  =>  G6 = [X'; Qt'/2] (6 x 3).
// Closed forms checked block by block against the dense KKT matrix of the reference algorithm (tests).
// Included from inside namespace dj by dojo_kernels.cuh, DJ_ANY_CONTACT compilation only (dojo_b200_cm.cu).
#pragma once

DJ_DEV void g6_apply(const M33& X, const M33& Qt, V3 f, V3& fl, V3& fa) { fl = tmul(X, f); fa = 0.5 * tmul(Qt, f); }

// out (6 x 6, row-major) += sgn * G6(X, Qt) * [Fv | Fw]   with Fv, Fw the 3 x 3 derivatives of a parent-frame force
DJ_DEV void g6_accumulate(double* out, const M33& X, const M33& Qt, const M33& Fv, const M33& Fw, double sgn) {
  M33 Xt = transpose(X), Qh = 0.5 * transpose(Qt);
  M33 a = Xt * Fv, b = Xt * Fw, cc = Qh * Fv, d = Qh * Fw;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      out[i * 6 + j] += sgn * a.m[i][j];
      out[i * 6 + 3 + j] += sgn * b.m[i][j];
      out[(3 + i) * 6 + j] += sgn * cc.m[i][j];
      out[(3 + i) * 6 + 3 + j] += sgn * d.m[i][j];
    }
}

// prologue: explicit spring impulse and the impulse maps t6 = G6 a_i of the limit duals (constant over the solve)
DJ_DEV void prologue_joint_tra(Ctx& c, const JointDev& jd, const JointGeom& g, V3& cl_p, V3& ca_p, V3& cl_c, V3& ca_c) {
  const Plan& P = *c.P;
  double* A = c.A;
  if (jd.flags & JF_TRA_SPRING) {
    const double* tp = joint_tra_params(jd);
    V3 f = v3zero();
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i < jd.nfree_t) { V3 a = ld3(jd.At + 3 * i); f += (tp[0] * (tp[2 + i] - dot(a, g.et))) * a; }
    f = P.h * f;
    V3 fl, fa;
    g6_apply(g.Xp, g.Qtp, f, fl, fa); cl_p += fl; ca_p += fa;
    g6_apply(g.Xc, g.Qtc, f, fl, fa); cl_c += fl; ca_c += fa;
  }
  if (jd.flags & JF_LIM_TRA) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i < jd.nb2_r) {
        double* lim = A + jd.lim_off + 2 * kLim * i;  // aP(6) aC(6) tP(6) tC(6)
        V3 a = ld3(jd.At + 3 * i), fl, fa;
        g6_apply(g.Xp, g.Qtp, a, fl, fa); st3(lim + 12, fl); st3(lim + 15, fa);
        g6_apply(g.Xc, g.Qtc, a, fl, fa); st3(lim + 18, fl); st3(lim + 21, fa);
      }
  }
}

// evaluation at sol + f * delta: damper impulse / velocity Jacobians, limit rows and their condensed coupling.
// K6pp / K6cc: the bodies do D -= K6;  B6pc / B6cp: the (parent, child) / (child, parent) blocks themselves.
template <bool JAC>
DJ_DEV void eval_joint_tra(Ctx& c, const JointDev& jd, double f, const Kin& ka, const Kin& kb, const JointGeom& g3, double* rr, const double* so,
                           const double* dd, double& bv, V3& fl_p, V3& fa_p, V3& fl_c, V3& fa_c, double* K6pp, double* K6cc, double* B6pc,
                           double* B6cp) {
  const Plan& P = *c.P;
  double* A = c.A;
  if (jd.flags & JF_TRA_DAMPER) {
    const double damper = joint_tra_params(jd)[1];
    M33 Ra2 = rotmat(ka.q2), Rb2 = rotmat(kb.q2);
    JointGeom g2 = joint_geom(jd, ka.x2, ka.q2, Ra2, kb.x2, kb.q2, Rb2);
    V3 xa1 = ka.x2 - P.h * ka.v, xb1 = kb.x2 - P.h * kb.v;
    Quat qa1 = qmul(ka.q2, qmap(-ka.w, P.h)), qb1 = qmul(kb.q2, qmap(-kb.w, P.h));
    M33 Ra1 = rotmat(qa1), Rb1 = rotmat(qb1);
    JointGeom g1 = joint_geom(jd, xa1, qa1, Ra1, xb1, qb1, Rb1);
    M33 AtA = m33zero();
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i < jd.nfree_t) { V3 a = ld3(jd.At + 3 * i); AtA = AtA + outer(a, a); }
    V3 fd = (-damper) * (AtA * (g2.et - g1.et));  // h * damper force
    V3 fl, fa;
    g6_apply(g2.Xp, g2.Qtp, fd, fl, fa); fl_p += fl; fa_p += fa;
    g6_apply(g2.Xc, g2.Qtc, fd, fl, fa); fl_c += fl; fa_c += fa;
    if (JAC) {
      // h d(force)/d v = -h damper A'A X1,  h d(force)/d w = -damper A'A Qt1 E(-w)   (e1 moves by -h X1 dv - Qt1 E(-w) dw)
      M33 Fva = (-P.h * damper) * (AtA * g1.Xp), Fvb = (-P.h * damper) * (AtA * g1.Xc);
      M33 Fwa = (-damper) * (AtA * (g1.Qtp * attitude_velocity_jacobian(-ka.w, P.h)));
      M33 Fwb = (-damper) * (AtA * (g1.Qtc * attitude_velocity_jacobian(-kb.w, P.h)));
      g6_accumulate(K6cc, g2.Xc, g2.Qtc, Fvb, Fwb, 1.0);      // D_child  -= d(impulse_c)/d(vel_c)
      if (jd.parent >= 0) {
        g6_accumulate(K6pp, g2.Xp, g2.Qtp, Fva, Fwa, 1.0);    // D_parent -= d(impulse_p)/d(vel_p)
        g6_accumulate(B6pc, g2.Xp, g2.Qtp, Fvb, Fwb, -1.0);   // (parent, child) = -d(impulse_p)/d(vel_c)
        g6_accumulate(B6cp, g2.Xc, g2.Qtc, Fva, Fwa, -1.0);   // (child, parent) = -d(impulse_c)/d(vel_p)
      }
    }
  }
  if (jd.flags & JF_LIM_TRA) {
    M33 QtpE, QtcE;
    if (JAC) { QtpE = g3.Qtp * ka.E; QtcE = g3.Qtc * kb.E; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (i < jd.nb2_r) {
        V3 ai = ld3(jd.At + 3 * i);
        double th = dot(ai, g3.et);
        const int is_u = jd.ne + i, is_l = jd.ne + jd.nb2_r + i;
        const int ig_u = is_u + jd.nb_r, ig_l = is_l + jd.nb_r;
        double su = so[is_u], sl = so[is_l], gu = so[ig_u], gl = so[ig_l];
        if (f != 0.0) { su += f * dd[is_u]; sl += f * dd[is_l]; gu += f * dd[ig_u]; gl += f * dd[ig_l]; }
        bv = nanmax(bv, nanmax(fabs(su * gu), fabs(sl * gl)));
        rr[is_u] = -(su * gu - c.mu);
        rr[is_l] = -(sl * gl - c.mu);
        rr[ig_u] = -(su - (jd.hi[i] - th));
        rr[ig_l] = -(sl - (th - jd.lo[i]));
        double* lim = A + jd.lim_off + 2 * kLim * i;
        V3 tPl = ld3(lim + 12), tPa = ld3(lim + 15), tCl = ld3(lim + 18), tCa = ld3(lim + 21);
        fl_p += (gl - gu) * tPl; fa_p += (gl - gu) * tPa;
        fl_c += (gl - gu) * tCl; fa_c += (gl - gu) * tCa;
        if (JAC) {
          V3 aPl = P.h * vtmul(ai, g3.Xp), aPa = vtmul(ai, QtpE), aCl = P.h * vtmul(ai, g3.Xc), aCa = vtmul(ai, QtcE);
          st3(lim, aPl); st3(lim + 3, aPa); st3(lim + 6, aCl); st3(lim + 9, aCa);
          // kept side / condensed side exactly as for the rotational limits (dojo_kernels.cuh limit_side, dojo_plan.h joint_nq)
          const LimitSide ls = limit_side(su, sl, gu, gl);
          const double kk = ls.kI;
          const double tP[6] = {tPl.x, tPl.y, tPl.z, tPa.x, tPa.y, tPa.z}, tC[6] = {tCl.x, tCl.y, tCl.z, tCa.x, tCa.y, tCa.z};
          const double aP[6] = {aPl.x, aPl.y, aPl.z, aPa.x, aPa.y, aPa.z}, aC[6] = {aCl.x, aCl.y, aCl.z, aCa.x, aCa.y, aCa.z};
#pragma unroll
          for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              K6pp[r * 6 + q] -= kk * tP[r] * aP[q];
              K6cc[r * 6 + q] -= kk * tC[r] * aC[q];
              B6pc[r * 6 + q] += kk * tP[r] * aC[q];
              B6cp[r * 6 + q] += kk * tC[r] * aP[q];
            }
          const int n = joint_nq(jd), q = jd.ne + i;
          double* D = A + jd.D_off;
          double* Uc = A + jd.Uc_off;
          double* Lcw = A + jd.Lc_off;
          D[q * n + q] = ls.sA;
#pragma unroll
          for (int r = 0; r < 6; ++r) { Uc[q * 6 + r] = (-ls.sg * ls.gA) * aC[r]; Lcw[r * n + q] = ls.sg * tC[r]; }
          if (jd.parent >= 0) {
            double* Up = A + jd.Up_off;
            double* Gpw = A + jd.Gp_off;
#pragma unroll
            for (int r = 0; r < 6; ++r) { Up[q * 6 + r] = (-ls.sg * ls.gA) * aP[r]; Gpw[r * n + q] = -ls.sg * tP[r]; }
          }
        }
      }
    }
  }
}

DJ_DEV void condense_joint_tra(Ctx& c, const JointDev& jd, double* x) {
  const Plan& P = *c.P;
  double* A = c.A;
  V3 pl = v3zero(), pa = v3zero(), cl = v3zero(), ca = v3zero();
  const double* so = A + P.sol_off + jd.sol_off;
  double* xr = x + jd.sol_off;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < jd.nb2_r) {
      const int is_u = jd.ne + i, is_l = jd.ne + jd.nb2_r + i, ig_u = is_u + jd.nb_r, ig_l = is_l + jd.nb_r;
      const LimitSide ls = limit_side(so[is_u], so[is_l], so[ig_u], so[ig_l]);
      const double rc_u = xr[is_u], rc_l = xr[is_l], rs_u = xr[ig_u], rs_l = xr[ig_l];
      const double cI = ((ls.up ? rc_l : rc_u) - ls.gI * (ls.up ? rs_l : rs_u)) / ls.sI;
      const double w = ls.sg * cI;
      const double* lim = A + jd.lim_off + 2 * kLim * i;
      pl += w * ld3(lim + 12); pa += w * ld3(lim + 15);
      cl += w * ld3(lim + 18); ca += w * ld3(lim + 21);
      xr[is_u] = (ls.up ? rc_u : rc_l) - ls.gA * (ls.up ? rs_u : rs_l);
      if (!ls.up) xr[is_l] = rc_u;
    }
  }
  double* sc = A + jd.slot_c;
  st3(sc, cl); st3(sc + 3, ca);
  if (jd.parent >= 0) { double* sp = A + jd.slot_p; st3(sp, pl); st3(sp + 3, pa); }
}

DJ_DEV void recover_joint_tra(Ctx& c, const JointDev& jd, double* x) {
  const Plan& P = *c.P;
  double* A = c.A;
  const double* so = A + P.sol_off + jd.sol_off;
  double* xr = x + jd.sol_off;
  V3 vp = v3zero(), wp = v3zero();
  if (jd.parent >= 0) { vp = ld3(x + c.bodies[jd.parent].sol_off); wp = ld3(x + c.bodies[jd.parent].sol_off + 3); }
  V3 vc = ld3(x + c.bodies[jd.child].sol_off), wc = ld3(x + c.bodies[jd.child].sol_off + 3);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < jd.nb2_r) {
      const int is_u = jd.ne + i, is_l = jd.ne + jd.nb2_r + i, ig_u = is_u + jd.nb_r, ig_l = is_l + jd.nb_r;
      const LimitSide ls = limit_side(so[is_u], so[is_l], so[ig_u], so[ig_l]);
      const double* lim = A + jd.lim_off + 2 * kLim * i;
      const double adw = dot(ld3(lim), vp) + dot(ld3(lim + 3), wp) + dot(ld3(lim + 6), vc) + dot(ld3(lim + 9), wc);
      const double dgA = xr[is_u], rcI = xr[is_l], rs_u = xr[ig_u], rs_l = xr[ig_l];
      const double ds_u = rs_u - adw, ds_l = rs_l + adw;
      const double dgI = (rcI - ls.gI * (ls.up ? ds_l : ds_u)) / ls.sI;
      xr[is_u] = ds_u;
      xr[is_l] = ds_l;
      xr[ig_u] = ls.up ? dgA : dgI;
      xr[ig_l] = ls.up ? dgI : dgA;
    }
  }
}
