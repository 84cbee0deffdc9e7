// dojo_contact_orthant.cuh -- the two positive-orthant contact models of the reference:
//
//   ImpactContact{T,2}   contacts/impact.jl:8-146   entry [s; gamma],                          rows [s gamma - mu; d - s]
//   LinearContact{T,12}  contacts/linear.jl:10-104  entry [s_gamma s_psi s_beta(4); gamma psi beta(4)],
//                        rows [s .* gamma - mu; d - s_gamma; mu_f gamma - sum(beta) - s_psi; P vt + psi 1 - s_beta]
//                        with the 4-sided friction pyramid P = [0 1; 0 -1; 1 0; -1 0] (linear.jl:29-34)
//
// Included by dojo_kernels.cuh only in the translation unit that is compiled with DJ_ANY_CONTACT (dojo_b200_cm.cu): the kernels of
// mechanisms whose contacts are all NonlinearContact (every BASELINE model) are compiled without it and stay bit-identical.
// Same structure as the NonlinearContact code in dojo_kernels.cuh: one lane per contact, closed-form elimination of the
// contact node (the first step of the reference's LDU: contacts are leaves), condensation onto the body's 6 x 6 block.
// (included from inside namespace dj, after Kin / body_kin / nanmax)
#pragma once

// Closed-form solve of the orthant contact block  D_c y = t,  y = [ds(nh); dgamma(nh)]  (impact.jl:58-64, linear.jl:49-70):
//   rows 0..nh-1   : g_i' ds_i + s_i' dg_i = t_i                      (s' = s + REG, g' = g + REG: neutral vector = ones)
//   rows nh..2nh-1 : -ds + C dg = t2,   C = 0 (impact),  C dg = [0; mu_f dg_0 - sum(dbeta); dpsi 1(4)] (linear)
// => (Diag(s') + Diag(g') C) dg = t + g' .* t2 =: b, solved by substitution:
//   dg_0 = b_0 / s_0';  dbeta_k = (b_k - g_k' dpsi) / s_k';  dpsi (s_1' + g_1' sum g_k'/s_k') = b_1 - g_1' mu_f dg_0 + g_1' sum b_k / s_k'
// (all terms of the dpsi pivot are positive: no cancellation while the iterate is strictly inside the orthant)
DJ_DEV void orthant_solve(int type, const double* s, const double* g, double muf, const double* t, double* y) {
  if (type == 0) {
    const double sp = s[0] + kReg, gp = g[0] + kReg;
    const double ds = -t[1];
    y[0] = ds;
    y[1] = (t[0] - gp * ds) / sp;
    return;
  }
  double sp[6], gp[6], b[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) { sp[i] = s[i] + kReg; gp[i] = g[i] + kReg; b[i] = t[i] + gp[i] * t[6 + i]; }
  const double dg0 = b[0] / sp[0];
  double piv = sp[1], rhs = b[1] - gp[1] * muf * dg0;
#pragma unroll
  for (int k = 2; k < 6; ++k) { const double r = 1.0 / sp[k]; piv += gp[1] * gp[k] * r; rhs += gp[1] * b[k] * r; }
  const double dpsi = rhs / piv;
  double dg[6];
  dg[0] = dg0; dg[1] = dpsi;
  double sb = 0.0;
#pragma unroll
  for (int k = 2; k < 6; ++k) { dg[k] = (b[k] - gp[k] * dpsi) / sp[k]; sb += dg[k]; }
  y[0] = -t[6];
  y[1] = muf * dg0 - sb - t[7];
#pragma unroll
  for (int k = 2; k < 6; ++k) y[k] = dpsi - t[6 + k];
#pragma unroll
  for (int i = 0; i < 6; ++i) y[6 + i] = dg[i];
}

// geometry shared by the three models (collisions/sphere_halfspace.jl, contacts/velocity.jl)
struct ContactGeom {
  V3 n, t0, t1, o, rc, ww, vc;
  double phi;
};
DJ_DEV ContactGeom contact_geom(const ContactDev& cd, const Kin& k) {
  ContactGeom q;
  q.n = ld3(cd.n); q.t0 = ld3(cd.t); q.t1 = ld3(cd.t + 3); q.o = ld3(cd.o);
  V3 off = ld3(cd.off);
  V3 ow = k.R3 * q.o;
  q.rc = ow - off - cd.radius * q.n;
  q.phi = dot(q.n, k.x3 + ow - off) - cd.radius;
  q.ww = k.R3 * k.w;
  q.vc = k.v + cross(q.ww, q.rc);
  return q;
}
// contact force X gamma: impact X = n' (impact.jl:103-115); linear X = [n' 0 T'P'] (contact.jl:141-155) = n g0 + t1 (b0 - b1) + t0 (b2 - b3)
DJ_DEV V3 orthant_force(int type, const ContactGeom& q, const double* g) {
  V3 F = g[0] * q.n;
  if (type == 1) F = F + (g[2] - g[3]) * q.t1 + (g[4] - g[5]) * q.t0;
  return F;
}
// rows of J (nh x 6) = d(constraint rows)/d(v25, w25) and columns of G (6 x nh) = impulse map from the three basic rows / columns
// [normal | tangent 0 | tangent 1]:  impact [normal];  linear [normal; 0; t1; -t1; t0; -t0]  (P = [0 1; 0 -1; 1 0; -1 0])
DJ_DEV void orthant_expand(const double* bn, const double* b0, const double* b1, double out[6][6]) {
#pragma unroll
  for (int cc = 0; cc < 6; ++cc) {
    out[0][cc] = bn[cc];
    out[1][cc] = 0.0;
    out[2][cc] = b1[cc]; out[3][cc] = -b1[cc];
    out[4][cc] = b0[cc]; out[5][cc] = -b0[cc];
  }
}

template <bool JAC>
DJ_DEV void eval_contact_orthant(Ctx& c, int idx, double f, double* res, double& rv, double& bv) {
  const Plan& P = *c.P;
  double* A = c.A;
  const double* sol = A + P.sol_off;
  const double* dl = A + P.rhs_off;
  const ContactDev& cd = c.contacts[idx];
  const int type = contact_type(cd), nh = contact_nh(cd);
  Kin k = body_kin(c, cd.body, f);
  double s[6], g[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    s[i] = 1.0; g[i] = 1.0;
    if (i < nh) {
      s[i] = sol[cd.sol_off + i];
      g[i] = sol[cd.sol_off + nh + i];
      if (f != 0.0) { s[i] += f * dl[cd.sol_off + i]; g[i] += f * dl[cd.sol_off + nh + i]; }
    }
  }
  ContactGeom q = contact_geom(cd, k);
  double r[6];
  r[0] = q.phi - s[0];
  if (type == 1) {
    const double vt0 = dot(q.t0, q.vc), vt1 = dot(q.t1, q.vc);
    r[1] = cd.mu * g[0] - (g[2] + g[3] + g[4] + g[5]) - s[1];
    r[2] = vt1 + g[1] - s[2];
    r[3] = -vt1 + g[1] - s[3];
    r[4] = vt0 + g[1] - s[4];
    r[5] = -vt0 + g[1] - s[5];
  }
  double* rr = res + cd.sol_off;
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (i < nh) {
      const double ci = g[i] * s[i];
      rv = nanmax(rv, fabs(r[i]));
      bv = nanmax(bv, fabs(ci));
      rr[i] = -(ci - c.mu);
      rr[nh + i] = -r[i];
    }
  V3 F = orthant_force(type, q, g);
  V3 tau = tmul(k.R3, cross(q.rc, F));
  st3(A + cd.slot + c.sd, F);
  st3(A + cd.slot + c.sd + 3, tau);
  if (JAC) {
    M33 R3so = k.R3 * skew(q.o);
    V3 nphi = (-2.0) * vtmul(q.n, R3so);
    V3 r4w = vtmul(nphi, k.E);
    V3 hn = P.h * q.n;
    M33 dvc_dw = (-1.0) * (skew(q.rc) * k.R3);
    M33 dvc_dd = 2.0 * (skew(q.rc) * (k.R3 * skew(k.w))) - 2.0 * (skew(q.ww) * R3so);
    M33 W3 = dvc_dw + dvc_dd * k.E;
    V3 r6w = vtmul(q.t0, W3), r7w = vtmul(q.t1, W3);
    const double Jn[6] = {hn.x, hn.y, hn.z, r4w.x, r4w.y, r4w.z};
    const double J0[6] = {q.t0.x, q.t0.y, q.t0.z, r6w.x, r6w.y, r6w.z};
    const double J1[6] = {q.t1.x, q.t1.y, q.t1.z, r7w.x, r7w.y, r7w.z};
    double Jr[6][6];
    orthant_expand(Jn, J0, J1, Jr);
    V3 qn = tmul(k.R3, cross(q.rc, q.n)), q0 = tmul(k.R3, cross(q.rc, q.t0)), q1 = tmul(k.R3, cross(q.rc, q.t1));
    const double Gn[6] = {q.n.x, q.n.y, q.n.z, qn.x, qn.y, qn.z};
    const double G0[6] = {q.t0.x, q.t0.y, q.t0.z, q0.x, q0.y, q0.z};
    const double G1[6] = {q.t1.x, q.t1.y, q.t1.z, q1.x, q1.y, q1.z};
    double Gt[6][6];  // Gt[col][row]: transposed impulse map
    orthant_expand(Gn, G0, G1, Gt);
    double* Jm = A + cd.J_off;
    double* Gm = A + cd.G_off;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i < nh) {
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) { Jm[i * 6 + cc] = Jr[i][cc]; Gm[cc * nh + i] = Gt[i][cc]; }
      }
    // condensation onto the body:  dgamma = w0 - W J dv  =>  D_b += G W J;  column cc of G W J = G * (D_c^-1 [0; J(:, cc)])_gamma
    M33 K = 2.0 * skew(tau) + 2.0 * (transpose(k.R3) * (skew(F) * R3so));
    M33 KE = K * k.E;
    double* slotK = A + cd.slot + 6;  // the body does D_b -= slotK
#pragma unroll 1
    for (int cc = 0; cc < 6; ++cc) {
      double t[12], y[12];
#pragma unroll
      for (int i = 0; i < 6; ++i) { t[i] = 0.0; t[6 + i] = 0.0; }
#pragma unroll
      for (int i = 0; i < 6; ++i) if (i < nh) t[nh + i] = Jr[i][cc];
      orthant_solve(type, s, g, cd.mu, t, y);
#pragma unroll
      for (int rI = 0; rI < 6; ++rI) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i < nh) acc += Gt[i][rI] * y[nh + i];
        double v = -acc;
        if (rI >= 3 && cc >= 3) v += KE.m[rI - 3][cc - 3];
        slotK[rI * 6 + cc] = v;
      }
    }
  }
}

DJ_DEV void condense_contact_orthant(Ctx& c, int idx, const double* x) {
  const Plan& P = *c.P;
  double* A = c.A;
  const ContactDev& cd = c.contacts[idx];
  const int type = contact_type(cd), nh = contact_nh(cd);
  const double* so = A + P.sol_off + cd.sol_off;
  double t[12], y[12];
  for (int i = 0; i < 2 * nh; ++i) t[i] = x[cd.sol_off + i];
  orthant_solve(type, so, so + nh, cd.mu, t, y);
  const double* G = A + cd.G_off;
  double* s = A + cd.slot;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double acc = 0.0;
    for (int i = 0; i < nh; ++i) acc += G[r * nh + i] * y[nh + i];
    s[r] = acc;
  }
}

DJ_DEV void recover_contact_orthant(Ctx& c, int idx, double* x) {
  const Plan& P = *c.P;
  double* A = c.A;
  const ContactDev& cd = c.contacts[idx];
  const int type = contact_type(cd), nh = contact_nh(cd);
  const double* so = A + P.sol_off + cd.sol_off;
  const double* J = A + cd.J_off;
  const double* dv = x + c.bodies[cd.body].sol_off;
  double t[12], y[12];
  for (int r = 0; r < nh; ++r) {
    t[r] = x[cd.sol_off + r];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += J[r * 6 + k] * dv[k];
    t[nh + r] = x[cd.sol_off + nh + r] - acc;
  }
  orthant_solve(type, so, so + nh, cd.mu, t, y);
  for (int r = 0; r < 2 * nh; ++r) x[cd.sol_off + r] = y[r];
}
