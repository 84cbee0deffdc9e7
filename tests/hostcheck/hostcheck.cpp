// hostcheck.cpp -- TEST INFRASTRUCTURE.  Compiles the per-environment kinematics code of the product
// (dojo.jl_b200/csrc/dojo_kin.cuh, dojo_kinjac.cuh: the exact functions the CUDA kernels call) with g++ and runs it with
// ONE "thread" (tid = 0, nthr = 1, no-op barrier), so that `pytest -m "not gpu"` can compare the device arithmetic with the
// oracle on a machine without a GPU.  Nothing in the product library links or loads this file; on the GPU the same
// functions are exercised through the C-ABI by tests/test_gpu_kinjac.py.
#include <cstring>
#include <vector>

#include "../../dojo.jl_b200/csrc/dojo_kinjac.cuh"
#include "../../dojo.jl_b200/csrc/dojo_envs.cuh"
#include "../../dojo.jl_b200/csrc/dojo_storage.cuh"

using namespace dj;

static void pad_mask(int nlambda, const double* axis_mask, double* A, double* Cm = nullptr) {  // nullspace / constraint mask rows (joints/joint.jl:56-64)
  double dummy[9];
  double* Cc = Cm ? Cm : dummy;
  std::memset(A, 0, 9 * sizeof(double));
  std::memset(Cc, 0, 9 * sizeof(double));
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (nlambda == 0) std::memcpy(A, I3, sizeof(I3));
  else if (nlambda == 1) { std::memcpy(Cc, axis_mask + 6, 3 * sizeof(double)); std::memcpy(A, axis_mask, 3 * sizeof(double)); std::memcpy(A + 3, axis_mask + 3, 3 * sizeof(double)); }
  else if (nlambda == 2) { std::memcpy(Cc, axis_mask, 6 * sizeof(double)); std::memcpy(A, axis_mask + 6, 3 * sizeof(double)); }
  else std::memcpy(Cc, I3, sizeof(I3));
}

struct Mech {
  std::vector<BodyDev> bodies;
  std::vector<JointDev> joints;
  std::vector<int> order;
  int Ne, Nb, nu;
  double h;
};

extern "C" {

// jint [Ne][4] = parent, child, nlambda_tra, nlambda_rot ;  jdbl [Ne][28] = pa(3) pb(3) qoff(4) axis_mask_tra(9) axis_mask_rot(9)
void* hostcheck_create(int Ne, int Nb, double h, const int* jint, const double* jdbl, const int* order) {
  Mech* m = new Mech;
  m->Ne = Ne; m->Nb = Nb; m->h = h;
  m->joints.resize(Ne);
  int uoff = 0;
  for (int j = 0; j < Ne; ++j) {
    JointDev& J = m->joints[j];
    std::memset(&J, 0, sizeof(J));
    J.parent = jint[4 * j]; J.child = jint[4 * j + 1];
    J.nl_t = jint[4 * j + 2]; J.nl_r = jint[4 * j + 3];
    J.nfree_t = 3 - J.nl_t; J.nfree_r = 3 - J.nl_r;
    J.u_off = uoff; uoff += J.nfree_t + J.nfree_r;
    const double* d = jdbl + 28 * j;
    std::memcpy(J.pa, d, 24); std::memcpy(J.pb, d + 3, 24); std::memcpy(J.qoff, d + 6, 32);
    pad_mask(J.nl_t, d + 10, J.At, J.Ct);
    pad_mask(J.nl_r, d + 19, J.Ar, J.Cr);
  }
  m->nu = uoff;
  m->order.assign(order, order + Ne);
  return m;
}
void hostcheck_destroy(void* p) { delete static_cast<Mech*>(p); }
int hostcheck_num_input(void* p) { return static_cast<Mech*>(p)->nu; }

static KinJacArgs base_args(Mech* m, int B) {
  KinJacArgs a;
  std::memset(&a, 0, sizeof(a));
  a.joints = m->joints.data(); a.order = m->order.data();
  a.Ne = m->Ne; a.Nb = m->Nb; a.nu = m->nu; a.B = B; a.h = m->h;
  return a;
}
static void run(const KinJacArgs& a0) {
  KinJacArgs a = a0;
  std::vector<double> ws(kinjac_ws_doubles(a.Nb, a.nu) + 1);
  a.ws = ws.data();
  for (int e = 0; e < a.B; ++e) kinjac_env(a, e, a.ws, 0, 1, [] {});
}
void hostcheck_minimal_to_maximal(void* p, int B, const double* X, double* Z) {
  Mech* m = static_cast<Mech*>(p);
  KinArgs a; a.joints = m->joints.data(); a.order = m->order.data(); a.Ne = m->Ne; a.Nb = m->Nb; a.nu = m->nu; a.B = B; a.h = m->h; a.in = X; a.out = Z;
  for (int e = 0; e < B; ++e) min_to_max_env(a, e);
}
void hostcheck_maximal_to_minimal(void* p, int B, const double* Z, double* X) {
  Mech* m = static_cast<Mech*>(p);
  KinArgs a; a.joints = m->joints.data(); a.order = m->order.data(); a.Ne = m->Ne; a.Nb = m->Nb; a.nu = m->nu; a.B = B; a.h = m->h; a.in = Z; a.out = X;
  for (int e = 0; e < B; ++e) max_to_min_env(a, e);
}
void hostcheck_max_to_min_jacobian(void* p, int B, const double* Z, double* J) {  // J zero-filled here
  Mech* m = static_cast<Mech*>(p);
  KinJacArgs a = base_args(m, B);
  std::memset(J, 0, sizeof(double) * (size_t)B * 2 * m->nu * 12 * m->Nb);
  a.Z = Z; a.Zm = Z; a.outM = J; a.mode = 0;
  run(a);
}
void hostcheck_min_to_max_jacobian(void* p, int B, const double* Z, double* J) {
  Mech* m = static_cast<Mech*>(p);
  KinJacArgs a = base_args(m, B);
  a.Z = Z; a.Zm = Z; a.outN = J; a.mode = 1;
  run(a);
}
void hostcheck_minimal_gradients(void* p, int B, const double* Z, const double* Zn, const double* Fz, const double* Fu, double* Gx, double* Gu) {
  Mech* m = static_cast<Mech*>(p);
  KinJacArgs a = base_args(m, B);
  a.Z = Z; a.Zm = Zn; a.Fz = Fz; a.Fu = Fu; a.Gx = Gx; a.Gu = Gu; a.mode = 2;
  run(a);
}

// environment layer (dojo_envs.cuh).  spec_i [5] / spec_d [7] = the DojoEnvSpec fields in declaration order.
static EnvArgs env_args(Mech* m, const int* spec_i, const double* spec_d, int Ni, int nres, const ContactDev* contacts, int B) {
  EnvArgs a;
  std::memset(&a, 0, sizeof(a));
  a.joints = m->joints.data(); a.contacts = contacts; a.order = m->order.data();
  a.Ne = m->Ne; a.Nb = m->Nb; a.Ni = Ni; a.nu = m->nu; a.nres = nres; a.B = B; a.h = m->h;
  a.spec.n_unactuated = spec_i[0]; a.spec.contact_obs = spec_i[1]; a.spec.forward_index = spec_i[2]; a.spec.healthy_index = spec_i[3];
  a.spec.bound_index = spec_i[4];
  a.spec.w_forward = spec_d[0]; a.spec.w_control = spec_d[1]; a.spec.w_contact = spec_d[2]; a.spec.survive_reward = spec_d[3];
  a.spec.healthy_min = spec_d[4]; a.spec.healthy_max = spec_d[5]; a.spec.bound_abs = spec_d[6];
  return a;
}
void hostcheck_env_pre(void* p, const int* spec_i, const double* spec_d, int Ni, int B, const double* S, const double* A, double* Z, double* U) {
  Mech* m = static_cast<Mech*>(p);
  EnvArgs a = env_args(m, spec_i, spec_d, Ni, 0, nullptr, B);
  a.S = S; a.A = A; a.Z = Z; a.U = U;
  for (int e = 0; e < B; ++e) env_pre(a, e);
}
void hostcheck_env_post(void* p, const int* spec_i, const double* spec_d, int Ni, int nres, const int* contact_sol_off, int B, const double* S,
                        const double* A, const double* Zn, const double* sol, double* Sn, double* reward, int32_t* done, double* ret, int32_t* dead) {
  Mech* m = static_cast<Mech*>(p);
  std::vector<ContactDev> contacts(Ni > 0 ? Ni : 1);
  std::memset(contacts.data(), 0, sizeof(ContactDev) * contacts.size());
  for (int c = 0; c < Ni; ++c) { contacts[c].sol_off = contact_sol_off[c]; contacts[c].tn = 2 | (4 << 8); }  // NonlinearContact entries [s(4); gamma(4)]
  EnvArgs a = env_args(m, spec_i, spec_d, Ni, nres, contacts.data(), B);
  a.S = S; a.A = A; a.Zn = Zn; a.sol = sol; a.Sn = Sn; a.reward = reward; a.done = done; a.ret = ret; a.dead = dead;
  for (int e = 0; e < B; ++e) env_post(a, e);
}

// storage / diagnostics (dojo_storage.cuh).  jext [Ne][12] = rot limits (Nb/2), spring_r, damper_r, spring_offset_r(3), then the same for tra;
// bdbl [Nb][10] = mass, inertia (row-major).  Solution offsets follow dojo_create: joints in order, n = nl_t + nl_r + 4 * limits.
void hostcheck_storage(void* p, const double* jext, const double* bdbl, int nres, double input_scaling, const double* g, int B, const double* Z,
                       const double* Zn, const double* U, const double* sol, double* body_out, double* diag) {
  Mech* m = static_cast<Mech*>(p);
  m->bodies.resize(m->Nb);
  int off = 0;
  for (int j = 0; j < m->Ne; ++j) {
    JointDev& J = m->joints[j];
    const double* d = jext + 12 * j;
    J.nb2_r = (int)d[0]; J.nb_r = 2 * J.nb2_r; J.spring_r = d[1]; J.damper_r = d[2];
    for (int i = 0; i < 3; ++i) J.spring_off_r[i] = d[3 + i];
    // translational half as dojo_create encodes it (JointDev::flags, joint_tra_params)
    J.flags = 0;
    if (J.nfree_t > 0) {
      if (d[7] != 0.0) J.flags |= JF_TRA_SPRING;
      if (d[8] != 0.0) J.flags |= JF_TRA_DAMPER | JF_FULL;
      if ((int)d[6] > 0) { J.flags |= JF_LIM_TRA | JF_FULL; J.nb2_r = (int)d[6]; J.nb_r = 2 * J.nb2_r; }
      if (J.flags & (JF_TRA_SPRING | JF_TRA_DAMPER)) { double* tp = joint_tra_params(J); tp[0] = d[7]; tp[1] = d[8]; for (int i = 0; i < J.nfree_t; ++i) tp[2 + i] = d[9 + i]; }
    }
    J.ne = J.nl_t + J.nl_r; J.n = J.ne + 2 * J.nb_r;
    J.sol_off = off; off += J.n;
  }
  for (int b = 0; b < m->Nb; ++b) {
    BodyDev& Bd = m->bodies[b];
    std::memset(&Bd, 0, sizeof(Bd));
    Bd.mass = bdbl[10 * b];
    std::memcpy(Bd.J, bdbl + 10 * b + 1, 9 * sizeof(double));
    Bd.sol_off = off; off += 6;
  }
  StorageArgs a;
  std::memset(&a, 0, sizeof(a));
  a.bodies = m->bodies.data(); a.joints = m->joints.data();
  a.Ne = m->Ne; a.Nb = m->Nb; a.nu = m->nu; a.nres = nres; a.B = B; a.h = m->h; a.input_scaling = input_scaling;
  for (int i = 0; i < 3; ++i) a.g[i] = g[i];
  a.Z = Z; a.Zn = Zn; a.U = U; a.sol = sol; a.body_out = body_out; a.diag = diag;
  for (int e = 0; e < B; ++e) storage_env(a, e);
}

void hostcheck_env_policy(int ns, int na, int B, const double* S, const double* Theta, const double* mean, const double* stdev, double* A) {
  PolicyArgs p;
  p.ns = ns; p.na = na; p.B = B; p.S = S; p.Theta = Theta; p.mean = mean; p.stdev = stdev; p.A = A;
  for (int e = 0; e < B; ++e) env_policy(p, e);
}
}
