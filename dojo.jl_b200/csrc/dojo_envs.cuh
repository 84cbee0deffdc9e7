// dojo_envs.cuh -- batched environment layer on the device: what DojoEnvironments does around every
// step_minimal_coordinates! call, fused into the two thread-per-environment kernels either side of the step kernel.
//
//   state_map / input_map         DojoEnvironments/src/environments/ant_ars.jl:53-61, quadruped_sampling.jl:51-58, pendulum.jl:41-47
//   step!(environment, x, u)      DojoEnvironments/src/environments.jl:77-84 (= step_minimal_coordinates!, simulation/step.jl:42-61)
//   get_state                     ant_ars.jl:72-79 (minimal state + clamp(gamma_1, -1, 1) per contact), others: minimal state
//   reward / failure test         examples/learning/ant_ars.jl:79-116 (rollout_policy), quadruped_sampling.jl:66-77
//
// pre  kernel: environment state s -> x = state_map(s) -> z = minimal_to_maximal(x);  action a -> u = input_map(a)
// post kernel: z' , solution -> s' = [maximal_to_minimal(z'); clamp(gamma_1)], reward, done
// One THREAD per environment (same reasoning as dojo_kin.cuh); HBM traffic per environment-step:
// 8 (2 ns + na + 2 * 13 Nb + nu + Ni) + 12 bytes on top of the step's own.
#pragma once
#include "dojo_kin.cuh"

namespace dj {

struct EnvSpec {  // = DojoEnvSpec (include/dojo_b200.h)
  int n_unactuated, contact_obs, forward_index, healthy_index, bound_index;
  double w_forward, w_control, w_contact, survive_reward, healthy_min, healthy_max, bound_abs;
};

struct EnvArgs {
  const JointDev* joints;
  const ContactDev* contacts;
  const int* order;
  int Ne, Nb, Ni, nu, nres, B;
  double h;
  EnvSpec spec;
  const double* S;       // [ns x B] environment states
  const double* A;       // [na x B] actions (nullable: zero input)
  double* Z;             // pre: out [13 Nb x B]
  double* U;             // pre: out [nu x B]
  const double* Zn;      // post: [13 Nb x B]
  const double* sol;     // post: [nres x B] solution in the reference ordering (contacts: [s(4); gamma(4)])
  double* Sn;            // post: out [ns x B]
  double* reward;        // post: out [B] (nullable)
  int32_t* done;         // post: out [B] (nullable)
  double* ret;           // post: in/out [B] (nullable) return accumulated over a rollout: += reward while the environment is alive
  int32_t* dead;         // post: in/out [B] (nullable) set once the failure test fires (the failing step's reward still counts,
                         //       as in rollout_policy, examples/learning/ant_ars.jl:106-115)
};

DJ_DEV int env_num_state(int nu, int Ni, const EnvSpec& sp) { return 2 * nu + (sp.contact_obs ? Ni : 0); }

DJ_DEV void env_pre(const EnvArgs& a, int e) {
  const int ns = env_num_state(a.nu, a.Ni, a.spec), na = a.nu - a.spec.n_unactuated;
  min_to_max_one(a.joints, a.order, a.Ne, a.h, a.S + (size_t)e * ns, a.Z + (size_t)e * 13 * a.Nb);  // state_map: s[1:2nu]
  double* u = a.U + (size_t)e * a.nu;
  for (int i = 0; i < a.spec.n_unactuated; ++i) u[i] = 0.0;                                        // input_map: [zeros; a]
  for (int i = 0; i < na; ++i) u[a.spec.n_unactuated + i] = a.A ? a.A[(size_t)e * na + i] : 0.0;
}

DJ_DEV double clamp_unit(double v) { return fmax(-1.0, fmin(1.0, v)); }  // max(-1, min(1, v)), ant_ars.jl:75

DJ_DEV void env_post(const EnvArgs& a, int e) {
  const EnvSpec& sp = a.spec;
  const int ns = env_num_state(a.nu, a.Ni, sp), na = a.nu - sp.n_unactuated;
  const double* s = a.S + (size_t)e * ns;
  double* sn = a.Sn + (size_t)e * ns;
  max_to_min_one(a.joints, a.Ne, a.h, a.Zn + (size_t)e * 13 * a.Nb, sn);
  double contact_cost = 0.0;
  for (int c = 0; c < a.Ni; ++c) {
    const double g = clamp_unit(a.sol[(size_t)e * a.nres + a.contacts[c].sol_off + (a.contacts[c].tn >> 8)]);  // contact.impulses[2][1] (entry = [s(N½); gamma(N½)])
    if (sp.contact_obs) sn[2 * a.nu + c] = g;
    contact_cost += g * g;
  }
  double r = 0.0;
  if (a.reward || a.ret) {
    double ctrl = 0.0;
    if (a.A) for (int i = 0; i < na; ++i) { const double v = a.A[(size_t)e * na + i]; ctrl += v * v; }
    r = sp.survive_reward - sp.w_control * ctrl - sp.w_contact * contact_cost;
    if (sp.forward_index >= 0) r += sp.w_forward * (sn[sp.forward_index] - s[sp.forward_index]) / a.h;
    if (a.reward) a.reward[e] = r;
    if (a.ret && !(a.dead && a.dead[e])) a.ret[e] += r;
  }
  if (a.done || a.dead) {
    bool ok = true;
    for (int i = 0; i < ns; ++i) ok = ok && (fabs(sn[i]) <= 1.79769313486231570e308);  // all(isfinite.(state_after))
    if (sp.healthy_index >= 0) ok = ok && (sn[sp.healthy_index] >= sp.healthy_min) && (sn[sp.healthy_index] <= sp.healthy_max);
    if (sp.bound_index >= 0) ok = ok && (fabs(sn[sp.bound_index]) <= sp.bound_abs);
    if (a.done) a.done[e] = ok ? 0 : 1;
    if (a.dead && !ok) a.dead[e] = 1;
  }
}

// linear policy with observation normalisation (examples/learning/ant_ars.jl:47-50 normalize, :88 action = theta * state):
//   a = Theta_e ((s - mean) ./ std),  Theta [na x ns x B] column-major per environment (ARS evaluates one perturbed policy per
//   environment), mean / std [ns] shared (nullable: no normalisation)
struct PolicyArgs {
  int ns, na, B;
  const double* S;
  const double* Theta;
  const double* mean;
  const double* stdev;
  double* A;
};
DJ_DEV void env_policy(const PolicyArgs& p, int e) {
  const double* s = p.S + (size_t)e * p.ns;
  const double* th = p.Theta + (size_t)e * p.ns * p.na;
  double* a = p.A + (size_t)e * p.na;
  for (int k = 0; k < p.na; ++k) a[k] = 0.0;
  for (int i = 0; i < p.ns; ++i) {
    const double o = p.mean ? (s[i] - p.mean[i]) / p.stdev[i] : s[i];
    for (int k = 0; k < p.na; ++k) a[k] += th[(size_t)i * p.na + k] * o;
  }
}

// reset: environments with mask != 0 (or all, mask == nullptr) get the initial state s0 [ns]
DJ_DEV void env_reset(int ns, const double* s0, const int32_t* mask, double* S, int e) {
  if (mask && mask[e] == 0) return;
  for (int i = 0; i < ns; ++i) S[(size_t)e * ns + i] = s0[i];
}

#ifdef __CUDACC__
__global__ void dojo_env_pre_kernel(const EnvArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.B) env_pre(a, e);
}
__global__ void dojo_env_post_kernel(const EnvArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < a.B) env_post(a, e);
}
__global__ void dojo_env_policy_kernel(const PolicyArgs p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < p.B) env_policy(p, e);
}
__global__ void dojo_env_reset_kernel(int ns, int B, const double* s0, const int32_t* mask, double* S) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < B) env_reset(ns, s0, mask, S, e);
}
#endif

}  // namespace dj
