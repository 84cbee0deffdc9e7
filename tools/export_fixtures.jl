# export_fixtures.jl -- deferred true-parity hook (SURVEY.md §8c).  Pure Julia; NOT runnable in the build image.
# On a machine with Julia + Dojo 0.7.6 + DojoEnvironments:   julia --project tools/export_fixtures.jl out_dir
# Dumps, for pendulum / ant / quadruped / atlas and the widened models (contact types, translational springs / dampers / limits):
# the flattened mechanism (same fields as include/dojo_b200.h, with explicit
# name -> index maps because the body order is Dict-hash order), N random (z, u), and the reference results
# z_next (true and Q1-literal), the solution vector, iteration counts, full_matrix(system) and get_maximal_gradients (consistent: right
# after mehrotra!; literal: get_maximal_gradients!, i.e. after update_state!).  Loader: tests/test_julia_fixtures.py.
using Dojo, DojoEnvironments, JSON, Random, LinearAlgebra
include(joinpath(@__DIR__, "..", "ext", "DojoB200.jl"))
out = length(ARGS) > 0 ? ARGS[1] : "fixtures"; mkpath(out)
Random.seed!(100)
# (mechanism, builder keywords, tag): the BASELINE models, then the widened models of round 1 -- the three contact models
# (contact_type) and the translational springs / dampers / limits (cartpole options, raiberthopper's default leg damper)
configs = [(:pendulum, (;), "pendulum"), (:ant, (;), "ant"), (:quadruped, (;), "quadruped"), (:atlas, (;), "atlas"),
           (:sphere, (; contact_type=:linear), "sphere_linear"), (:sphere, (; contact_type=:impact), "sphere_impact"),
           (:block, (; contact_type=:linear), "block_linear"), (:block, (; contact_type=:impact), "block_impact"),
           (:cartpole, (; springs=2.0, dampers=0.3, joint_limits=Dict([(:cart_joint, [-0.3, 0.25]), (:pole_joint, [-1.2, 1.4])])), "cartpole"),
           (:raiberthopper, (;), "raiberthopper")]
for (name, kw, tag) in configs
    mech = get_mechanism(name; kw...)
    bodies, joints, contacts = DojoB200.flatten(mech)
    z = get_maximal_state(mech); nu = input_dimension(mech)
    cases = []
    for k in 1:20
        floating = length(mech.joints[1]) == 0   # a Floating joint to the origin carries no impulses: its 6 inputs stay zero
        u = floating ? [zeros(6); 0.5 .* randn(nu - 6)] : 0.2 .* randn(nu)
        m1 = deepcopy(mech)
        set_maximal_state!(m1, z); set_input!(m1, u)
        status = Dojo.mehrotra!(m1, opts = SolverOptions())
        sol = Dojo.get_solution(m1)
        znext_true = Dojo.get_next_state(m1)
        solmat = Dojo.full_matrix(m1.system)
        Fz, Fu = Dojo.get_maximal_gradients(m1)
        znext_q1 = step!(deepcopy(mech), z, u)
        Fz_lit, Fu_lit = get_maximal_gradients!(deepcopy(mech), z, u)   # the literal result (data Jacobian after update_state!, SURVEY Q2)
        push!(cases, Dict("z" => z, "u" => u, "status" => String(status), "sol" => sol, "z_next" => znext_true, "z_next_q1" => znext_q1,
                          "solmat" => vec(solmat), "Fz" => vec(Fz), "Fu" => vec(Fu), "Fz_literal" => vec(Fz_lit), "Fu_literal" => vec(Fu_lit)))
        z = znext_true
    end
    open(joinpath(out, "$(tag).json"), "w") do io
        JSON.print(io, Dict("body_names" => [String(b.name) for b in mech.bodies], "joint_names" => [String(j.name) for j in mech.joints],
                            "contact_names" => [String(c.name) for c in mech.contacts], "timestep" => mech.timestep, "cases" => cases))
    end
end
