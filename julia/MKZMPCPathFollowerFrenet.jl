#!/usr/bin/env julia
#
# Drop-in replacement for scripts/mpc_utils/MKZMPCPathFollowerFrenet.jl of govvijaycal/mkz_mpc_path_follower
# (the module scripts/nodes_gazebo_sim/gazebo_sim_mpc_cmd_pub_frenet.jl drives): same module name, functions,
# argument order and return tuples; JuMP/Ipopt replaced by libmpc_b200.so (mpcb200_*_frenet of
# include/mpc_b200.h) through `ccall`, batch of one.
#
# NOT EXECUTED in the build container (no Julia there); the identical call sequence is exercised by the Python
# mirror mkz_mpc_path_follower_b200/mpc_path_follower.py::MKZMPCPathFollowerFrenet.
# Julia versions: the shim itself is written to load under Julia 0.6 (`mutable struct` exists there; `Cvoid` is aliased
# below) as well as >= 1.0.  The reference's own node scripts are Julia 0.6-era code (`unshift!`, JuMP `sum{}`, `@sprintf`
# without `using Printf`): on Julia 0.6 they run on top of this shim as they are; on Julia >= 1.0 they need the same
# one-line modernisations they would need with the original module.  Neither combination has been executed (no Julia in
# the build container); what HAS driven this exact call sequence through the C ABI is the Python mirror and the plain-C
# caller tests/c_abi/mpc_cmd_loop.c, and tests/test_c_abi.py checks this struct's field list against the C header.

module MKZMPCPathFollowerFrenet

const libmpc = get(ENV, "MPCB200_LIB", "libmpc_b200.so")
@static if VERSION < v"0.7"
    const Cvoid = Void
end

mutable struct Config   # mpcb200_config
    N::Int32; max_iter::Int32; start_mode::Int32; device::Int32
    dt::Float64; dt_control::Float64; L_a::Float64; L_b::Float64
    v_min::Float64; v_max::Float64; a_max::Float64; steer_max::Float64
    a_dmax::Float64; steer_dmax::Float64; tol::Float64
    n_devices::Int32; devices::NTuple{8,Int32}
    Config() = new()
end

function check(rc::Cint, h::Ptr{Cvoid})
    if rc != 0
        error("libmpc_b200: ", unsafe_string(ccall((:mpcb200_last_error, libmpc), Cstring, (Ptr{Cvoid},), h)))
    end
end

const N = 8                       # horizon (:34)
const cfg = Config()
check(ccall((:mpcb200_default_config, libmpc), Cint, (Ref{Config}, Int32), cfg, Int32(N)), C_NULL)
const dt = cfg.dt
const dt_control = cfg.dt_control

const handle = Ref{Ptr{Cvoid}}(C_NULL)
check(ccall((:mpcb200_create_frenet, libmpc), Cint, (Ref{Ptr{Cvoid}}, Ref{Config}), handle, cfg), C_NULL)
atexit(() -> ccall((:mpcb200_destroy, libmpc), Cint, (Ptr{Cvoid},), handle[]))

path_ref = Dict()                                      # :38
const state = zeros(4)                                 # s0, ey0, epsi0, v0 (:106-109)
const k_poly = zeros(4)                                # [a3, a2, a1, a0] (:40-41)
const v_target = [15.0]                                # v_ref (:39)
const u_curr = zeros(2)                                # d_f_current, acc_current
const warm = zeros(6 * N + 4)                          # start = 0.0 (:65-73); afterwards the last solution
const u0 = zeros(2); const cost = zeros(1)
const status = zeros(Int32, 1); const iters = zeros(Int32, 1)
const STATUS = (:Optimal, :Infeasible, :Unbounded, :UserLimit, :Error)

function solve_batch_of_one()
    check(ccall((:mpcb200_solve_batch_frenet, libmpc), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Int32),
                handle[], 1, state, k_poly, v_target, u_curr, warm, u0, cost, status, iters, C_NULL, 0), handle[])
    return STATUS[status[1] + 1]
end

println("MPC: Initial solve ...")                                                    # :126-128
println("MPC: Finished initial solve: ", solve_batch_of_one())

function update_init_cond(s::Float64, ey::Float64, epsi::Float64, vel::Float64)      # :132-138
    state[1] = s; state[2] = ey; state[3] = epsi; state[4] = vel
end

function update_reference(path::Dict, k_coeffs::Array{Float64,1}, v_des::Float64)    # :142-147
    global path_ref
    path_ref = path
    k_poly[1:4] = k_coeffs
    v_target[1] = v_des
end

function update_current_input(c_swa::Float64, c_acc::Float64)                        # :151-154
    u_curr[1] = c_swa; u_curr[2] = c_acc
end

function update_cost(cey::Float64, cep::Float64, cev::Float64,
                     cda::Float64, cdd::Float64, ca::Float64, cd::Float64)           # :158-169
    w = [cey, cep, cev, cda, cdd, ca, cd]
    check(ccall((:mpcb200_set_cost_frenet, libmpc), Cint, (Ptr{Cvoid}, Ptr{Float64}), handle[], w), handle[])
end

function solve_model()                                                               # :173-183
    st = solve_batch_of_one()
    return u0[1], u0[2], st
end

function get_solver_results()                                                        # :188-207
    s_mpc = warm[1:(N + 1)]; ey_mpc = warm[(N + 2):(2N + 2)]
    v_mpc = warm[(2N + 3):(3N + 3)]; epsi_mpc = warm[(3N + 4):(4N + 4)]
    d_f_opt = warm[(4N + 5):(5N + 4)]; acc_opt = warm[(5N + 5):(6N + 4)]
    return s_mpc, ey_mpc, v_mpc, epsi_mpc, copy(k_poly), path_ref, d_f_opt, acc_opt
end

end
