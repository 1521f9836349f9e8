#!/usr/bin/env julia
#
# Drop-in replacement for scripts/mpc_utils/MKZMPCPathFollower.jl of govvijaycal/mkz_mpc_path_follower:
# same module name, the same six functions with the same argument order and return tuples, and the
# same readable constants `dt` and `N` (mpc_cmd_pub.jl:51).  JuMP/Ipopt are replaced by libmpc_b200.so
# (include/mpc_b200.h) through `ccall`, with a batch of one; mpc_cmd_pub.jl keeps calling the same six functions.
#
# NOT EXECUTED in the build container (no Julia there); the identical call sequence is exercised by
# the Python mirror mkz_mpc_path_follower_b200/mpc_path_follower.py, which binds the same C ABI.
# Julia versions: the shim itself is written to load under Julia 0.6 (`mutable struct` exists there; `Cvoid` is aliased
# below) as well as >= 1.0.  The reference's own node scripts are Julia 0.6-era code (`unshift!`, JuMP `sum{}`, `@sprintf`
# without `using Printf`): on Julia 0.6 they run on top of this shim as they are; on Julia >= 1.0 they need the same
# one-line modernisations they would need with the original module.  Neither combination has been executed (no Julia in
# the build container); what HAS driven this exact call sequence through the C ABI is the Python mirror and the plain-C
# caller tests/c_abi/mpc_cmd_loop.c, and tests/test_c_abi.py checks this struct's field list against the C header.

module MKZMPCPathFollower

const libmpc = get(ENV, "MPCB200_LIB", "libmpc_b200.so")
@static if VERSION < v"0.7"
    const Cvoid = Void
end

# mpcb200_config (include/mpc_b200.h); field order and types must match the C struct
mutable struct Config
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

#### (1) model constants: MKZMPCPathFollower.jl:28-48 come back from mpcb200_default_config
const N = 8                       # horizon (:34); change here for another horizon
const cfg = Config()
check(ccall((:mpcb200_default_config, libmpc), Cint, (Ref{Config}, Int32), cfg, Int32(N)), C_NULL)
const dt = cfg.dt                 # model discretization time (:33)
const dt_control = cfg.dt_control # control period (:28)
const L_a = cfg.L_a
const L_b = cfg.L_b

const handle = Ref{Ptr{Cvoid}}(C_NULL)
check(ccall((:mpcb200_create, libmpc), Cint, (Ref{Ptr{Cvoid}}, Ref{Config}), handle, cfg), C_NULL)
atexit(() -> ccall((:mpcb200_destroy, libmpc), Cint, (Ptr{Cvoid},), handle[]))

#### parameters (the @NLparameter values of :51-59, :75, :82, :91-94, :110-113)
const v_ref = 15.0
const state = zeros(4)                                 # x0, y0, psi0, v0
const ref = zeros(3 * (N + 1))                         # x_r, y_r, psi_r
ref[1:(N + 1)] = v_ref * collect(0.0:dt:N * dt)
const v_target = [v_ref]
const u_curr = zeros(2)                                # d_f_current, acc_current
const warm = zeros(6 * N + 4)                          # start = 0.0 (:65-72); afterwards the last solution
const u0 = zeros(2); const cost = zeros(1)
const status = zeros(Int32, 1); const iters = zeros(Int32, 1)
const STATUS = (:Optimal, :Infeasible, :Unbounded, :UserLimit, :Error)

function solve_batch_of_one()
    check(ccall((:mpcb200_solve_batch, libmpc), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Int32),
                handle[], 1, state, ref, v_target, u_curr, warm, u0, cost, status, iters, C_NULL, 0), handle[])
    return STATUS[status[1] + 1]
end

#### (5) Initialize Solver (:126-128)
println("MPC: Initial solve ...")
println("MPC: Finished initial solve: ", solve_batch_of_one())

function update_init_cond(x::Float64, y::Float64, psi::Float64, vel::Float64)       # :132-138
    state[1] = x; state[2] = y; state[3] = psi; state[4] = vel
end

function update_reference(x_ref::Array{Float64,1}, y_ref::Array{Float64,1}, psi_ref::Array{Float64,1}, v_des::Float64)  # :142-147
    ref[1:(N + 1)] = x_ref; ref[(N + 2):(2N + 2)] = y_ref; ref[(2N + 3):(3N + 3)] = psi_ref
    v_target[1] = v_des
end

function update_current_input(c_swa::Float64, c_acc::Float64)                        # :151-154 (steering first)
    u_curr[1] = c_swa; u_curr[2] = c_acc
end

function update_cost(cx::Float64, cy::Float64, cp::Float64, cv::Float64,
                     cda::Float64, cdd::Float64, ca::Float64, cd::Float64)          # :158-169
    w = [cx, cy, cp, cv, cda, cdd, ca, cd]
    check(ccall((:mpcb200_set_cost, libmpc), Cint, (Ptr{Cvoid}, Ptr{Float64}), handle[], w), handle[])
end

function solve_model()                                                               # :173-183
    st = solve_batch_of_one()
    return u0[1], u0[2], st          # acc_opt[1], d_f_opt[1], status
end

function get_solver_results()                                                        # :188-207
    x_mpc = warm[1:(N + 1)]; y_mpc = warm[(N + 2):(2N + 2)]
    v_mpc = warm[(2N + 3):(3N + 3)]; psi_mpc = warm[(3N + 4):(4N + 4)]
    d_f_opt = warm[(4N + 5):(5N + 4)]; acc_opt = warm[(5N + 5):(6N + 4)]
    return x_mpc, y_mpc, v_mpc, psi_mpc, ref[1:(N + 1)], ref[(N + 2):(2N + 2)], ref[(2N + 3):(3N + 3)], d_f_opt, acc_opt
end

end
