# baseline/julia_threads_ensemble.jl — the north-star's CPU baseline: the UNMODIFIED reference functions, one
# ensemble member per Julia thread.  NOT MEASURED HERE (no Julia in the build image / on the GPU box); run with
#     julia -t auto --project=/path/to/MovingBoundaryMinerals.jl baseline/julia_threads_ensemble.jl 64 100
using MovingBoundaryMinerals, Random
function member_steps!(p, nsteps)
    V_ip, t_tot, t_ar, Di, D0, Ri, Lsc, Dsc, Vsc, tsc = scaling(p.Ri, p.Di, [NaN NaN], p.V_ip, p.t_tot, LinRange(0.0, p.t_tot, 1000))
    res = copy(p.res); resmin = copy(res)
    x_left, x_right, dx1, dx2, x0 = create_grid!(Ri, res, 20.0, false)
    C_left = 0.5 * ones(res[1], 1); C_right = (0.5 / p.KD) * ones(res[2], 1)
    KD_ar = LinRange(p.KD, p.KD, 1000); T_ar = LinRange(1273.15, 950.0, 1000)
    rho = [1.0 1.0]; n = 2; t = 0.0; it = 0; dt = 0.0
    Mass0 = calc_mass_vol(x_left, x_right, C_left, C_right, n, rho)
    D_l, D_r, KD, T = update_t_dependent_param!(D0, Di, NaN, NaN, KD_ar, NaN, T_ar, t_ar, t, t_tot)
    nodes = 0
    for _ in 1:nsteps
        dt = find_dt(dx1, dx2, V_ip, D_l, D_r, 0.3)
        t, dt, it = update_time!(t, dt, it, t_tot)
        D_l, D_r, KD, T = update_t_dependent_param!(D0, Di, NaN, NaN, KD_ar, NaN, T_ar, t_ar, t, t_tot)
        Fl, x_left, x_right, C_left, C_right, res, Ri = advect_interface_regrid!(Ri, V_ip, dt, x_left, x_right, C_left, C_right, res)
        _, _, dVolC = calc_volume(x_left, x_right, n)
        L_g, R_g, Co_l, Co_r = construct_matrix_fem(x_left, x_right, C_left, C_right, D_l, D_r, dt, n, res)
        L_g, R_g, ScF = set_inner_bc_mb!(L_g, R_g, dVolC, Mass0, KD, res)
        L_g, R_g = set_outer_bc!([0 0], L_g, R_g, Co_l[1], Co_r[end], ScF)
        C_left, C_right = solve_soe(L_g, R_g, res)
        nodes += length(C_left) + length(C_right)
        x_left, x_right, C_left, C_right, dx1, dx2, res = regrid!(Fl, x_left, x_right, C_left, C_right, Ri, V_ip, res, resmin, 20.0, false)
    end
    return nodes
end
function main(M, nsteps)
    rng = MersenneTwister(20251019)
    pars = [(Ri = [rand(rng) * 0.2 + 0.4 1.0], Di = [exp(rand(rng) * log(4) + log(2e-7)) exp(rand(rng) * log(4) + log(2.5e-8))],
             V_ip = (isodd(m) ? 1 : -1) * (rand(rng) * 1.5e-7 + 0.5e-7), t_tot = 1e9, KD = rand(rng) * 0.15 + 0.05,
             res = [1024 1024]) for m in 1:M]
    member_steps!(pars[1], 2)
    total = Threads.Atomic{Int}(0)
    el = @elapsed Threads.@threads for m in 1:M
        Threads.atomic_add!(total, member_steps!(pars[m], nsteps))
    end
    println("threads=", Threads.nthreads(), " members=", M, " steps=", nsteps, " node-steps/s=", total[] / el)
end
main(parse(Int, get(ARGS, 1, "16")), parse(Int, get(ARGS, 2, "20")))
