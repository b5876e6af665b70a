# MovingBoundaryMineralsB200.jl — ccall shim over libmbm_b200.so (include/mbm_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It is the binding a maintainer of
# MovingBoundaryMinerals.jl would add; struct layouts mirror include/mbm_b200.h field by field (the same
# layouts are checked against the C header for the Python ctypes mirror in tests/test_abi.py).
#
# Usage (drop-ins for the four main codes; same keyword names as the literals at the top of each driver, same return tuples):
#     using MovingBoundaryMinerals, MovingBoundaryMineralsB200
#     x_left, x_right, dx1, dx2, x0, res, Ri, C_left, C_right, C0 = MovingBoundaryMineralsB200.DCGMB()
#     ..., C0, Mass, Mass0                                         = MovingBoundaryMineralsB200.DCGF()
#     x, C, x0, C0, D, t, t_tot, Mass, Mass0                        = MovingBoundaryMineralsB200.SDG()
#     x_left, x_right, x0, C_left, C_right, C0, maxC, ...           = MovingBoundaryMineralsB200.CSP()
# and the batched forms:
#     out = MovingBoundaryMineralsB200.couple_ensemble(pars::Vector{<:NamedTuple}; kind = :mb)
#     out = MovingBoundaryMineralsB200.simple_ensemble(pars);  out = MovingBoundaryMineralsB200.stefan_ensemble(pars)
module MovingBoundaryMineralsB200

using MovingBoundaryMinerals: scaling, rescale            # host-side pre/post steps stay the reference's own
using MovingBoundaryMinerals: calc_mass_vol_simple_diff, coeff_trans_line, composition

const LIB = get(ENV, "MBM_B200_LIB", joinpath(@__DIR__, "..", "movingboundaryminerals.jl_b200", "csrc", "libmbm_b200.so"))
const ABI_VERSION = Int32(4)
const BC_FLUX, BC_MB, BC_STEFAN, BC_HOLD, BC_LASAGA = Int32(1), Int32(2), Int32(3), Int32(4), Int32(5)
const DT_FIND_DT, DT_ADVECTIVE, DT_STEFAN, DT_SIMPLE = Int32(0), Int32(1), Int32(2), Int32(3)
const FLAG_RECORD, FLAG_MASS_FINAL = UInt32(0x0002), UInt32(0x0080)

struct Common                      # mbm_common
    abi_version::Int32; n_members::Int32; ndim::Int32; hist_len::Int32
    cap_left::Int32; cap_right::Int32; inner_bc::Int32; dt_mode::Int32
    flags::UInt32; n_devices::Int32; max_steps::Int64; device_ids::Ptr{Int32}; rec_cap::Int64
end

struct CoupleParams                # mbm_couple_params
    Di::Ptr{Float64}; D0::Ptr{Float64}; Ea::Ptr{Float64}; rho::Ptr{Float64}; V_ip::Ptr{Float64}
    t_tot::Ptr{Float64}; CFL::Ptr{Float64}; MRefin::Ptr{Float64}; BCout::Ptr{Int32}; resmin::Ptr{Int64}
    t_ar::Ptr{Float64}; T_ar::Ptr{Float64}; KD_ar::Ptr{Float64}
    t_stride::Int64; T_stride::Int64; KD_stride::Int64; R::Float64
    coeff_up::Ptr{Float64}; coeff_do::Ptr{Float64}; coeff_stride::Int64
    C_leftB::Ptr{Float64}; C_rightB::Ptr{Float64}; nd::Int32; law::Int32
    dens_X::Ptr{Float64}; dens_T::Ptr{Float64}; dens_left::Ptr{Float64}; dens_right::Ptr{Float64}
    law_par::Ptr{Float64}         # [8,M] for law = 1 (examples/D1.jl:126-132), else C_NULL
end

struct CoupleIO                    # mbm_couple_io
    x_left::Ptr{Float64}; C_left::Ptr{Float64}; x_right::Ptr{Float64}; C_right::Ptr{Float64}
    res::Ptr{Int64}; Ri::Ptr{Float64}; t::Ptr{Float64}; dt::Ptr{Float64}; it::Ptr{Int64}
    dx1::Ptr{Float64}; dx2::Ptr{Float64}; Mass0::Ptr{Float64}; Mass::Ptr{Float64}; newton_R::Ptr{Float64}
    last::Ptr{Float64}; status::Ptr{Int32}; warn::Ptr{Int32}; idx_hash::Ptr{UInt64}
    node_steps::Ptr{Int64}; rec::Ptr{Float64}
end

struct SimpleParams                # mbm_simple_params
    Di::Ptr{Float64}; D0::Ptr{Float64}; Ea1::Ptr{Float64}; rho::Ptr{Float64}; dx::Ptr{Float64}
    t_tot::Ptr{Float64}; CFL::Ptr{Float64}; BCout::Ptr{Int32}; t_ar::Ptr{Float64}; T_ar::Ptr{Float64}
    t_stride::Int64; T_stride::Int64; R::Float64
end

struct SimpleIO                    # mbm_simple_io
    x::Ptr{Float64}; C::Ptr{Float64}; res::Ptr{Int64}; t::Ptr{Float64}; dt::Ptr{Float64}; it::Ptr{Int64}
    D::Ptr{Float64}; Mass::Ptr{Float64}; status::Ptr{Int32}; node_steps::Ptr{Int64}
end

struct ValidateArgs                # mbm_validate_args
    abi_version::Int32; n_members::Int32; ndim::Int32; cap_left::Int32; cap_right::Int32; kind::Int32
    flags::UInt32; device::Int32
    x_left::Ptr{Float64}; C_left::Ptr{Float64}; x_right::Ptr{Float64}; C_right::Ptr{Float64}
    res::Ptr{Int64}; Ri::Ptr{Float64}; rho::Ptr{Float64}; Mass0::Ptr{Float64}; Ri0::Ptr{Float64}
    C0_r::Ptr{Float64}; KD0::Ptr{Float64}; Di_r::Ptr{Float64}; V_ip::Ptr{Float64}; t_tot::Ptr{Float64}
    out::Ptr{Float64}; relerr::Ptr{Float64}
    zeta::Ptr{Float64}; C0_left::Ptr{Float64}; C0_right::Ptr{Float64}     # MBM_VAL_SPHERE (ABI 4)
end
const VAL_MASS_ERR = Int32(0); const VAL_RAYLEIGH = Int32(1); const VAL_SMITH = Int32(2); const VAL_SPHERE = Int32(3)

"""
    rayleigh_gate(x_left, C_left, x_right, C_right, res, Ri, Ri0, C0_r_end, KD0; n)

test/test_C1.jl's gate for every member of an ensemble (`[cap, M]` state arrays as the batch calls return them):
`(out[2, M], relerr[M])` with `out = (int_Ray, int_model)`; `relerr[m] <= 1e-4` is the reference's tolerance.
"""
function rayleigh_gate(x_left, C_left, x_right, C_right, res, Ri, Ri0, C0_r_end, KD0; n::Integer)
    M = size(x_left, 2); out = zeros(2, M); relerr = zeros(M)
    GC.@preserve x_left C_left x_right C_right res Ri Ri0 C0_r_end KD0 out relerr begin
        va = Ref(ValidateArgs(ABI_VERSION, M, n, size(x_left, 1), size(x_right, 1), VAL_RAYLEIGH, 0, 0,
                              pointer(x_left), pointer(C_left), pointer(x_right), pointer(C_right), pointer(res),
                              pointer(Ri), C_NULL, C_NULL, pointer(Ri0), pointer(C0_r_end), pointer(KD0),
                              C_NULL, C_NULL, C_NULL, pointer(out), pointer(relerr), C_NULL, C_NULL, C_NULL))
        rc = ccall((:mbm_validate_batch, LIB), Cint, (Ref{ValidateArgs},), va)
        rc == 0 || error("mbm_validate_batch failed with rc = $rc")
    end
    return out, relerr
end

status_string(s::Integer) = unsafe_string(ccall((:mbm_status_string, LIB), Cstring, (Int32,), Int32(s)))

"""
    couple_ensemble(pars; kind, dt_mode, max_steps = 0, n_devices = 1, record = false, rec_cap = 400_000)

`pars` is a vector of NamedTuples with the reference's parameter names (Di, D0, rho, Ri, Cl_i, Cr_i, V_ip, R,
Ea1, Ea2, t_tot, n, KD_ar, t_ar, T_ar, CFL, res, resmin, MRefin, BCout), one per member; all members share
`n`.  Mirrors the pre-loop part of DCGMB()/DCGF() (main_codes/Diff_couple_1D_general_MB.jl:35-85), runs the
time loop in the library, and returns the state in the library's (non-dimensional) units together with the scales
(`Lsc`, `tsc`) and the initial data the drivers return (`x0`, `C0`).  `record = true` keeps the per-step history
`rec[:, it, m] = (Mass, T, KD, Ri[1])` - the `Mass` vector of DCGF (`..._Flux.jl:137`).
"""
function couple_ensemble(pars::AbstractVector; kind::Symbol = :mb, dt_mode::Int32 = DT_FIND_DT,
                         max_steps::Integer = 0, n_devices::Integer = 1, cap_factor::Real = 2.0,
                         record::Bool = false, rec_cap::Integer = 400_000)
    M = length(pars)
    n = Int32(pars[1].n)
    H = length(pars[1].t_ar)
    capL = Int32(ceil(cap_factor * maximum(p.res[1] for p in pars)) + 8)
    capR = Int32(ceil(cap_factor * maximum(p.res[2] for p in pars)) + 8)
    Di = zeros(2, M); D0 = zeros(2, M); Ea = zeros(2, M); rho = ones(2, M); V = zeros(M); ttot = zeros(M)
    CFL = zeros(M); MRefin = zeros(M); BCout = zeros(Int32, 2, M); resmin = zeros(Int64, 2, M)
    t_ar = zeros(H, M); T_ar = zeros(H, M); KD_ar = zeros(H, M)
    res = zeros(Int64, 2, M); Ri = zeros(2, M); Lsc = ones(M); tsc = ones(M)
    for (m, p) in enumerate(pars)
        V_ip, t_tot, tar, Dis, D0s, Ris, Ls, Dsc, Vsc, ts = scaling(p.Ri, p.Di, p.D0, p.V_ip, p.t_tot, p.t_ar)
        Ris[1] < Ris[2] || error("Please change the size of the system. Increase Ri[2].")                    # ..._MB.jl:44-45
        p.res[1] <= p.res[2] || error("Please change the resolution of the system. res[2] >= res[1].")      # :46-47
        tv = collect(tar)
        (any(diff(tv) .<= 0.0) || any(tv .< 0.0) || any(tv .> t_tot)) &&
            error("The time array is not valid. Please check your inputs.")                                  # :38-42
        if kind === :mb
            (p.BCout[1] != 0 || p.BCout[2] != 0) &&
                error("The code is only valid for a closed system. Please set outer BC to Neumann conditions (0).")  # :73-74
        elseif p.BCout[1] != 0 && (p.n == 3 || p.n == 2)
            error("The code is only valid for cylindrical/spherical geometry, where the left outer BC has Neumann conditions (0).")
        end
        Di[:, m] .= vec(Dis); D0[:, m] .= vec(D0s); Ea[:, m] .= (p.Ea1, p.Ea2); rho[:, m] .= vec(p.rho)
        V[m] = V_ip; ttot[m] = t_tot; CFL[m] = p.CFL; MRefin[m] = p.MRefin
        BCout[:, m] .= vec(p.BCout); resmin[:, m] .= vec(p.resmin); res[:, m] .= vec(p.res); Ri[:, m] .= vec(Ris)
        t_ar[:, m] .= tv; T_ar[:, m] .= collect(p.T_ar); KD_ar[:, m] .= collect(p.KD_ar); Lsc[m] = Ls; tsc[m] = ts
    end
    x_left = zeros(capL, M); C_left = zeros(capL, M); x_right = zeros(capR, M); C_right = zeros(capR, M)
    dx1 = zeros(M); dx2 = zeros(M); nR = zeros(M); status = zeros(Int32, M)
    # create_grid! (functions.jl:364-370), batched
    rc = ccall((:mbm_create_grid, LIB), Cint,
               (Int32, Int32, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Int32}), M, capL, capR, Ri, res, MRefin, x_left, x_right, dx1, dx2, nR, status)
    rc == 0 || error("mbm_create_grid failed rc=$rc")
    for (m, p) in enumerate(pars)                             # initial condition, ..._MB.jl:54-55
        C_left[1:p.res[1], m] .= p.Cl_i
        C_right[1:p.res[2], m] .= p.Cr_i
    end
    x0 = [vcat(x_left[1:res[1, m], m], x_right[1:res[2, m], m]) for m in 1:M]       # x0 of create_grid!, C0 of :56
    C0 = [vcat(C_left[1:res[1, m], m], C_right[1:res[2, m], m]) for m in 1:M]
    Ri0 = copy(Ri)
    Mass0 = zeros(M)
    rc = ccall((:mbm_calc_mass_vol, LIB), Cint,
               (Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64},
                Ptr{Float64}, Ptr{Float64}), M, n, capL, capR, x_left, x_right, C_left, C_right, res, rho, Mass0)
    rc == 0 || error("mbm_calc_mass_vol failed rc=$rc")
    t = zeros(M); dt = zeros(M); it = zeros(Int64, M); Mass = zeros(M); last = zeros(8, M)
    warn = zeros(Int32, M); hash = zeros(UInt64, M); node_steps = zeros(Int64, M)
    rec = record ? zeros(4, rec_cap, M) : zeros(0, 0, 0)
    flags = record ? (FLAG_MASS_FINAL | FLAG_RECORD) : FLAG_MASS_FINAL
    inner = kind === :mb ? BC_MB : kind === :flux ? BC_FLUX : kind === :lasaga ? BC_LASAGA : BC_HOLD
    GC.@preserve Di D0 Ea rho V ttot CFL MRefin BCout resmin t_ar T_ar KD_ar x_left C_left x_right C_right res Ri t dt it dx1 dx2 Mass0 Mass nR last status warn hash node_steps rec begin
        c = Ref(Common(ABI_VERSION, M, n, H, capL, capR, inner, dt_mode, flags, Int32(n_devices), max_steps, C_NULL,
                       record ? rec_cap : 0))
        p = Ref(CoupleParams(pointer(Di), pointer(D0), pointer(Ea), pointer(rho), pointer(V), pointer(ttot), pointer(CFL),
                             pointer(MRefin), pointer(BCout), pointer(resmin), pointer(t_ar), pointer(T_ar), pointer(KD_ar),
                             H, H, H, pars[1].R, C_NULL, C_NULL, 0, C_NULL, C_NULL, 0, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL))
        io = Ref(CoupleIO(pointer(x_left), pointer(C_left), pointer(x_right), pointer(C_right), pointer(res), pointer(Ri),
                          pointer(t), pointer(dt), pointer(it), pointer(dx1), pointer(dx2), pointer(Mass0), pointer(Mass),
                          pointer(nR), pointer(last), pointer(status), pointer(warn), pointer(hash), pointer(node_steps),
                          record ? pointer(rec) : Ptr{Float64}(C_NULL)))
        rc = ccall((:mbm_couple_batch, LIB), Cint, (Ref{Common}, Ref{CoupleParams}, Ref{CoupleIO}), c, p, io)
        rc == 0 || error("mbm_couple_batch failed rc=$rc")
    end
    return (; x_left, x_right, C_left, C_right, res, Ri, Ri0, t, it, dx1, dx2, Mass0, Mass, status, warn, last, Lsc, tsc,
            x0, C0, rec)
end

# one member of a couple ensemble in the reference's return conventions (lengths back in [m] as `rescale` does,
# functions.jl:941-955; concentrations and dx1 / dx2 are returned as the drivers return them: unscaled)
function couple_member(o, m::Integer)
    o.status[m] == 0 || error(status_string(o.status[m]))            # the reference throws at the same sites
    n1, n2 = o.res[1, m], o.res[2, m]
    L = o.Lsc[m]
    return (; x_left = o.x_left[1:n1, m] .* L, x_right = o.x_right[1:n2, m] .* L, dx1 = o.dx1[m], dx2 = o.dx2[m],
            x0 = o.x0[m] .* L, res = reshape(o.res[:, m], 1, 2), Ri = reshape(o.Ri[:, m] .* L, 1, 2),
            Ri0 = reshape(o.Ri0[:, m] .* L, 1, 2), C_left = o.C_left[1:n1, m], C_right = o.C_right[1:n2, m], C0 = o.C0[m],
            Mass = size(o.rec, 1) == 4 ? o.rec[1, 1:min(o.it[m], size(o.rec, 2)), m] : Float64[], Mass0 = o.Mass0[m])
end

"DCGMB(): drop-in for main_codes/Diff_couple_1D_general_MB.jl:4-140 (same parameter block, same return tuple)"
function DCGMB(; Di = [4e-10 5e-8], D0 = [1e-4 5e-4], rho = [1.0 1.0], Ri = [0.2 0.8], Cl_i = 0.5, Cr_i = 0.7,
               V_ip = -1e-7, R = 8.314472, Ea1 = 292879.6767, Ea2 = 300660.4018, t_tot = 1e4, n = 2,
               KD_ar = LinRange(0.5 / 0.7, 0.5 / 0.7 * 0.9, 1000), t_ar = LinRange(0.0, 1e4, 1000),
               T_ar = LinRange(973.15, 723.15, 1000), CFL = 0.5, res = [100 150], resmin = copy(res), MRefin = 2.0,
               BCout = [0 0])
    p = (; Di, D0, rho, Ri, Cl_i, Cr_i, V_ip, R, Ea1, Ea2, t_tot, n, KD_ar, t_ar, T_ar, CFL, res, resmin, MRefin, BCout)
    r = couple_member(couple_ensemble([p]; kind = :mb), 1)
    return r.x_left, r.x_right, r.dx1, r.dx2, r.x0, r.res, r.Ri, r.C_left, r.C_right, r.C0
end

"DCGF(): drop-in for main_codes/Diff_couple_1D_general_Flux.jl:4-138 (pass that file's literals; returns its 12-tuple incl. the `Mass` history)"
function DCGF(; Di, D0, rho, Ri, Cl_i, Cr_i, V_ip, R, Ea1, Ea2, t_tot, n, KD_ar, t_ar, T_ar, CFL, res, resmin = copy(res),
              MRefin, BCout, rec_cap::Integer = 400_000)
    p = (; Di, D0, rho, Ri, Cl_i, Cr_i, V_ip, R, Ea1, Ea2, t_tot, n, KD_ar, t_ar, T_ar, CFL, res, resmin, MRefin, BCout)
    r = couple_member(couple_ensemble([p]; kind = :flux, record = true, rec_cap = rec_cap), 1)
    return r.x_left, r.x_right, r.dx1, r.dx2, r.x0, r.res, r.Ri, r.C_left, r.C_right, r.C0, r.Mass, r.Mass0
end

"""
    simple_ensemble(pars; max_steps = 0, n_devices = 1)

Single-phase members (main_codes/Simple_Diff_1D_general.jl:6-66): NamedTuples with `Di, D0, L, rho, R, Ea1, t_tot, n, res, CFL,
BCout, T_ar, t_ar` and the initial profile `C` (length `res`) on the driver's mesh `x = [0:dx:L;]`.
"""
function simple_ensemble(pars::AbstractVector; max_steps::Integer = 0, n_devices::Integer = 1)
    M = length(pars); n = Int32(pars[1].n); H = length(pars[1].t_ar)
    cap = Int32(maximum(p.res for p in pars))
    Di = zeros(M); D0 = zeros(M); Ea1 = zeros(M); rho = zeros(M); dx = zeros(M); ttot = zeros(M); CFL = zeros(M)
    BCout = zeros(Int32, 2, M); t_ar = zeros(H, M); T_ar = zeros(H, M)
    x = zeros(cap, M); C = zeros(cap, M); res = zeros(Int64, M)
    for (m, p) in enumerate(pars)
        (p.BCout[1] != 0 && (p.n == 3 || p.n == 2)) &&
            error("The code is only valid for cylindrical/spherical geometry, where the left outer BC has Neumann conditions (0).")
        tv = collect(p.t_ar)
        (any(diff(tv) .<= 0.0) || any(tv .< 0.0) || any(tv .> p.t_tot)) && error("The time array is not valid. Please check your inputs.")
        xs = [0:(p.L * inv(p.res - 1)):p.L;]                      # :22-23
        length(xs) == p.res || error("mesh of $(length(xs)) nodes for res = $(p.res)")
        x[1:p.res, m] .= xs; C[1:p.res, m] .= vec(p.C); res[m] = p.res
        Di[m] = p.Di; D0[m] = p.D0; Ea1[m] = p.Ea1; rho[m] = p.rho; dx[m] = p.L * inv(p.res - 1.0)   # :45
        ttot[m] = p.t_tot; CFL[m] = p.CFL; BCout[:, m] .= vec(p.BCout); t_ar[:, m] .= tv; T_ar[:, m] .= collect(p.T_ar)
    end
    t = zeros(M); dt = zeros(M); it = zeros(Int64, M); D = zeros(M); Mass = zeros(M); status = zeros(Int32, M)
    node_steps = zeros(Int64, M)
    GC.@preserve Di D0 Ea1 rho dx ttot CFL BCout t_ar T_ar x C res t dt it D Mass status node_steps begin
        c = Ref(Common(ABI_VERSION, M, n, H, cap, 0, 0, DT_SIMPLE, FLAG_MASS_FINAL, Int32(n_devices), max_steps, C_NULL, 0))
        sp = Ref(SimpleParams(pointer(Di), pointer(D0), pointer(Ea1), pointer(rho), pointer(dx), pointer(ttot), pointer(CFL),
                              pointer(BCout), pointer(t_ar), pointer(T_ar), H, H, pars[1].R))
        io = Ref(SimpleIO(pointer(x), pointer(C), pointer(res), pointer(t), pointer(dt), pointer(it), pointer(D), pointer(Mass),
                          pointer(status), pointer(node_steps)))
        rc = ccall((:mbm_simple_diff_batch, LIB), Cint, (Ref{Common}, Ref{SimpleParams}, Ref{SimpleIO}), c, sp, io)
        rc == 0 || error("mbm_simple_diff_batch failed rc=$rc")
    end
    return (; x, C, res, t, it, D, Mass, status)
end

"SDG(): drop-in for main_codes/Simple_Diff_1D_general.jl:4-103 -> (x, C, x0, C0, D, t, t_tot, Mass, Mass0)"
function SDG(; Di = 2.65 * 1e-18, D0 = 2.75 * 1e-6, L = 0.001, Cinf = 0.0, Cstart = 4.0, rho = 2700.0, R = 8.314472,
             Ea1 = 292880.0, t_tot = 1e-4 * (60 * 60 * 24 * 365.25 * 1e6), n = 1, res = 100, CFL = 0.99, BCout = [1, 1],
             T_ar = LinRange(1273.15, 923.15, 1000), t_ar = LinRange(0.0, t_tot, 1000))
    C = Cinf * ones(res); C[1] = Cstart                            # :27-28
    x0 = [0:(L * inv(res - 1)):L;]; C0 = copy(C)
    Mass0 = calc_mass_vol_simple_diff(x0, C0, n, rho)               # :54
    o = simple_ensemble([(; Di, D0, L, rho, R, Ea1, t_tot, n, res, CFL, BCout, T_ar, t_ar, C)])
    o.status[1] == 0 || error(status_string(o.status[1]))
    return o.x[1:res, 1], o.C[1:res, 1], x0, C0, o.D[1], o.t[1], t_tot, o.Mass[1], Mass0
end

"""
    stefan_ensemble(pars; max_steps = 0, n_devices = 1, rec_cap = 200_000)

Members of the thermodynamically constrained Stefan problem (main_codes/Chemical_Stefan_problem_XT.jl:8-98): NamedTuples with
`Di, Ri, Tstart, Tstop, t_tot, n, CompInt, coeff (3 x 2), CFL, res, MRefin, BCout`.  The density table is replaced by ones, as
examples/D1.jl:54-55 does (the shipped .tab file is not part of the repository); a table goes into `nd`, `dens_X`, `dens_T`,
`dens_left`, `dens_right` of `CoupleParams` (calculate_density, functionsStefan.jl:23-31) - not wired here.
"""
function stefan_ensemble(pars::AbstractVector; max_steps::Integer = 0, n_devices::Integer = 1, rec_cap::Integer = 200_000,
                         hist_len::Integer = 10_000, cap_factor::Real = 4.0)
    M = length(pars); n = Int32(pars[1].n); H = hist_len
    capL = Int32(ceil(cap_factor * maximum(p.res[1] for p in pars)) + 8)
    capR = Int32(ceil(cap_factor * maximum(p.res[2] for p in pars)) + 8)
    Di = zeros(2, M); D0 = zeros(2, M); Ea = zeros(2, M); rho = ones(2, M); V = zeros(M); ttot = zeros(M)
    CFL = zeros(M); MRefin = zeros(M); BCout = zeros(Int32, 2, M); resmin = zeros(Int64, 2, M)
    t_ar = zeros(H, M); T_ar = zeros(H, M); KD_ar = zeros(H, 1)
    cu = zeros(3, M); cd = zeros(3, M); CLB = zeros(M); CRB = zeros(M)
    res = zeros(Int64, 2, M); Ri = zeros(2, M)
    for (m, p) in enumerate(pars)
        eq_values = [p.coeff[1, 1] p.coeff[2, 1] p.coeff[3, 1]; p.coeff[1, 2] p.coeff[2, 2] p.coeff[3, 2]]      # :17-19
        coeff_up, coeff_do = coeff_trans_line(eq_values)                                                          # :29
        C_leftB, C_rightB = composition(coeff_up, coeff_do, p.Tstart)                                             # :56
        Xc = (1.0 - p.CompInt) * C_rightB + p.CompInt * C_leftB                                                   # :57
        L = (p.Ri ^ p.n * (Xc - C_leftB) * inv((C_rightB - Xc)) ^ (1 * inv(p.n))) + p.Ri                          # :58
        p.Ri < L || error("Please change the size of the system. Increase Ri[2].")
        p.res[1] <= p.res[2] || error("Please change the resolution of the system. res[2] >= res[1].")
        cu[:, m] .= vec(coeff_up); cd[:, m] .= vec(coeff_do); CLB[m] = C_leftB; CRB[m] = C_rightB
        Ri[:, m] .= (p.Ri, L); res[:, m] .= vec(p.res); resmin[:, m] .= vec(p.res); Di[:, m] .= vec(p.Di)
        ttot[m] = p.t_tot; CFL[m] = p.CFL; MRefin[m] = p.MRefin; BCout[:, m] .= vec(p.BCout)
        T_ar[:, m] .= collect(LinRange(p.Tstart, p.Tstop, H)); t_ar[:, m] .= collect(LinRange(0, p.t_tot + 1e-10, H))   # :50-51
    end
    x_left = zeros(capL, M); C_left = zeros(capL, M); x_right = zeros(capR, M); C_right = zeros(capR, M)
    dx1 = zeros(M); dx2 = zeros(M); nR = zeros(M); status = zeros(Int32, M)
    rc = ccall((:mbm_create_grid, LIB), Cint,
               (Int32, Int32, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Int32}), M, capL, capR, Ri, res, MRefin, x_left, x_right, dx1, dx2, nR, status)
    rc == 0 || error("mbm_create_grid failed rc=$rc")
    dt = zeros(M)
    for (m, p) in enumerate(pars)
        C_left[1:p.res[1], m] .= CLB[m]; C_right[1:p.res[2], m] .= CRB[m]                                        # :76-77
        dt[m] = minimum([dx1[m], dx2[m]]) ^ 2 * inv(maximum(p.Di))                                                # :79
    end
    x0 = [vcat(x_left[1:res[1, m], m], x_right[1:res[2, m], m]) for m in 1:M]
    C0 = [vcat(C_left[1:res[1, m], m], C_right[1:res[2, m], m]) for m in 1:M]
    Mass0 = zeros(M)
    rc = ccall((:mbm_calc_mass_vol, LIB), Cint,
               (Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64},
                Ptr{Float64}, Ptr{Float64}), M, n, capL, capR, x_left, x_right, C_left, C_right, res, rho, Mass0)
    rc == 0 || error("mbm_calc_mass_vol failed rc=$rc")
    t = zeros(M); it = zeros(Int64, M); Mass = zeros(M); last = zeros(8, M)
    warn = zeros(Int32, M); hash = zeros(UInt64, M); node_steps = zeros(Int64, M); rec = zeros(4, rec_cap, M)
    GC.@preserve Di D0 Ea rho V ttot CFL MRefin BCout resmin t_ar T_ar KD_ar cu cd CLB CRB x_left C_left x_right C_right res Ri t dt it dx1 dx2 Mass0 Mass nR last status warn hash node_steps rec begin
        c = Ref(Common(ABI_VERSION, M, n, H, capL, capR, BC_STEFAN, DT_STEFAN, FLAG_MASS_FINAL | FLAG_RECORD, Int32(n_devices),
                       max_steps, C_NULL, rec_cap))
        p = Ref(CoupleParams(pointer(Di), pointer(D0), pointer(Ea), pointer(rho), pointer(V), pointer(ttot), pointer(CFL),
                             pointer(MRefin), pointer(BCout), pointer(resmin), pointer(t_ar), pointer(T_ar), pointer(KD_ar),
                             H, H, 0, 8.314472, pointer(cu), pointer(cd), 3, pointer(CLB), pointer(CRB), 0, 0,
                             C_NULL, C_NULL, C_NULL, C_NULL, C_NULL))
        io = Ref(CoupleIO(pointer(x_left), pointer(C_left), pointer(x_right), pointer(C_right), pointer(res), pointer(Ri),
                          pointer(t), pointer(dt), pointer(it), pointer(dx1), pointer(dx2), pointer(Mass0), pointer(Mass),
                          pointer(nR), pointer(last), pointer(status), pointer(warn), pointer(hash), pointer(node_steps),
                          pointer(rec)))
        rc = ccall((:mbm_stefan_xt_batch, LIB), Cint, (Ref{Common}, Ref{CoupleParams}, Ref{CoupleIO}), c, p, io)
        rc == 0 || error("mbm_stefan_xt_batch failed rc=$rc")
    end
    return (; x_left, x_right, C_left, C_right, res, Ri, t, it, dx1, dx2, Mass0, Mass, status, last, x0, C0, rec)
end

"""
CSP(): drop-in for main_codes/Chemical_Stefan_problem_XT.jl:4-208 with the density table replaced by ones (examples/D1.jl:54-55).
Returns the entries of its 20-tuple that the time loop produces, under the reference's names:
`(x_left, x_right, x0, C_left, C_right, C0, maxC, T, Tstart, Tstop, KD_sim, T_sim, Mass0, Mass)`; the plot-only entries
(`Tlin, XC_left, XC_right, KDlin`: the phase diagram, :31-35,47-49; `Mass01, Mass2`: the plot normalisation of the mass) are
host-side one-liners of the reference and are not repeated here.
"""
function CSP(; Di = [0.002 0.0004;], Ri = 2.0, Tstart = 1400.0 + 273.0, Tstop = 1300.0 + 273.0, t_tot = 1e4, n = 1,
             CompInt = 0.3, coeff, CFL = 0.8, res = [15 25;], MRefin = 5.0, BCout = [0 0], rec_cap::Integer = 200_000)
    o = stefan_ensemble([(; Di, Ri, Tstart, Tstop, t_tot, n, CompInt, coeff, CFL, res, MRefin, BCout)]; rec_cap = rec_cap)
    o.status[1] == 0 || error(status_string(o.status[1]))
    n1, n2 = o.res[1, 1], o.res[2, 1]
    k = min(o.it[1], rec_cap)
    C_left = o.C_left[1:n1, 1]; C_right = o.C_right[1:n2, 1]
    maxC = maximum([maximum(C_left), maximum(C_right)])
    return (o.x_left[1:n1, 1], o.x_right[1:n2, 1], o.x0[1], C_left, C_right, o.C0[1], maxC, o.last[4, 1], Tstart, Tstop,
            o.rec[3, 1:k, 1], o.rec[2, 1:k, 1], o.Mass0[1], o.rec[1, 1:k, 1])
end

end # module
