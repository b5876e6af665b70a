# MovingBoundaryMineralsB200.jl — ccall shim over libmbm_b200.so (include/mbm_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It is the binding a maintainer of
# MovingBoundaryMinerals.jl would add; struct layouts mirror include/mbm_b200.h field by field (the same
# layouts are checked against the C header for the Python ctypes mirror in tests/test_abi.py).
#
# Usage (drop-in for main_codes/Diff_couple_1D_general_MB.jl):
#     using MovingBoundaryMinerals, MovingBoundaryMineralsB200
#     x_left, x_right, dx1, dx2, x0, res, Ri, C_left, C_right, C0 = MovingBoundaryMineralsB200.DCGMB()
# and the batched form:
#     out = MovingBoundaryMineralsB200.couple_ensemble(pars::Vector{<:NamedTuple}; kind = :mb)
module MovingBoundaryMineralsB200

using MovingBoundaryMinerals: scaling, rescale            # host-side pre/post steps stay the reference's own

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
    couple_ensemble(pars; kind, dt_mode, max_steps = 0, n_devices = 1)

`pars` is a vector of NamedTuples with the reference's parameter names (Di, D0, rho, Ri, Cl_i, Cr_i, V_ip, R,
Ea1, Ea2, t_tot, n, KD_ar, t_ar, T_ar, CFL, res, resmin, MRefin, BCout), one per member; all members share
`n`.  Mirrors the pre-loop part of DCGMB()/DCGF() (main_codes/Diff_couple_1D_general_MB.jl:35-85), runs the
time loop in the library, rescales like :134.
"""
function couple_ensemble(pars::AbstractVector; kind::Symbol = :mb, dt_mode::Int32 = DT_FIND_DT,
                         max_steps::Integer = 0, n_devices::Integer = 1, cap_factor::Real = 2.0)
    M = length(pars)
    n = Int32(pars[1].n)
    H = length(pars[1].t_ar)
    capL = Int32(ceil(cap_factor * maximum(p.res[1] for p in pars)) + 8)
    capR = Int32(ceil(cap_factor * maximum(p.res[2] for p in pars)) + 8)
    Di = zeros(2, M); D0 = zeros(2, M); Ea = zeros(2, M); rho = ones(2, M); V = zeros(M); ttot = zeros(M)
    CFL = zeros(M); MRefin = zeros(M); BCout = zeros(Int32, 2, M); resmin = zeros(Int64, 2, M)
    t_ar = zeros(H, M); T_ar = zeros(H, M); KD_ar = zeros(H, M)
    res = zeros(Int64, 2, M); Ri = zeros(2, M); Lsc = ones(M)
    for (m, p) in enumerate(pars)
        V_ip, t_tot, tar, Dis, D0s, Ris, Ls, Dsc, Vsc, tsc = scaling(p.Ri, p.Di, p.D0, p.V_ip, p.t_tot, p.t_ar)
        Di[:, m] .= vec(Dis); D0[:, m] .= vec(D0s); Ea[:, m] .= (p.Ea1, p.Ea2); rho[:, m] .= vec(p.rho)
        V[m] = V_ip; ttot[m] = t_tot; CFL[m] = p.CFL; MRefin[m] = p.MRefin
        BCout[:, m] .= vec(p.BCout); resmin[:, m] .= vec(p.resmin); res[:, m] .= vec(p.res); Ri[:, m] .= vec(Ris)
        t_ar[:, m] .= collect(tar); T_ar[:, m] .= collect(p.T_ar); KD_ar[:, m] .= collect(p.KD_ar); Lsc[m] = Ls
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
    Mass0 = zeros(M)
    rc = ccall((:mbm_calc_mass_vol, LIB), Cint,
               (Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64},
                Ptr{Float64}, Ptr{Float64}), M, n, capL, capR, x_left, x_right, C_left, C_right, res, rho, Mass0)
    rc == 0 || error("mbm_calc_mass_vol failed rc=$rc")
    t = zeros(M); dt = zeros(M); it = zeros(Int64, M); Mass = zeros(M); last = zeros(8, M)
    warn = zeros(Int32, M); hash = zeros(UInt64, M); node_steps = zeros(Int64, M)
    inner = kind === :mb ? BC_MB : kind === :flux ? BC_FLUX : kind === :lasaga ? BC_LASAGA : BC_HOLD
    GC.@preserve Di D0 Ea rho V ttot CFL MRefin BCout resmin t_ar T_ar KD_ar x_left C_left x_right C_right res Ri t dt it dx1 dx2 Mass0 Mass nR last status warn hash node_steps begin
        c = Ref(Common(ABI_VERSION, M, n, H, capL, capR, inner, dt_mode, FLAG_MASS_FINAL, Int32(n_devices), max_steps, C_NULL, 0))
        p = Ref(CoupleParams(pointer(Di), pointer(D0), pointer(Ea), pointer(rho), pointer(V), pointer(ttot), pointer(CFL),
                             pointer(MRefin), pointer(BCout), pointer(resmin), pointer(t_ar), pointer(T_ar), pointer(KD_ar),
                             H, H, H, pars[1].R, C_NULL, C_NULL, 0, C_NULL, C_NULL, 0, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL))
        io = Ref(CoupleIO(pointer(x_left), pointer(C_left), pointer(x_right), pointer(C_right), pointer(res), pointer(Ri),
                          pointer(t), pointer(dt), pointer(it), pointer(dx1), pointer(dx2), pointer(Mass0), pointer(Mass),
                          pointer(nR), pointer(last), pointer(status), pointer(warn), pointer(hash), pointer(node_steps), C_NULL))
        rc = ccall((:mbm_couple_batch, LIB), Cint, (Ref{Common}, Ref{CoupleParams}, Ref{CoupleIO}), c, p, io)
        rc == 0 || error("mbm_couple_batch failed rc=$rc")
    end
    return (; x_left, x_right, C_left, C_right, res, Ri, t, it, dx1, dx2, Mass0, Mass, status, Lsc)
end

"DCGMB(): drop-in for main_codes/Diff_couple_1D_general_MB.jl:4-140 (same parameter block, same return tuple)"
function DCGMB(; Di = [4e-10 5e-8], D0 = [1e-4 5e-4], rho = [1.0 1.0], Ri = [0.2 0.8], Cl_i = 0.5, Cr_i = 0.7,
               V_ip = -1e-7, R = 8.314472, Ea1 = 292879.6767, Ea2 = 300660.4018, t_tot = 1e4, n = 2,
               KD_ar = LinRange(0.5 / 0.7, 0.5 / 0.7 * 0.9, 1000), t_ar = LinRange(0.0, 1e4, 1000),
               T_ar = LinRange(973.15, 723.15, 1000), CFL = 0.5, res = [100 150], resmin = copy(res), MRefin = 2.0,
               BCout = [0 0])
    p = (; Di, D0, rho, Ri, Cl_i, Cr_i, V_ip, R, Ea1, Ea2, t_tot, n, KD_ar, t_ar, T_ar, CFL, res, resmin, MRefin, BCout)
    o = couple_ensemble([p]; kind = :mb)
    o.status[1] == 0 || error(status_string(o.status[1]))
    n1, n2 = o.res[1, 1], o.res[2, 1]
    L = o.Lsc[1]
    x_left = o.x_left[1:n1, 1] .* L; x_right = o.x_right[1:n2, 1] .* L
    return x_left, x_right, o.dx1[1], o.dx2[1], nothing, o.res[:, 1], o.Ri[:, 1] .* L, o.C_left[1:n1, 1], o.C_right[1:n2, 1], nothing
end

end # module
