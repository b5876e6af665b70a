"""Analytical / semi-analytical solutions used by the reference's known-answer tests, restated in
numpy/scipy (test harness only).  Follows reference src/Benchmarks.jl line by line, including
the as-shipped `LinearIndices` quirk of rayleigh_fractionation (:250-258) that the shipped
tolerances depend on."""
import numpy as np
from scipy.special import erfc


def isapprox(a, b, rtol):
    """Julia isapprox for arrays: ||a-b||_2 <= rtol*max(||a||_2, ||b||_2); returns (ok, relerr)."""
    a = np.asarray(a, float).ravel()
    b = np.asarray(b, float).ravel()
    den = max(np.linalg.norm(a), np.linalg.norm(b))
    err = np.linalg.norm(a - b) / den if den > 0 else 0.0
    return err <= rtol, err


def trapezoidal_integration(x, fx):
    """src/functions.jl:1310-1323"""
    x = np.asarray(x, float).ravel()
    fx = np.asarray(fx, float).ravel()
    dx = x[1:] - x[:-1]
    return np.cumsum(0.5 * dx * (fx[:-1] + fx[1:]))[-1]


def linear_interpolation_1D(x, y, xi):
    """src/functions.jl:565-585"""
    x = np.asarray(x, float).ravel()
    y = np.asarray(y, float).ravel()
    if xi < x.min() or xi > x.max():
        raise ValueError("x_interp is out of bounds of x.")
    for i in range(len(x) - 1):
        if x[i] <= xi <= x[i + 1]:
            N2 = (xi - x[i]) * (1.0 / (x[i + 1] - x[i]))
            return (1 - N2) * y[i] + N2 * y[i + 1]
    return None


def calc_sinus_sphere(x0, C0, D, tot, nterms):
    """src/Benchmarks.jl:82-99 (x[1] = 0 gives NaN/Inf in element 1, overwritten at :96)"""
    C = np.array(C0, float).ravel().copy()
    x = np.array(x0, float).ravel().copy()
    C1 = C[0]
    Cs = C[-1]
    L = x[-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        for nt in range(1, nterms + 1):
            e = np.exp(-D * nt**2 * np.pi**2 * tot * (1.0 / L**2))
            C = C + 2.0 * L * (1.0 / (np.pi * x)) * ((-1.0) ** nt) * (1.0 / nt) * np.sin(nt * np.pi * x * (1.0 / L)) * e
            C1 = C1 + 2.0 * (((-1.0) ** nt) * e)
    C = C + 1.0
    C[0] = C1 + 1.0
    C[-1] = Cs
    return x, C


def crank_time_transformation3(C0, x0, T0, T, E, R, D0, t, nterms):
    """src/Benchmarks.jl:192-212"""
    Tan = np.linspace(T0, T, 1000)
    tan = np.linspace(0, t, 1000)
    Dan = D0 * np.exp(-E * (1.0 / (R * Tan)))
    zeta = trapezoidal_integration(tan, Dan)
    C = np.array(C0, float).ravel().copy()
    x = np.array(x0, float).ravel().copy()
    C1 = C[0]
    Cs = C[-1]
    L = x[-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        for nt in range(1, nterms + 1):
            e = np.exp(-zeta * nt**2 * np.pi**2 * (1.0 / L**2))
            C = C + 2.0 * L * (1.0 / (np.pi * x)) * ((-1.0) ** nt) * (1.0 / nt) * np.sin(nt * np.pi * x * (1.0 / L)) * e
            C1 = C1 + 2.0 * (((-1.0) ** nt) * e)
    C = C + 1.0
    C[0] = C1 + 1.0
    C[-1] = Cs
    return C, x


def rayleigh_fractionation(x_left, C_left, Ri0, Ri, C0_r, KD0, n):
    """src/Benchmarks.jl:239-260, with `x_left[LinearIndices(indi)]` == x_left[1:length(indi)] as shipped"""
    x_left = np.asarray(x_left, float).ravel()
    C_left = np.asarray(C_left, float).ravel()
    Ray_Fs = np.linspace(Ri0[0] * (1.0 / Ri0[1]), Ri[0] * (1.0 / Ri0[1]), 1000) ** n
    Ray_Fl = 1.0 - Ray_Fs
    Ray_Cl = Ray_Fl ** (KD0 - 1.0) * C0_r[-1]
    Ray_Cs = Ray_Cl * KD0
    k = int(np.count_nonzero(x_left > Ri0[0]))
    x_l_p = np.concatenate([[Ri0[0]], x_left[:k]])            # the quirk: first k nodes, not the selected ones
    phi_solid = x_l_p * (1.0 / Ri0[1])
    C_bc_p = linear_interpolation_1D(x_left, C_left, Ri0[0])
    C_l_p = np.concatenate([[C_bc_p], C_left[:k]])
    return Ray_Fs, Ray_Fl, Ray_Cl, Ray_Cs, C_l_p, phi_solid


def smith(x_right, C_right, Ri, Di, t_tot, KD, V_ip, n):
    """src/Benchmarks.jl:284-301"""
    assert n == 1
    xan = np.asarray(x_right, float) - Ri[0]
    Dan, tan, Kan, R = Di[1], t_tot, KD, V_ip
    ksi = 0.5 * (Dan * tan) ** (-0.5)
    Can0 = C_right[-1]
    Can = (1.0 + ((1.0 - Kan) * (1.0 / (2.0 * Kan))) * np.exp(-R * (1.0 / Dan) * xan) * erfc(ksi * (xan - R * tan))
           - 0.5 * erfc(ksi * (xan + R * tan))
           + ((1.0 - Kan) * (1.0 / 2.0)) * (1.0 * (1.0 / (1.0 - Kan)) - 1.0 * (1.0 / Kan))
           * np.exp(-(1.0 - Kan) * (R * (1.0 / Dan)) * (xan + Kan * R * tan))
           * erfc(ksi * (xan + (2.0 * Kan - 1.0) * R * tan)))
    return xan, Can * Can0
