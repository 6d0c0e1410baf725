"""Spin-resolved energy expressions of the supported functionals in torch float64 (test helper).

Third, independent statement of the functional arithmetic: the oracle (oracle/oracle_functionals.c) uses
closed-shell formulas with hand-derived derivatives, the device code (serenity_b200/csrc/functionals.cuh) uses
forward-mode AD of spin-resolved expressions, and this file lets torch.autograd differentiate the published
spin-resolved energy densities F(rho_a, rho_b, s_aa, s_ab, s_bb) (XCFun parametrisation).
"""
import math

import torch

PI = math.pi
CF = 0.3 * (3.0 * PI * PI) ** (2.0 / 3.0)


def _f_zeta(z):
    return ((1 + z) ** (4.0 / 3.0) + (1 - z) ** (4.0 / 3.0) - 2.0) / (2.0 ** (4.0 / 3.0) - 2.0)


FPP0 = 4.0 / (9.0 * (2.0 ** (1.0 / 3.0) - 1.0))  # f''(0) = 1.709921


def slaterx(a, b, gaa, gab, gbb):
    return -0.75 * (6.0 / PI) ** (1.0 / 3.0) * (a ** (4.0 / 3.0) + b ** (4.0 / 3.0))


def _vwn_eps(x, A, x0, b, c):
    Q = math.sqrt(4 * c - b * b)
    X = x * x + b * x + c
    X0 = x0 * x0 + b * x0 + c
    at = torch.atan(Q / (2 * x + b))
    return A * (torch.log(x * x / X) + 2 * b / Q * at
                - b * x0 / X0 * (torch.log((x - x0) ** 2 / X) + 2 * (b + 2 * x0) / Q * at))


def vwn5c(a, b, gaa, gab, gbb):
    n = a + b
    z = (a - b) / n
    rs = (3.0 / (4.0 * PI * n)) ** (1.0 / 3.0)
    x = torch.sqrt(rs)
    eP = _vwn_eps(x, 0.0310907, -0.10498, 3.72744, 12.9352)
    eF = _vwn_eps(x, 0.01554535, -0.32500, 7.06042, 18.0578)
    ac = _vwn_eps(x, -1.0 / (6.0 * PI * PI), -0.0047584, 1.13107, 13.0045)
    fz = _f_zeta(z)
    return n * (eP + ac * fz / FPP0 * (1 - z ** 4) + (eF - eP) * fz * z ** 4)


def tfk(a, b, gaa, gab, gbb):
    return 2.0 ** (2.0 / 3.0) * CF * (a ** (5.0 / 3.0) + b ** (5.0 / 3.0))


def _pbex_spin(n, g):  # E_x[n] for a closed-shell density n with |grad n|^2 = g
    kappa, mu = 0.804, 0.2195149727645171
    s2 = g / (4.0 * (3.0 * PI * PI) ** (2.0 / 3.0) * n ** (8.0 / 3.0))
    Fx = 1 + kappa - kappa / (1 + mu * s2 / kappa)
    return -0.75 * (3.0 / PI) ** (1.0 / 3.0) * n ** (4.0 / 3.0) * Fx


def pbex(a, b, gaa, gab, gbb):
    return 0.5 * (_pbex_spin(2 * a, 4 * gaa) + _pbex_spin(2 * b, 4 * gbb))


def _b88_corr_spin(r, g):
    beta = 0.0042
    x = torch.sqrt(g) / r ** (4.0 / 3.0)
    return -beta * r ** (4.0 / 3.0) * x * x / (1 + 6 * beta * x * torch.asinh(x))


def beckecorrx(a, b, gaa, gab, gbb):
    return _b88_corr_spin(a, gaa) + _b88_corr_spin(b, gbb)


def beckex(a, b, gaa, gab, gbb):
    return slaterx(a, b, gaa, gab, gbb) + beckecorrx(a, b, gaa, gab, gbb)


def lypc(a, b, gaa, gab, gbb):
    A, B, C, D = 0.04918, 0.132, 0.2533, 0.349
    n = a + b
    g = gaa + 2 * gab + gbb
    q = n ** (-1.0 / 3.0)
    omega = torch.exp(-C * q) * n ** (-11.0 / 3.0) / (1 + D * q)
    delta = C * q + D * q / (1 + D * q)
    t = (a * b * (2.0 ** (11.0 / 3.0) * CF * (a ** (8.0 / 3.0) + b ** (8.0 / 3.0))
                  + (47.0 / 18.0 - 7.0 * delta / 18.0) * g - (2.5 - delta / 18.0) * (gaa + gbb)
                  - (delta - 11.0) / 9.0 * (a * gaa + b * gbb) / n)
         - 2.0 / 3.0 * n * n * g + (2.0 / 3.0 * n * n - a * a) * gbb + (2.0 / 3.0 * n * n - b * b) * gaa)
    return -A * 4 * a * b / ((1 + D * q) * n) - A * B * omega * t


def _pw92_G(rs, A, a1, b1, b2, b3, b4):
    return -2 * A * (1 + a1 * rs) * torch.log(1 + 1 / (2 * A * (b1 * rs ** 0.5 + b2 * rs + b3 * rs ** 1.5 + b4 * rs ** 2)))


def _pw92_eps(rs, z):
    e0 = _pw92_G(rs, 0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294)
    e1 = _pw92_G(rs, 0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517)
    mac = _pw92_G(rs, 0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671)  # = -alpha_c
    fz = _f_zeta(z)
    return e0 - mac * fz / FPP0 * (1 - z ** 4) + (e1 - e0) * fz * z ** 4


def pbec(a, b, gaa, gab, gbb):
    beta = 0.06672455060314922
    gamma = (1 - math.log(2.0)) / (PI * PI)
    n = a + b
    g = gaa + 2 * gab + gbb
    z = (a - b) / n
    rs = (3.0 / (4.0 * PI * n)) ** (1.0 / 3.0)
    eps = _pw92_eps(rs, z)
    phi = 0.5 * ((1 + z) ** (2.0 / 3.0) + (1 - z) ** (2.0 / 3.0))
    kF = (3 * PI * PI * n) ** (1.0 / 3.0)
    ks2 = 4 * kF / PI
    t2 = g / (4 * phi * phi * ks2 * n * n)
    Aa = beta / gamma / (torch.exp(-eps / (gamma * phi ** 3)) - 1)
    H = gamma * phi ** 3 * torch.log(1 + beta / gamma * t2 * (1 + Aa * t2) / (1 + Aa * t2 + Aa * Aa * t2 * t2))
    return n * (eps + H)


def _pz81_eps(rs, z):
    def branch(rs, g, b1, b2, A, B, C, D):
        hi = g / (1 + b1 * torch.sqrt(rs) + b2 * rs)
        lo = A * torch.log(rs) + B + C * rs * torch.log(rs) + D * rs
        return torch.where(rs >= 1.0, hi, lo)
    eU = branch(rs, -0.1423, 1.0529, 0.3334, 0.0311, -0.048, 0.0020, -0.0116)
    eP = branch(rs, -0.0843, 1.3981, 0.2611, 0.01555, -0.0269, 0.0007, -0.0048)
    return eU + _f_zeta(z) * (eP - eU)


def p86c(a, b, gaa, gab, gbb):
    n = a + b
    g = gaa + 2 * gab + gbb
    z = (a - b) / n
    rs = (3.0 / (4.0 * PI * n)) ** (1.0 / 3.0)
    Cn = 0.001667 + (0.002568 + 0.023266 * rs + 7.389e-6 * rs * rs) / (1 + 8.723 * rs + 0.472 * rs * rs + 0.07389 * rs ** 3)
    Phi = (9.0 * PI) ** (1.0 / 6.0) * 0.11 * 0.004235 / Cn * torch.sqrt(g) / n ** (7.0 / 6.0)
    d = 2.0 ** (1.0 / 3.0) * torch.sqrt(((1 + z) / 2) ** (5.0 / 3.0) + ((1 - z) / 2) ** (5.0 / 3.0))
    return n * _pz81_eps(rs, z) + torch.exp(-Phi) * Cn * g / (d * n ** (4.0 / 3.0))


def _lc94_F(s):
    a1, a2, a3, a4, aa, bb = 0.093907, 76.320, 0.26608, 0.0809615, 100.0, 0.57767e-4
    L = a1 * s * torch.asinh(a2 * s)
    return (1 + L + (a3 - a4 * torch.exp(-aa * s * s)) * s * s) / (1 + L + bb * s ** 4)


def pw91k(a, b, gaa, gab, gbb):
    def spin(r, g):
        s = torch.sqrt(g) / (2 * (6 * PI * PI) ** (1.0 / 3.0) * r ** (4.0 / 3.0))
        return 2.0 ** (2.0 / 3.0) * CF * r ** (5.0 / 3.0) * _lc94_F(s)
    return spin(a, gaa) + spin(b, gbb)


def llp91k(a, b, gaa, gab, gbb):
    def spin(r, g):
        x = torch.sqrt(g) / r ** (4.0 / 3.0)
        return 2.0 ** (2.0 / 3.0) * CF * r ** (5.0 / 3.0) * (1 + 0.0044188 * x * x / (1 + 0.0253 * x * torch.asinh(x)))
    return spin(a, gaa) + spin(b, gbb)


# BASIC_FUNCTIONALS enum value -> expression
BY_ID = {2: slaterx, 45: vwn5c, 66: tfk, 80: beckex, 81: beckecorrx, 135: pbex, 184: lypc, 193: p86c, 197: pbec,
         283: pw91k, 286: llp91k}


def closed_shell(fid, rho, sigma):
    """F, dF/drho, dF/dsigma at rho_a = rho_b = rho/2 by autograd (numpy in, numpy out)."""
    r = torch.tensor(rho, dtype=torch.float64, requires_grad=True)
    s = torch.tensor(sigma, dtype=torch.float64, requires_grad=True)
    F = BY_ID[fid](r / 2, r / 2, s / 4, s / 4, s / 4)
    vr, vs = torch.autograd.grad(F.sum(), [r, s], allow_unused=True)
    if vs is None:
        vs = torch.zeros_like(r)
    return F.detach().numpy(), vr.numpy(), vs.numpy()


def spin_resolved(fid, ra, rb, gaa, gab, gbb):
    """F and dF/d(rho_a, rho_b, s_aa, s_ab, s_bb) by autograd (numpy arrays in, F [n] and d [5, n] out)."""
    xs = [torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (ra, rb, gaa, gab, gbb)]
    F = BY_ID[fid](*xs)
    gr = torch.autograd.grad(F.sum(), xs, allow_unused=True)
    d = [torch.zeros_like(xs[0]) if g is None else g for g in gr]
    return F.detach().numpy(), torch.stack(d).numpy()
