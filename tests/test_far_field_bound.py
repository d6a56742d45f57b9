"""CPU check of the error bound the far-field treatment of distant far wings relies on
(k_far_nodes + the polynomial term of k_voigt_tile, csrc/sr_voigt.cu):
the region-1 rational of humliv_bb (lineshape.f:456-477), sampled at 12 Chebyshev nodes of a
512-point tile, converted to monomial coefficients and evaluated by Horner's rule, reproduces every
point of the tile to <= 4e-11 of the line's own value whenever the kernel's predicate holds - the
whole tile on one side of the line and its nearest point at least 2 tile lengths + ry + 1 Doppler
widths away (far_full_half).  Pure NumPy: the numbers do not depend on the GPU."""
import numpy as np
from numpy.polynomial import chebyshev as C

TP, NN = 512, 12


def reg1(x, ry):
    """K = (a + x^2 b) / (c + x^2 (d + 4 x^2)), lineshape.f:456-477, in the kernel's u-form."""
    u = x * x + ry * ry - 0.5
    return 0.5641896 * ry * (u + 1.0) / (u * u + 2.0 * ry * ry)


def node_to_monomial():
    t = np.cos((2 * np.arange(NN) + 1) * np.pi / (2 * NN))
    Tm = np.array([[(1.0 if j == 0 else 2.0) / NN * np.cos(j * (2 * n + 1) * np.pi / (2 * NN))
                    for n in range(NN)] for j in range(NN)])
    c2p = np.zeros((NN, NN))
    for j in range(NN):
        e = np.zeros(NN)
        e[j] = 1.0
        p = C.cheb2poly(e)
        c2p[:len(p), j] = p
    return t, c2p @ Tm


def test_twelve_nodes_reproduce_a_distant_far_wing_to_4e_11():
    t, M = node_to_monomial()
    rng = np.random.default_rng(1)
    P = np.arange(TP)
    s = 2.0 * P / (TP - 1) - 1.0
    worst = 0.0
    for ry in (1e-4, 0.05, 1.0, 10.0, 80.0):
        for xs in (0.02, 0.12, 0.5):
            L = (TP - 1) * xs
            d_min = 2.0 * L + ry + 1.0                       # the predicate's boundary: worst case
            for d in [d_min] + list(d_min * (1.0 + rng.uniform(0, 3, 40) ** 3)):
                for side in (1, -1):
                    x = d + P * xs if side > 0 else -(d + (TP - 1 - P) * xs)
                    Pn = 0.5 * (TP - 1) * (1.0 + t)
                    xn = d + Pn * xs if side > 0 else -(d + (TP - 1 - Pn) * xs)
                    coef = M @ reg1(xn, ry)
                    fi = np.zeros(TP)
                    for c in coef[::-1]:
                        fi = fi * s + c
                    f = reg1(x, ry)
                    worst = max(worst, float(np.max(np.abs(fi - f) / np.abs(f))))
    assert worst <= 4e-11, worst
    # and the bound is what buys the 1e-6 parity gate its margin: five orders of magnitude
    assert worst < 1e-6 * 1e-4


def test_one_tile_length_would_not_be_enough():
    """Why the predicate asks for TWO tile lengths: at one tile length the same scheme is three
    orders of magnitude worse."""
    t, M = node_to_monomial()
    P = np.arange(TP)
    s = 2.0 * P / (TP - 1) - 1.0
    ry, xs = 0.05, 0.5
    L = (TP - 1) * xs
    err = []
    for dfac in (1.0, 2.0):
        d = dfac * L + ry + 1.0
        coef = M @ reg1(d + 0.5 * (TP - 1) * (1.0 + t) * xs, ry)
        fi = np.zeros(TP)
        for c in coef[::-1]:
            fi = fi * s + c
        f = reg1(d + P * xs, ry)
        err.append(float(np.max(np.abs(fi - f) / f)))
    assert err[1] <= 4e-11 < err[0] and err[0] > 100 * err[1]
