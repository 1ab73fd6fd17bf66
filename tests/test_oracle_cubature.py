"""CPU: pin the oracle's sigma-point rules with known-answer identities, and the
product's host-side tables against the oracle (utp_ws.m, ut{3,5,7,9}_ws.m,
sym_set.m, gauher.m, mvhermgauss.m)."""
import itertools
import math

import numpy as np
import pytest

from oracle import cubature as oc


def _gauss_moment(powers):
    """E[prod x_i^p_i] for a standard normal vector."""
    out = 1.0
    for p in powers:
        if p % 2:
            return 0.0
        out *= float(np.prod(np.arange(p - 1, 0, -2))) if p else 1.0
    return out


@pytest.mark.parametrize("p", [3, 5, 7, 9])
@pytest.mark.parametrize("n", [2, 3, 4])
def test_rule_integrates_monomials_exactly(p, n):
    """A p-th order rule integrates every monomial of total degree <= p exactly.
    The 9th-order rule with n >= 3 carries the reference's sign quirk in the
    centre weight (ut9_ws.m:78-79): only the constant monomial is affected."""
    W, SX = oc.utp_ws(p, n)
    quirk = (p == 9 and n >= 3)
    for powers in itertools.product(range(p + 1), repeat=n):
        if sum(powers) > p:
            continue
        val = float(np.sum(W * np.prod(SX ** np.array(powers)[:, None], axis=0)))
        ref = _gauss_moment(powers)
        if quirk and sum(powers) == 0:
            continue
        assert abs(val - ref) < 1e-10 * max(1.0, abs(ref)), (powers, val, ref)


def test_ut9_weight_sum_quirk():
    """sum(W) = 1 for N=2, 1.2637037037... for N=3 (doubled minus sign)."""
    assert abs(oc.ut9_ws(2)[0].sum() - 1.0) < 1e-12
    assert abs(oc.ut9_ws(3)[0].sum() - 1.263703703703704) < 1e-10
    # the excess is exactly 16*C(n,3)*(A111+A222)
    W, SX = oc.ut9_ws(3)
    nz = np.count_nonzero(SX, axis=0)
    A3sum = W[nz == 3].sum() / 8.0        # 8 sign patterns per orbit; both generators u and v
    assert abs((W.sum() - 1.0) - 16 * A3sum) < 1e-12


@pytest.mark.parametrize("p,n,S", [(3, 2, 5), (5, 3, 19), (7, 3, 45), (9, 2, 25), (9, 3, 77), (9, 4, 193)])
def test_point_counts(p, n, S):
    W, SX = oc.utp_ws(p, n)
    assert W.shape == (S,) and SX.shape == (n, S)


def test_ut3_centre_weight_is_zero():
    W, SX = oc.ut3_ws(3)
    assert W[0] == 0.0 and np.allclose(np.abs(SX).max(), math.sqrt(3))


def test_roots_order_larger_first():
    """u is the larger generator (order of MATLAB/NumPy `roots` on these quartics)."""
    _, SX7 = oc.ut7_ws(2)
    assert abs(SX7[0, 1] - math.sqrt(3 + math.sqrt(6))) < 1e-12
    _, SX9 = oc.ut9_ws(2)
    assert abs(SX9[0, 1] - math.sqrt(5 + math.sqrt(10))) < 1e-12


def test_gauher_table_and_golub_welsch():
    x20, w20 = oc.gauher(20)
    assert x20[0] == -7.619048541679757 and w20[9] == 0.260793063449555     # verbatim table
    assert abs(w20.sum() - 1) < 1e-9 and abs((w20 * x20 ** 2).sum() - 1) < 1e-8
    for n in (3, 5, 9, 12):
        x, w = oc.gauher(n)
        for k in range(0, 2 * n, 2):
            assert abs((w * x ** k).sum() - _gauss_moment([k])) < 1e-9 * max(1, _gauss_moment([k]))


def test_mvhermgauss_ndgrid_order():
    xn, wn = oc.mvhermgauss(np.array([1.0, -2.0]), np.array([4.0, 0.25]), 3)
    t, w = oc.gauher(3)
    assert xn.shape == (9, 2)
    assert np.allclose(xn[1], [1.0 + 2.0 * t[1], -2.0 + 0.5 * t[0]])     # first dimension varies fastest
    assert np.allclose(wn[5], w[2] * w[1])


@pytest.mark.parametrize("p", [3, 5, 7, 9])
@pytest.mark.parametrize("n", [2, 3, 4])
def test_product_tables_match_oracle(nsagp, p, n):
    W, SX = nsagp.utp_ws(p, n)
    Wo, SXo = oc.utp_ws(p, n)
    assert np.allclose(W, Wo, rtol=0, atol=1e-13) and np.allclose(SX, SXo, rtol=0, atol=1e-14)


def test_product_gauss_hermite_matches_oracle(nsagp):
    for N, p in ((2, 3), (3, 5), (2, 20)):
        wn, xn = nsagp.mvhermgauss_unit(N, p)
        xo, wo = oc.mvhermgauss_unit(N, p)
        assert np.array_equal(wn, wo) and np.array_equal(xn, xo.T)
