"""axisem3d_b200/spectral.py (GLL / GLJ nodes, weights and derivative matrices derived numerically) against the
reference's hard-coded nPol = 4 table (SpectralConstants.cpp:43-49; fixture tests/golden/spectral_npol4.json written by
oracle/make_golden_spectral.py), and nextLuckyNumber against known values of PreloopFFTW.cpp:59-109."""
import json
import os

import numpy as np

from axisem3d_b200 import spectral as SP

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spectral_npol4.json")))
TOL = 2e-12   # the table has 12 decimals


def test_nodes_and_weights_match_reference_table():
    assert np.abs(SP.P_GLL - np.array(GOLD["pgll_4"])).max() < TOL
    assert np.abs(SP.P_GLJ - np.array(GOLD["pglj_4"])).max() < TOL
    assert np.abs(SP.W_GLL - np.array(GOLD["wgll_4"])).max() < TOL
    assert np.abs(SP.W_GLJ - np.array(GOLD["wglj_4"])).max() < TOL


def test_gradient_matrices_match_reference_table():
    # the table is consumed row-major into G(i, j) = l_i'(x_j) (SpectralConstants.cpp:104-110)
    assert np.abs(SP.G_GLL.reshape(-1) - np.array(GOLD["Ggll_4"])).max() < TOL
    assert np.abs(SP.G_GLJ.reshape(-1) - np.array(GOLD["Gglj_4"])).max() < TOL
    # sum_i l_i'(x) = 0: the columns of G sum to zero
    assert np.abs(SP.G_GLL.sum(axis=0)).max() < 1e-12 and np.abs(SP.G_GLJ.sum(axis=0)).max() < 1e-12


def test_next_lucky_number():
    # Nu -> Nr under the even rule and the forced-odd rule
    for nu, even, odd in ((2, 5, 5), (20, 42, 45), (100, 208, 225), (200, 416, 405), (500, 1008, 1029), (1000, 2016, 2025)):
        assert SP.next_lucky_number(2 * nu + 1) == even
        assert SP.next_lucky_number(2 * nu + 1, force_odd=True) == odd
    assert SP.next_lucky_number(17 * 2) == 36          # prime factor > 13 is skipped
    assert SP.next_lucky_number(11 * 13 * 2) == 288    # at most one factor of 11 or 13 in total: 286 = 2 * 11 * 13 is not lucky
