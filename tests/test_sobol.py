"""Known-answer tests for the sobol_seq stand-in (reference use: sub_uts/utilities_2.py:139, :441)."""
import numpy as np

from gp_rto_ei_b200.compat import sobol_seq

# First rows of Burkardt's sobol_test output for DIM_NUM = 5 with skip = 1
KAT5 = np.array([[.5, .5, .5, .5, .5], [.75, .25, .75, .25, .75], [.25, .75, .25, .75, .25],
                 [.375, .375, .625, .125, .875], [.875, .875, .125, .625, .375], [.625, .125, .375, .375, .125]])


def test_kat_dim5():
    assert np.array_equal(sobol_seq.i4_sobol_generate(5, 6), KAT5)


def test_prefix_and_dims_consistent():
    a = sobol_seq.i4_sobol_generate(4, 64)
    b = sobol_seq.i4_sobol_generate(5, 64)
    assert np.array_equal(a, b[:, :4])            # dimension j does not depend on dim_num
    assert np.array_equal(a[:15], sobol_seq.i4_sobol_generate(4, 15))


def test_is_a_digital_net():
    # any 2^k consecutive points starting at index 0 (skip=0) hit every dyadic interval of length 2^-k once per dimension
    for dim in (2, 3, 5, 8):
        p = sobol_seq.i4_sobol_generate(dim, 64, skip=0)
        for j in range(dim):
            assert sorted(np.floor(p[:, j] * 64).astype(int)) == list(range(64))


def test_first_dimension_is_van_der_corput():
    p = sobol_seq.i4_sobol_generate(1, 8, skip=0)[:, 0]
    assert np.array_equal(p, [0, .5, .75, .25, .375, .875, .625, .125])   # Gray-code order


def test_i4_sobol_single_point():
    q, nxt = sobol_seq.i4_sobol(5, 4)
    assert nxt == 5 and np.array_equal(q, KAT5[3])
