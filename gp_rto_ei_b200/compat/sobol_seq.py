"""Host-side stand-in for the PyPI ``sobol_seq`` module (absent from this image).

The reference draws its hyper-parameter multistarts and initial designs with
``sobol_seq.i4_sobol_generate(dim_num, n)`` (reference: sub_uts/utilities_2.py:139, :441).
That package is a Python port of J. Burkardt's ``i4_sobol`` (Bratley & Fox, ACM TOMS
Algorithm 659; 30-bit direction numbers, Gray-code ordering, ``skip=1`` so the first
returned row is the all-0.5 point).  This file restates the published algorithm from
its description; it does not share code with the PyPI package.

Pinned by the known-answer rows of Burkardt's ``sobol_test`` output in
tests/test_sobol.py.  Only the first eight dimensions are tabulated (the path needs
dim = d+2 <= 5 for the hot path and dim = d <= 3 for initial designs).
"""
import numpy as np

_BITS = 30

# (degree, interior coefficient bits a_1..a_{deg-1} of the primitive polynomial, initial m values)
# Bratley-Fox / Burkardt table, dimensions 2..8 (dimension 1 is the van der Corput sequence).
_POLY = {
    2: (1, [], [1]),             # x + 1
    3: (2, [1], [1, 1]),          # x^2 + x + 1
    4: (3, [0, 1], [1, 3, 7]),    # x^3 + x + 1
    5: (3, [1, 0], [1, 1, 5]),    # x^3 + x^2 + 1
    6: (4, [0, 0, 1], [1, 3, 1, 1]),      # x^4 + x + 1
    7: (4, [1, 0, 0], [1, 1, 3, 7]),      # x^4 + x^3 + 1
    8: (5, [0, 0, 1, 0], [1, 3, 3, 9, 9]),  # x^5 + x^2 + 1
}
MAX_DIM = 8


def _direction_numbers(dim_num):
    """v[i, j]: direction integer j of dimension i, already scaled to 30-bit fractions."""
    if not 1 <= dim_num <= MAX_DIM:
        raise ValueError("sobol_seq shim supports 1 <= dim_num <= %d" % MAX_DIM)
    m = np.zeros((dim_num, _BITS), dtype=np.int64)
    m[0, :] = 1
    for dim in range(2, dim_num + 1):
        deg, coef, init = _POLY[dim]
        row = m[dim - 1]
        row[:deg] = init
        for j in range(deg, _BITS):
            new = row[j - deg] ^ (row[j - deg] << deg)
            for k in range(1, deg):
                if coef[k - 1]:
                    new ^= row[j - k] << k
            row[j] = new
    shifts = np.arange(_BITS - 1, -1, -1, dtype=np.int64)
    return m << shifts[None, :]


def _lowest_zero_bit(i):
    c = 0
    while i & 1:
        i >>= 1
        c += 1
    return c


def i4_sobol_generate(dim_num, n, skip=1):
    """Return ``r[n, dim_num]``; row j is the Sobol point with Gray-code index ``j + skip``."""
    v = _direction_numbers(dim_num)
    state = np.zeros(dim_num, dtype=np.int64)
    for i in range(skip):                      # advance from the origin to index `skip`
        state ^= v[:, _lowest_zero_bit(i)]
    out = np.empty((n, dim_num))
    scale = 1.0 / float(1 << _BITS)
    for j in range(n):
        out[j, :] = state * scale
        state = state ^ v[:, _lowest_zero_bit(j + skip)]
    return out


def i4_sobol(dim_num, seed):
    """Single point + next seed, in the calling convention of the PyPI module."""
    return i4_sobol_generate(dim_num, 1, skip=seed)[0], seed + 1
