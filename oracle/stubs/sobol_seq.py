"""sobol_seq stand-in for importing the reference: re-exports the restated generator
(gp_rto_ei_b200/compat/sobol_seq.py, pinned by Burkardt's published rows in tests/test_sobol.py)."""
from gp_rto_ei_b200.compat.sobol_seq import i4_sobol_generate, i4_sobol  # noqa: F401
