"""The reference's own driver scripts, UNCHANGED, on top of this package - as far as a machine without a GPU can take
them: imports (casadi / matplotlib / sobol_seq / plots_RTO stand-ins), the star-imported names the scripts rely on,
plant / model set-up, ITR_GP_RTO construction and initial data handling all run; the first GP fit then fails loudly
because there is no CUDA device (no CPU fallback).  Only runs where the reference tree exists (the build container)."""
import contextlib
import io
import os

import pytest
import torch

from gp_rto_ei_b200 import run_main
from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available() or torch.cuda.is_available(),
                                reason='needs /root/reference and no GPU (on a GPU box the full scripts would run for hours)')


@pytest.mark.parametrize('script', ['run_simple/main_simple.py', 'run_WO/main_WO_prior_EI.py'])
def test_reference_script_reaches_the_gpu_boundary(script, tmp_path):
    path = os.path.join(ref_harness.REFERENCE_ROOT, script)
    out = io.StringIO()
    with pytest.raises(RuntimeError, match='CUDA device'):
        with contextlib.redirect_stdout(out):
            run_main.run(path, episodes=1, out=str(tmp_path))
    log = out.getvalue()
    # the script got through ITR_GP_RTO.__init__ (data handling prints) and into RTO_routine before the GP fit
    assert 'preliminar data supplied' in log and 'constraints are set as positive' in log
    import sub_uts.utilities_2 as u2
    import sub_uts.systems as systems
    assert u2.__file__.startswith(os.path.dirname(run_main.HERE)) and hasattr(systems, 'WO_system')
