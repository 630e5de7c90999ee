"""No-op stand-in for the reference's plotting module (plots_RTO.py / run_simple/plots_RTO_simple.py need matplotlib
with LaTeX): every plotting entry point the driver scripts import accepts its arguments and does nothing."""


def _noop(*a, **k):
    return None


Plot = Plot_simple = compute_obj = compute_obj_simple = plot_obj = plot_obj_noise = _noop
