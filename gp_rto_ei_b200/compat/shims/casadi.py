"""Minimal stand-in used ONLY when the real casadi is not importable, so that the reference's driver scripts, which
start with ``from casadi import *``, can run on this package.  The scripts themselves use nothing from casadi except
the names it happens to re-export (``os``: run_simple/main_simple.py:21); the Williams-Otto plant/model of
gp_rto_ei_b200/sub_uts/systems.py solve their balances with numpy instead of casadi's rootfinder."""
import os  # noqa: F401

__all__ = ['os']
