"""Mirror of the reference's ``sub_uts`` package: put gp_rto_ei_b200/ on sys.path (gp_rto_ei_b200/run_main.py does)
and ``from sub_uts.utilities_2 import *`` / ``from sub_uts.systems import *`` resolve here."""
