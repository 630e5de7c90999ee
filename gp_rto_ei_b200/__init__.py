"""gp_rto_ei_b200 - B200-native (sm_100a) GP fitting / inference / acquisition hot path of GP-RTO-EI.

The directory is named with underscores so that it is importable; it is the package the build brief
calls `gp-rto-ei_b200/`.  Layout: csrc/ (CUDA kernels + C ABI, built into libgprto.so), _lib.py (ctypes
binding), api.py (batched device API), parallel.py (one-process-per-GPU sharding + NCCL gather),
sub_uts/ (drop-in mirror of the reference's sub_uts/utilities_2.py module), compat/ (stand-ins for
third-party modules missing from the image), workloads.py (synthetic benchmark inputs).
Importing the package never touches CUDA; using it without libgprto.so or without a B200 raises.
"""
__version__ = '0.1'
