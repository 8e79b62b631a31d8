"""photic_b200 -- B200-native (sm_100a, FP64) per-pixel semi-analytical inversion.

A from-scratch CUDA implementation of ONE hot path of stblake/photic: the HOPE/Lee forward model
(model/samodel.c) minimised per pixel by the AS 047 Nelder-Mead simplex (model/asa047.c), behind
the reference's ``samodel()`` call surface. The product is the C-ABI library
``photic_b200/csrc/libphotic_b200.so`` (see include/photic_b200.h); this package is the thin
Python host mirror used by tests and bench.py.
"""
__version__ = "0.1.0"
