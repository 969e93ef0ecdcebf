"""Stand-in for the reference's pybind11 extension module `cauchy_mult` (extensions/cauchy/cauchy.cpp:86-95) on libdwb.

Put this directory on sys.path and the reference's own `extensions/cauchy/cauchy.py` (which does
`from cauchy_mult import cauchy_mult_fwd, cauchy_mult_bwd, cauchy_mult_sym_fwd, cauchy_mult_sym_bwd`) and therefore
`models/s4.py:35-42` run unmodified on the B200 kernels:

    import sys, diffwave_sashimi_b200.shims as shims
    sys.path.insert(0, shims.PATH)            # before importing models.s4 / extensions.cauchy.cauchy

Same signatures, CUDA complex64 tensors in, freshly allocated outputs, current torch stream.
"""
from diffwave_sashimi_b200.ops import cauchy_mult_bwd, cauchy_mult_fwd, cauchy_mult_sym_bwd, cauchy_mult_sym_fwd  # noqa: F401
