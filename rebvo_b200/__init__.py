"""rebvo_b200 -- B200 (sm_100a) implementation of REBVO's per-frame edge pipeline.

The product is librebvo_b200.so (hand-written CUDA behind the C ABI of include/rebvo_b200.h); this package
is only the Python-side binding used by the tests and bench.py.  There is no CPU fallback: importing
`rebvo_b200.capi` fails loudly when the library has not been built.
"""
