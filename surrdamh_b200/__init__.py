"""surrdamh_b200 -- the data-parallel hot path of surrDAMH (lockstep MH / delayed-acceptance MH
chains with polynomial / RBF surrogates) on NVIDIA B200 (sm_100a).

Layout: csrc/ (CUDA kernels + the C ABI of include/surrdamh_b200.h), _lib.py (ctypes binding),
device.py / chains.py (device state), surrogate_poly.py / surrogate_rbf.py (the reference's
surrogate interface), configuration.py / sampler.py / collector.py (the reference's JSON
configuration and process_SAMPLER / process_COLLECTOR logic for lockstep chains).
There is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
