"""fullysparsefusion_b200 — the B200-native hot path of Fully Sparse Fusion.

Layout:
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/fsf_b200.h)
  _capi.py   ctypes binding of libfsf_b200.so (fails loudly when the library is missing)
  ops.py     torch-tensor front end (device memory + streams only)
  shims/     drop-in modules under the names the reference plugin imports
  synth.py   seeded synthetic inputs shared by tests and bench
Nothing in this package imports oracle/.
"""
__version__ = "0.1.0"
