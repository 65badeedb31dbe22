"""fgnn_b200 — Python-side plumbing around the sm_100a kernel library.

`kernels`  ctypes binding of include/fgnn_kernels.h on torch device tensors
`synth`    deterministic synthetic power-law graphs of the reference's dataset shapes
"""
