"""b200-typicality: B200-native engine for the typicality / DIFT hot path of ysig/diff-mining.

Host side stays Python (mirroring the reference's `SD` / `D` / `SDFeaturizer` surface); all device work goes
through the C ABI of `libdm_b200.so` (include/dm_abi.h) into hand-written sm_100a CUDA.  There is no CPU or
PyTorch fallback on the product path: if the shared library is missing, importing `_abi` raises.
"""
__version__ = "0.1.0"
