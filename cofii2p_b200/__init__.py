"""cofii2p_b200 -- B200-native (sm_100a) implementation of CoFiI2P's coarse-to-fine correspondence hot path.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI declared in `include/cofi_b200.h`),
`lib.py` (ctypes binding, fails loudly when the library is missing), `ops.py` (tensor-level wrappers),
`model/` (host-side mirror of the reference's `model` package: same class names, same state_dict) and
`frames.py` (synthetic KITTI-shaped frames). There is no CPU fallback: the CPU restatement of the algorithm
is test infrastructure under `oracle/`.
"""
__version__ = "0.1.0"
