"""Host-side mirror of the reference's `model` package for the hot path: same module paths, class names,
constructor arguments and `state_dict` keys (430 entries), so that the reference's `train.py` /
`evaluation/eval_all.py` call sites work unchanged.  The forward passes run exclusively on the sm_100a
kernels of libcofi_b200.so through `cofii2p_b200.ops`."""
