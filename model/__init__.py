"""Drop-in alias: `import model...` resolves to `cofii2p_b200.model...` so that the reference's scripts
(`from model.network import CoFiI2P`, `from model.loss import *`, `from model.kpconv.preprocess_data import ...`;
reference train.py:13-17, evaluation/eval_all.py:9-14) run against the B200 implementation unchanged when this
repository precedes the reference on sys.path.  Nothing is implemented here."""
import importlib
import sys

_SUBMODULES = ("network", "imagenet", "loss", "kpconv", "kpconv.kp_backbone", "kpconv.modules", "kpconv.kpconv",
               "kpconv.functional", "kpconv.kernel_points", "kpconv.preprocess_data", "transformer", "transformer.transformer",
               "transformer.position_encoding", "transformer.linear_attention")

_pkg = importlib.import_module("cofii2p_b200.model")
for _name in _SUBMODULES:
    _mod = importlib.import_module(f"cofii2p_b200.model.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    if "." not in _name:
        globals()[_name] = _mod
