"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (WHU-USI3DV/CoFiI2P) as an oracle.

This module exists only in the build container: `/root/reference` is not present on the GPU box,
so nothing under `tests/ -m gpu`, `bench.py` or `__graft_entry__.smoke()` may call `load_reference()`.
It is used by `oracle/make_golden.py` (generates `tests/golden/*.npz`) and by the `-m "not gpu"` tests
that pin `oracle/restate.py` against the reference when the reference tree is available.

What the shim does:
  1. registers stub modules for `open3d` / `matplotlib` (imported at reference `model/network.py:12`,
     `model/kpconv/kernel_points.py:21-23`, used there only for PLY I/O of the kernel-point disposition);
  2. when CUDA is absent, makes `torch.Tensor.cuda` a no-op (reference `model/network.py:105,156,180,181`
     hard-codes `.cuda()` inside forward);
  3. imports the reference `model` package under the alias `cofi_ref_model` from a scratch copy in /tmp
     (the reference writes `dispositions/k_015_center_3D.ply` beside its sources on first use,
     `model/kpconv/kernel_points.py:392-394,421`; /root/reference must stay untouched).
No reference source is copied into this repository.
"""
import importlib
import importlib.util
import os
import shutil
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("COFI_REFERENCE_ROOT", "/root/reference")
SCRATCH = os.environ.get("COFI_REFERENCE_SCRATCH", "/tmp/cofi_ref_scratch")
ALIAS = "cofi_ref_model"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "network.py"))


def _install_stubs():
    if "open3d" not in sys.modules:
        o3d = types.ModuleType("open3d")
        geometry = types.ModuleType("open3d.geometry")
        utility = types.ModuleType("open3d.utility")
        io = types.ModuleType("open3d.io")

        class PointCloud:  # bare container, `.points` attribute only
            def __init__(self):
                self.points = None

        class KDTreeFlann:  # only has to exist (dead `search_point_index`)
            def __init__(self, *a, **k):
                raise NotImplementedError("open3d stub")

        def write_point_cloud(path, pcd):
            np.save(path + ".npy", np.asarray(pcd.points))
            with open(path, "wb") as f:  # `load_kernels` tests exists(kernel_file)
                f.write(b"stub")
            return True

        def read_point_cloud(path):
            pcd = PointCloud()
            pcd.points = np.load(path + ".npy")
            return pcd

        geometry.PointCloud = PointCloud
        geometry.KDTreeFlann = KDTreeFlann
        utility.Vector3dVector = np.asarray
        io.write_point_cloud = write_point_cloud
        io.read_point_cloud = read_point_cloud
        o3d.geometry, o3d.utility, o3d.io = geometry, utility, io
        sys.modules["open3d"] = o3d
        sys.modules["open3d.geometry"] = geometry
        sys.modules["open3d.utility"] = utility
        sys.modules["open3d.io"] = io
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt


def _patch_cuda_noop():
    import torch

    if not torch.cuda.is_available() and not getattr(torch.Tensor.cuda, "_cofi_noop", False):
        def _noop(self, *a, **k):
            return self
        _noop._cofi_noop = True
        torch.Tensor.cuda = _noop


def load_reference():
    """Returns the reference `model` package (alias `cofi_ref_model`) and its Options_KITTI class."""
    if not reference_available():
        raise RuntimeError("reference tree not available at %s (oracle shim runs in the build container only)"
                           % REFERENCE_ROOT)
    _install_stubs()
    _patch_cuda_noop()
    if ALIAS in sys.modules:
        pkg = sys.modules[ALIAS]
    else:
        dst = os.path.join(SCRATCH, "model")
        if not os.path.isdir(dst):
            os.makedirs(SCRATCH, exist_ok=True)
            shutil.copytree(os.path.join(REFERENCE_ROOT, "model"), dst)
            os.makedirs(os.path.join(SCRATCH, "data"), exist_ok=True)
            shutil.copy(os.path.join(REFERENCE_ROOT, "data", "options.py"), os.path.join(SCRATCH, "data", "options.py"))
        spec = importlib.util.spec_from_file_location(
            ALIAS, os.path.join(dst, "__init__.py"), submodule_search_locations=[dst])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules[ALIAS] = pkg
        spec.loader.exec_module(pkg)
        importlib.import_module(ALIAS + ".network")
        importlib.import_module(ALIAS + ".loss")
    ospec = importlib.util.spec_from_file_location("cofi_ref_options", os.path.join(SCRATCH, "data", "options.py"))
    omod = importlib.util.module_from_spec(ospec)
    ospec.loader.exec_module(omod)
    return pkg, omod.Options_KITTI


def load_reference_preprocess():
    """The reference's model/kpconv/preprocess_data.py (its `knn`, `square_distance`, `precompute_point_cloud_*`).
    That module imports `open3d.ml.torch.layers` (FixedRadiusSearch, KNNSearch: native, absent here); the stub
    KNNSearch is an exact brute-force search (fp64 direct distances, ties to the lower index) so that the reference's own
    stack-mode driver code -- the sampling and the which-cloud-queries-which wiring -- can run unmodified."""
    import torch

    load_reference()
    if "open3d.ml" not in sys.modules:
        ml = types.ModuleType("open3d.ml")
        mlt = types.ModuleType("open3d.ml.torch")
        layers = types.ModuleType("open3d.ml.torch.layers")

        class _Result:
            def __init__(self, idx):
                self.neighbors_index = idx

        class KNNSearch:
            def __init__(self, return_distances=False, **kw):
                pass

            def __call__(self, points, queries, k):
                p, q = points.double(), queries.double()
                rows = []
                for a in range(0, q.shape[0], 1024):
                    d = ((q[a:a + 1024, None, :] - p[None, :, :]) ** 2).sum(-1)
                    rows.append(torch.argsort(d, dim=1, stable=True)[:, :k])
                return _Result(torch.cat(rows, 0).reshape(-1))

        class FixedRadiusSearch:
            def __init__(self, *a, **kw):
                raise NotImplementedError("open3d stub")

        layers.KNNSearch, layers.FixedRadiusSearch = KNNSearch, FixedRadiusSearch
        ml.torch, mlt.layers = mlt, layers
        sys.modules["open3d"].ml = ml
        sys.modules["open3d.ml"], sys.modules["open3d.ml.torch"], sys.modules["open3d.ml.torch.layers"] = ml, mlt, layers
    return importlib.import_module(ALIAS + ".kpconv.preprocess_data")


def build_reference_model(seed: int = 0):
    """Seeded reference CoFiI2P (eval mode). Construction draws from both torch and numpy RNGs
    (reference `model/kpconv/kernel_points.py:426-453` uses np.random for the per-layer kernel rotation)."""
    import torch

    pkg, Options = load_reference()
    torch.manual_seed(seed)
    np.random.seed(seed)
    opt = Options()
    net = sys.modules[ALIAS + ".network"].CoFiI2P(opt)
    net.eval()
    return net, opt
