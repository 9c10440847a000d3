"""ORACLE — test infrastructure only; runs ONLY in the build container (needs /root/reference).

Loads the reference's own DA head classes IN PLACE from /root/reference under a minimal mmcv stub
(mmcv / mmdet are not installable here: SURVEY.md §8c, Appendix D).  Nothing is copied; the
classes are used by oracle/make_golden.py to produce tests/golden/*.pt.
"""
import importlib.util
import os
import sys
import types

import torch.nn as nn

REF = os.environ.get("DA_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "mmdet"))


def _install_stubs():
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_da_stub", False):
        return

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    mmcv = types.ModuleType("mmcv")
    mmcv._da_stub = True
    cnn = types.ModuleType("mmcv.cnn")
    cnn.build_conv_layer = lambda cfg, *a, **k: nn.Conv2d(*a, **k)

    def build_norm_layer(cfg, n, postfix=""):
        return f"bn{postfix}", nn.BatchNorm2d(n)

    cnn.build_norm_layer = build_norm_layer
    cnn.build_plugin_layer = None
    runner = types.ModuleType("mmcv.runner")

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    runner.BaseModule = BaseModule
    runner.Sequential = nn.Sequential
    ops = types.ModuleType("mmcv.ops")
    ops.sigmoid_focal_loss = None
    mmcv.cnn, mmcv.runner, mmcv.ops = cnn, runner, ops
    mmcv.jit = lambda *a, **k: (lambda f: f)
    sys.modules.update({"mmcv": mmcv, "mmcv.cnn": cnn, "mmcv.runner": runner, "mmcv.ops": ops})

    for name in ["mmdet_ref", "mmdet_ref.models", "mmdet_ref.models.backbones", "mmdet_ref.models.roi_heads",
                 "mmdet_ref.models.losses", "mmdet_ref.models.utils"]:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    builder = types.ModuleType("mmdet_ref.models.builder")
    for k in ["BACKBONES", "HEADS", "LOSSES", "DETECTORS", "NECKS", "ROI_EXTRACTORS", "SHARED_HEADS"]:
        setattr(builder, k, _Reg())
    sys.modules["mmdet_ref.models.builder"] = builder


def _load(modname, relpath):
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference classes/functions used for golden vectors."""
    if not available():
        raise RuntimeError(f"{REF} not present: golden vectors can only be regenerated in the build container")
    _install_stubs()
    res_layer = _load("mmdet_ref.models.utils.res_layer", "mmdet/models/utils/res_layer.py")
    sys.modules["mmdet_ref.models.utils"].ResLayer = res_layer.ResLayer
    ns = types.SimpleNamespace()
    ns.daf_org = _load("mmdet_ref.models.backbones.resnet_da_daf_org", "mmdet/models/backbones/resnet_da_daf_org.py")
    ns.maf = _load("mmdet_ref.models.backbones.resnet_da", "mmdet/models/backbones/resnet_da.py")
    ns.cbam = _load("mmdet_ref.models.backbones.resnet_da_cbam", "mmdet/models/backbones/resnet_da_cbam.py")
    ns.deep = _load("mmdet_ref.models.backbones.resnet_da_deep", "mmdet/models/backbones/resnet_da_deep.py")
    ns.instance = _load("mmdet_ref.models.roi_heads.instance_da", "mmdet/models/roi_heads/instance_da.py")
    ns.local_da = _load("mmdet_ref.models.roi_heads.local_da", "mmdet/models/roi_heads/local_da.py")
    ns.loss_utils = _load("mmdet_ref.models.losses.utils", "mmdet/models/losses/utils.py")
    ns.focal = _load("mmdet_ref.models.losses.focal_loss", "mmdet/models/losses/focal_loss.py")
    ns.ce = _load("mmdet_ref.models.losses.cross_entropy_loss", "mmdet/models/losses/cross_entropy_loss.py")
    ns.smooth_l1 = _load("mmdet_ref.models.losses.smooth_l1_loss", "mmdet/models/losses/smooth_l1_loss.py")
    ns.accuracy = _load("mmdet_ref.models.losses.accuracy", "mmdet/models/losses/accuracy.py")
    return ns


# ----------------------------------------------------------------------------------------------
# detector methods of the instance-level "group" loss (L5).  The detector modules cannot be imported
# (they pull in the whole of mmdet) and hard-code device='cuda'; the METHODS are plain torch.  They are
# compiled in place from the reference source (ast -> exec, nothing copied into this repo) with a `torch`
# proxy that drops the device= keyword, so the reference's own code runs on the CPU of the build container.
# ----------------------------------------------------------------------------------------------
class _TorchCpu:
    """Forwards to torch; factory functions ignore device= (the reference passes device='cuda')."""

    def __getattr__(self, name):
        import torch
        obj = getattr(torch, name)
        if name in ("zeros", "ones", "tensor", "randn", "rand", "empty", "full"):
            def wrapped(*a, **k):
                k.pop("device", None)
                return obj(*a, **k)
            return wrapped
        return obj


def load_group_loss(flavour):
    """flavour: 'daf' (DAFaster_rcnn.py:198-327), 'maf' (MAFaster_rcnn.py:204-299), 'deep' (DAFaster_rcnn_Deep.py:232-329).
    Returns a dict of the reference's functions {group_local_da_loss, group?, complete?} taking `self` first."""
    import ast
    import torch.nn.functional as F
    if not available():
        raise RuntimeError(f"{REF} not present")
    _install_stubs()
    cl = _load("mmdet_ref.models.utils.cluster", "mmdet/models/utils/cluster.py")
    cl.torch = _TorchCpu()                       # cluster.py:98 draws its centroids with device='cuda'
    path = {"daf": "DAFaster_rcnn.py", "maf": "MAFaster_rcnn.py", "deep": "DAFaster_rcnn_Deep.py"}[flavour]
    src = open(os.path.join(REF, "mmdet/models/detectors", path)).read()
    tree = ast.parse(src)
    wanted = {"group_local_da_loss", "group", "complete"}
    fns = [f for c in tree.body if isinstance(c, ast.ClassDef) for f in c.body
           if isinstance(f, ast.FunctionDef) and f.name in wanted]
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {"torch": _TorchCpu(), "F": F, "cluster": cl.cluster}
    exec(compile(mod, os.path.join(REF, "mmdet/models/detectors", path), "exec"), ns)
    return {f.name: ns[f.name] for f in fns}


def load_batch_sampler():
    """The reference's BatchSchedulerSampler class (mmdet/datasets/samplers/batch_sampler.py), loaded in place."""
    if not available():
        raise RuntimeError(f"{REF} not present")
    _install_stubs()
    sys.modules["mmcv.runner"].get_dist_info = lambda: (0, 1)
    return _load("mmdet_ref.datasets.samplers.batch_sampler", "mmdet/datasets/samplers/batch_sampler.py").BatchSchedulerSampler
