"""Drop-in import names for the reference. `install()` registers this package's modules
under the names vision3d imports, so `vision3d.detector.{Second, PV_RCNN}` run unchanged:

    import vision3d_b200.compat as compat; compat.install()
    sys.path.insert(0, "<reference checkout>"); from vision3d.detector import Second

  vision3d._C                    -> compat.c_ext        (rotated IoU / NMS)
  spconv, spconv.utils           -> compat.spconv       (SparseConvTensor, SubMConv3d, SparseConv3d, ...)
  pointnet2.pointnet2_utils/...  -> compat.pointnet2    (FPS, gather, ball query, grouping, SA-MSG)
  torchsearchsorted              -> compat.torchsearchsorted

The reference tree is read-only, so `vision3d._C` cannot be dropped next to it; sys.modules injection
is the packaging. `yacs` / `visdom` are third-party config/plot packages absent from this
image; minimal stand-ins are registered ONLY if the real ones cannot be imported.
"""
import importlib
import sys
import types


def _standin_yacs():
    import copy

    import yaml

    class CfgNode(dict):
        def __init__(self, init=None):
            super().__init__()
            for k, v in (init or {}).items():
                self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

        def clone(self):
            return copy.deepcopy(self)

        def merge_from_other_cfg(self, other):
            for k, v in other.items():
                if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                    self[k].merge_from_other_cfg(v)
                else:
                    self[k] = CfgNode(v) if isinstance(v, dict) else v

        def merge_from_file(self, path):
            with open(path) as f:
                self.merge_from_other_cfg(yaml.safe_load(f) or {})

        def merge_from_list(self, lst):
            for k, v in zip(lst[0::2], lst[1::2]):
                node = self
                parts = k.split(".")
                for p in parts[:-1]:
                    node = node[p]
                node[parts[-1]] = v

        def freeze(self):
            pass

        def defrost(self):
            pass

    yacs = types.ModuleType("yacs")
    cfgmod = types.ModuleType("yacs.config")
    cfgmod.CfgNode = CfgNode
    yacs.config = cfgmod
    return {"yacs": yacs, "yacs.config": cfgmod}


def _standin_visdom():
    m = types.ModuleType("visdom")

    class Visdom:
        def __init__(self, *a, **k):
            pass

        def line(self, *a, **k):
            return None

    m.Visdom = Visdom
    return {"visdom": m}


def install(standins=True):
    from . import c_ext, pointnet2, spconv, torchsearchsorted
    from .pointnet2 import pointnet2_modules, pointnet2_utils
    from .spconv import utils as spconv_utils
    sys.modules["vision3d._C"] = c_ext
    sys.modules["spconv"] = spconv
    sys.modules["spconv.utils"] = spconv_utils
    sys.modules["pointnet2"] = pointnet2
    sys.modules["pointnet2.pointnet2_utils"] = pointnet2_utils
    sys.modules["pointnet2.pointnet2_modules"] = pointnet2_modules
    sys.modules["torchsearchsorted"] = torchsearchsorted
    if standins:
        for name, factory in (("yacs", _standin_yacs), ("visdom", _standin_visdom)):
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules.update(factory())
    return sys.modules["vision3d._C"]
