"""TEST INFRASTRUCTURE — not part of the product path.

Loads the *unmodified* reference hot-path modules behind namespace shells: from /root/reference in the build
container (golden vectors), or from the staged copy baseline/_ref (oracle/install_ref.py; git-ignored, it travels to
the GPU box) for `bench.py --impl reference`, `cpu_baseline` and the stock-PyTorch GPU probe.
Recipe follows SURVEY.md Appendix B.  Only `oracle/make_*.py`, `tests/` and `bench.py`'s reference legs import this.
"""
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REF_ROOT = os.environ.get("RTPOSE_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/det3d") else _STAGED)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "det3d", "models"))


class AttrDict(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


_loaded = None


def load():
    """Returns the reference's `build_detector` with the four registry names populated."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    R = REF_ROOT
    for name in [n for n in sys.modules if n == "det3d" or n.startswith("det3d.")]:
        del sys.modules[name]  # e.g. aliases left by rtpose_b200.det3d_compat.install_as_det3d() earlier in this process

    def ns(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        m.__package__ = name
        sys.modules[name] = m
        return m

    ns("yacs")
    yc = ns("yacs.config")

    class CfgNode(dict):
        __getattr__ = dict.__getitem__

        def __setattr__(s, k, v):
            s[k] = v

    yc.CfgNode = CfgNode
    pc = ns("pycocotools")
    pc.mask = ns("pycocotools.mask")
    d = ns("det3d", R + "/det3d")
    t = ns("det3d.torchie", R + "/det3d/torchie")
    t.is_str = lambda x: isinstance(x, str)
    d.torchie = t
    ns("det3d.torchie.cnn", R + "/det3d/torchie/cnn")
    sys.modules["det3d.torchie.cnn"].kaiming_init = importlib.import_module(
        "det3d.torchie.cnn.weight_init"
    ).kaiming_init
    ns("det3d.torchie.trainer").load_checkpoint = None
    u = ns("det3d.utils", R + "/det3d/utils")
    reg = importlib.import_module("det3d.utils.registry")
    u.Registry, u.build_from_cfg = reg.Registry, reg.build_from_cfg
    ns("det3d.core", R + "/det3d/core").box_torch_ops = None
    ns("det3d.core.utils", R + "/det3d/core/utils")
    m = ns("det3d.models", R + "/det3d/models")
    for s in ["backbones", "backbones.hr_util", "pose_heads", "losses", "detectors", "readers"]:
        ns("det3d.models." + s, R + "/det3d/models/" + s.replace(".", "/"))
    mu = ns("det3d.models.utils", R + "/det3d/models/utils")
    mu.Sequential = importlib.import_module("det3d.models.utils.misc").Sequential
    ns("det3d.models.utils.finetune_utils").FrozenBatchNorm2d = object
    ns("det3d.ops")
    m.builder = importlib.import_module("det3d.models.builder")
    for s in ["readers.radar_encoder", "backbones.hrnet3d", "pose_heads.center_head", "detectors.radar_pose_net"]:
        importlib.import_module("det3d.models." + s)
    from det3d.models.builder import build_detector  # noqa

    mods = {
        "build_detector": build_detector,
        "center_utils": importlib.import_module("det3d.core.utils.center_utils"),
        "centernet_loss": importlib.import_module("det3d.models.losses.centernet_loss"),
    }
    # the reference keeps the `det3d.*` names in sys.modules (it does call-time relative imports);
    # the product package lives under `rtpose_b200.*`, so both can coexist in one test process.
    _loaded = mods
    return mods
