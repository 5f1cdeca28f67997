"""TEST / BASELINE INFRASTRUCTURE ONLY -- loader for the *live* reference.

Two places hold it:
  * ``/root/reference`` (the build container): the whole tree, imported in place.  Used by ``oracle/make_golden.py`` (fixture
    generation) and by the ``-m "not gpu"`` tests that pin the oracle against the reference.
  * ``oracle/_ref`` (staged by ``oracle/build_ref.py``, git-ignored, travels to the GPU box): only the three files that hold the
    hot path.  Everything else they import at module level is replaced by stub modules.  Used by ``bench.py``'s CPU arm so that
    the baseline it times is the reference's OWN code.
``load_reference()`` returns ``None`` when neither is present; every caller must cope with that.

The reference parses its CLI at import time (uest_seg_multi_os.py:305) and drags in matplotlib / skimage through
data_loader/segmentation/utils.py:3-8 for superpixel code that the hot path never calls; both are shimmed here (recipe from
SURVEY.md section 8c).  Its modules are imported and only ever *called*.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MSPL_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

_cached = None


class RefModules:
    """Handles to the reference modules that own the hot path."""

    def __init__(self, uest, seg_loss, greenhouse, utils, origin):
        self.uest = uest            # uest_seg_multi_os.py  (get_output, merge_outputs, ...)
        self.seg_loss = seg_loss    # loss_fns/segmentation_loss.py (PixelwiseKLD, UncertaintyWeighted...)
        self.greenhouse = greenhouse  # data_loader/segmentation/greenhouse.py (LUTs)
        self.utils = utils          # utilities/utils.py (import_os_model); None when loaded from oracle/_ref
        self.origin = origin        # "tree" (/root/reference) or "staged" (oracle/_ref)


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "uest_seg_multi_os.py"))


def staged_available():
    return os.path.isfile(os.path.join(STAGED_ROOT, "uest_seg_multi_os.py"))


class _Stub(types.ModuleType):
    """Stands in for a module the staged files import but the hot path never calls: any attribute is a do-nothing class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        value = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, value)
        return value


def _install_stubs(names):
    for name in names:
        if name in sys.modules:
            continue
        mod = _Stub(name)
        mod.__path__ = []       # importable as a package, so that `import a.b.c` resolves through sys.modules
        sys.modules[name] = mod
        parent, _, child = name.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], child, mod)


def _import_hot_path(root, origin):
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.data", "skimage.color",
                 "skimage.filters", "skimage.util", "skimage.segmentation"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if origin == "staged":
        # the rest of the reference tree is absent: stub what the three files import at module level
        _install_stubs(("transforms", "transforms.segmentation", "transforms.segmentation.data_transforms", "utilities",
                        "utilities.utils", "utilities.metrics", "utilities.metrics.segmentation_miou",
                        "data_loader.segmentation.utils"))
        for optional in ("cv2", "scipy.io", "tqdm", "torch.utils.tensorboard", "torchvision.models", "torchvision.transforms"):
            try:
                __import__(optional)
            except Exception:
                _install_stubs((optional,))
    if root not in sys.path:
        sys.path.insert(0, root)
    saved_argv = sys.argv
    sys.argv = ["uest_seg_multi_os.py"]
    try:
        import uest_seg_multi_os as uest
        from loss_fns import segmentation_loss as seg_loss
        from data_loader.segmentation import greenhouse
        utils = None
        if origin == "tree":
            from utilities import utils
    finally:
        sys.argv = saved_argv
    # values main() would have set (uest_seg_multi_os.py:383) and that merge_outputs/get_output read
    uest.args.classes = 5
    uest.args.use_depth = False
    return RefModules(uest, seg_loss, greenhouse, utils, origin)


def load_reference(allow_staged=False):
    """Import the reference's hot-path modules from /root/reference, or (allow_staged) from oracle/_ref; None when absent."""
    global _cached
    if _cached is not None:
        return _cached
    if reference_available():
        _cached = _import_hot_path(REFERENCE_ROOT, "tree")
    elif allow_staged and staged_available():
        _cached = _import_hot_path(STAGED_ROOT, "staged")
    return _cached


class FixedLogitsModel:
    """A 'model' that ignores its input and returns preset (main, aux) logits -- lets the reference's
    get_output consume the very same synthetic logits the kernels see, with no CNN in the way."""

    def __init__(self, main, aux):
        self.main, self.aux = main, aux

    def __call__(self, image):
        return (self.main, self.aux)


def build_espdnetue(num_classes, seed=3):
    """Random-init ESPDNetUE source model exactly as main() builds it (uest_seg_multi_os.py:426-435
    -> utilities/utils.py:274-298); BASELINE config 1 uses num_classes=20, seed 3."""
    import copy
    import torch
    ref = load_reference()
    args = copy.deepcopy(ref.uest.args)
    args.weights = ''
    args.s = 2.0
    args.channels = 3
    args.num_classes = 1000
    torch.manual_seed(seed)
    model = ref.utils.import_os_model(args, 'espdnetue', '', num_classes)
    model.eval()
    return model
