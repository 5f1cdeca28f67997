"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference (``/root/reference``).

Used by ``oracle/make_golden.py`` (fixture generation) and by the ``-m "not gpu"`` tests that pin the
oracle against the reference in the build container.  ``/root/reference`` does not exist on the GPU
box, so every caller must cope with ``load_reference()`` returning ``None``.

The reference parses its CLI at import time (uest_seg_multi_os.py:305) and drags in matplotlib /
skimage through data_loader/segmentation/utils.py:3-8 for superpixel code that the hot path never
calls; both are shimmed here (recipe from SURVEY.md section 8c).  Nothing is copied from the
reference: its modules are imported in place and only ever *called*.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MSPL_REFERENCE_ROOT", "/root/reference")

_cached = None


class RefModules:
    """Handles to the reference modules that own the hot path."""

    def __init__(self, uest, seg_loss, greenhouse, utils):
        self.uest = uest            # uest_seg_multi_os.py  (get_output, merge_outputs, ...)
        self.seg_loss = seg_loss    # loss_fns/segmentation_loss.py (PixelwiseKLD, UncertaintyWeighted...)
        self.greenhouse = greenhouse  # data_loader/segmentation/greenhouse.py (LUTs)
        self.utils = utils          # utilities/utils.py (import_os_model)


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "uest_seg_multi_os.py"))


def load_reference():
    """Import the reference's hot-path modules, or return None when the tree is absent."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        return None
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.data", "skimage.color",
                 "skimage.filters", "skimage.util", "skimage.segmentation"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    saved_argv = sys.argv
    sys.argv = ["uest_seg_multi_os.py"]
    try:
        import uest_seg_multi_os as uest
        from loss_fns import segmentation_loss as seg_loss
        from data_loader.segmentation import greenhouse
        from utilities import utils
    finally:
        sys.argv = saved_argv
    # values main() would have set (uest_seg_multi_os.py:383) and that merge_outputs/get_output read
    uest.args.classes = 5
    uest.args.use_depth = False
    _cached = RefModules(uest, seg_loss, greenhouse, utils)
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
