"""Drop-in replacement for the reference's utilities/metrics/segmentation_miou.py:MIOU, computed on the GPU.

The reference moves ``pred``/``target`` to the CPU on every call and runs ``torch.histc`` three times
(call sites uest_seg_multi_os.py:1032, 1198; eval_label.py:201); here one kernel counts the per-class intersection,
prediction and mask areas where the tensors already live.  Same constructor, same ``get_iou`` return value:
``(area_inter, area_union)`` as float32 NumPy arrays of length ``num_classes``."""
import numpy as np

from ... import ops


class MIOU(object):
    def __init__(self, num_classes=21):
        self.num_classes = num_classes
        self.epsilon = 1e-6

    def get_iou(self, output, target):
        if isinstance(output, tuple):
            output = output[0]
        counts = ops.miou_counts(output, target, self.num_classes).cpu().numpy()
        area_inter = counts[0].astype(np.float32)
        area_union = (counts[1].astype(np.float32) + counts[2].astype(np.float32) - area_inter + np.float32(self.epsilon))
        return area_inter, area_union.astype(np.float32)
