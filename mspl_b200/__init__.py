"""mspl_b200 -- B200-native (sm_100a) implementation of MSPL's multi-source pseudo-label generation path and its
uncertainty-weighted loss, behind the reference's own Python call surface.

  mspl_b200.uest_seg_multi_os            get_output, merge_outputs, generate_pseudo_label[_multi_model], ...
  mspl_b200.loss_fns.segmentation_loss   PixelwiseKLD, UncertaintyWeightedSegmentationLoss
  mspl_b200.data_loader.segmentation.greenhouse   id_{camvid,cityscapes,forest}_to_greenhouse
  mspl_b200.ops                          batched tensor fast paths (fuse_sources, cb_thresholds, apply_thresholds, uw_ce_loss)
  mspl_b200.pipeline                     LabelGenerator: whole-job label generation, one process per GPU, histogram all-reduce

All compute goes through libmspl_b200.so (include/mspl_b200.h); there is no CPU fallback.
"""
__version__ = "0.2.0"
