#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the reference's own hot-path files for the GPU box.

    python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference is a flat tree of Python scripts with no setup.py, so it cannot be pip-installed, and /root/reference does not
exist on the GPU box.  This recipe copies the THREE files that hold the path,

    uest_seg_multi_os.py                        get_output :669-693, merge_outputs :695-718, the generation loop :888-950
    loss_fns/segmentation_loss.py               PixelwiseKLD :177-189, UncertaintyWeightedSegmentationLoss :146-175
    data_loader/segmentation/greenhouse.py      id_*_to_greenhouse :15-58

byte for byte into oracle/_ref/ (git-ignored: never committed; NOT gpurun-ignored, so it travels to the GPU box like the built
.so files) together with empty package markers.  oracle/ref_import.py imports them from there with stub modules standing in for
everything else those files import at module level (the networks, transforms, TensorBoard ... -- none of it on the path), so
bench.py's CPU arm can time the reference's OWN code (`cpu_baseline.kind: "reference"`).  Nothing here is used by the product.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MSPL_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = ("uest_seg_multi_os.py", "loss_fns/segmentation_loss.py", "data_loader/segmentation/greenhouse.py")
PACKAGES = ("loss_fns", "data_loader", "data_loader/segmentation")


def build(verbose=True):
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, FILES[0])):
        if verbose:
            print("build_ref: %s not found; keeping whatever oracle/_ref already holds" % REFERENCE_ROOT)
        return False
    manifest = []
    for rel in FILES:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFERENCE_ROOT, rel), dst)
        manifest.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), rel))
    for pkg in PACKAGES:
        open(os.path.join(DEST, pkg, "__init__.py"), "w").close()        # empty markers (the reference's own import other loaders)
    with open(os.path.join(DEST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print("build_ref: staged %d reference files under %s" % (len(FILES), DEST))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
