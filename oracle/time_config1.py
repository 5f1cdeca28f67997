#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE -- BASELINE.json configs[0]: 1-source (random-init 20-class ESPDNetUE) pseudo-label generation
on 8 synthetic 480x256 images through the LIVE reference's own CPU path (needs /root/reference: the network definition is not
staged under oracle/_ref).  Prints one JSON line: seconds per image for the network forward and for the post-network path this
repo replaces (softmax + KLD + host copies inside get_output, argmax, table, merge_outputs, class_array), and Mpix/s of both."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402


def main():
    ref = ref_import.load_reference()
    if ref is None or ref.utils is None:
        print(json.dumps({"unavailable": "/root/reference not present"}))
        return
    U = ref.uest
    model = ref_import.build_espdnetue(20, seed=3)
    g = torch.Generator().manual_seed(3)
    images = [torch.randn(1, 3, 256, 480, generator=g) for _ in range(8)]
    table = ref.greenhouse.id_cityscapes_to_greenhouse
    class_array = np.zeros(5)
    t_total = t_net = 0.0
    with torch.no_grad():
        model(images[0])                                       # warm-up
        for im in images:
            t0 = time.perf_counter()
            out = model(im)
            t1 = time.perf_counter()
            t_net += t1 - t0
            output, _ = U.get_output(ref_import.FixedLogitsModel(out[0], out[1]), im, device='cpu')
            amax = np.asarray(np.argmax(output.transpose(1, 2, 0), axis=2), dtype=np.uint8)
            lab = U.merge_outputs(np.array([table[amax]]), seg_classes=5, thresh='all')
            for k in range(5):
                class_array[k] += (lab == k).sum()
            t_total += time.perf_counter() - t0
    mpix = 8 * 256 * 480 / 1e6
    print(json.dumps({"config": "configs[0]: ESPDNetUE (20 classes, random init, seed 3), 8 x 480x256, CPU reference path",
                      "torch_threads": torch.get_num_threads(), "host_cpus": os.cpu_count(),
                      "network_forward_s_per_image": round(t_net / 8, 4),
                      "post_network_path_s_per_image": round((t_total - t_net) / 8, 4),
                      "post_network_path_mpix_s": round(mpix / (t_total - t_net), 2),
                      "end_to_end_mpix_s": round(mpix / t_total, 2), "class_array": class_array.tolist()}))


if __name__ == "__main__":
    main()
