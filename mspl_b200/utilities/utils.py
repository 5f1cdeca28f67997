"""Drop-in replacement for the hot half of the reference's utilities/utils.py:in_training_visualization_img (:76-133), the
TensorBoard hook train()/val() call every epoch (uest_seg_multi_os.py:1074-1077, 1224-1227).

The reference computes ``PixelwiseKLD`` with library passes, synchronises for ``torch.max(kld).item()``, moves every map to the
CPU and colours the label maps class by class with ``masked_fill_`` per image (LongTensorToRGBPIL, :188-237).  Here the argmax of
``main + 0.5*aux``, the KLD map, its maximum, the ``-kld/max + 1`` heat map and the label colouring all run on the device
(``ops.prediction_maps``, ``ops.label_colors``); only the finished grids are copied to the host.  Same signature, same
``writer.add_image`` tags, order and array layouts."""
from collections import OrderedDict

import torch
import torchvision

from .. import ops


def _split_predictions(predictions):
    """(main, aux or None) / a ready (N,H,W) label map, following the branches at utilities/utils.py:88-114."""
    if type(predictions) is tuple:
        return predictions[0], predictions[1], None
    if isinstance(predictions, OrderedDict):
        return predictions['out'], (predictions['aux'] if len(predictions) == 2 else None), None
    if len(predictions.size()) == 3:
        return None, None, predictions
    return predictions, None, None


def prediction_maps(predictions):
    """-> (pred_labels int64 (N,H,W) on the device, heat f32 (N,1,H,W) or None): what the reference derives at :88-114."""
    main, aux, ready = _split_predictions(predictions)
    if ready is not None:
        return ready, None
    return ops.prediction_maps(main, aux)


def in_training_visualization_img(model, images, depths=None, labels=None, predictions=None, class_encoding=None, writer=None,
                                  epoch=None, data=None, device=None):
    if predictions is None:
        model.eval()
        with torch.no_grad():
            predictions = model(images, depths) if depths is not None else model(images)
    pred_labels, heat = prediction_maps(predictions)
    if heat is not None:
        writer.add_image(data + '/kld', torchvision.utils.make_grid(heat.cpu()).numpy(), epoch)
    colors = [tuple(c) for c in class_encoding.values()]
    dev = pred_labels.device
    color_train = ops.label_colors(labels.to(dev), colors).cpu() if labels is not None else None
    color_predictions = ops.label_colors(pred_labels, colors).cpu()
    # write_summary_batch (utilities/utils.py:172-186)
    writer.add_image(data + '/images', torchvision.utils.make_grid(images.data.cpu()).numpy(), epoch)
    if color_train is not None:
        writer.add_image(data + '/train_labels', torchvision.utils.make_grid(color_train).numpy(), epoch)
    writer.add_image(data + '/pred_labels', torchvision.utils.make_grid(color_predictions).numpy(), epoch)
