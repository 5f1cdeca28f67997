"""CPU: the PNG writer produces files PIL decodes to the identical uint8 array (what the reference's dataset loader does,
data_loader/segmentation/greenhouse.py:232-234)."""
import io

import numpy as np
import pytest
from PIL import Image

from mspl_b200.label_io import encode_png_gray8


@pytest.mark.parametrize("shape", [(256, 480), (1, 1), (7, 13), (512, 1024)])
def test_png_roundtrip(shape):
    rng = np.random.default_rng(shape[0])
    arr = rng.integers(0, 5, shape, dtype=np.uint8)
    for level in (0, 1, 6):
        im = Image.open(io.BytesIO(encode_png_gray8(arr, level)))
        assert im.mode == "L" and im.size == (shape[1], shape[0])
        assert np.array_equal(np.array(im), arr)
    full = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(np.array(Image.open(io.BytesIO(encode_png_gray8(full)))), full)


def test_png_same_pixels_as_reference_writer():
    arr = (np.arange(64 * 96, dtype=np.int64).reshape(64, 96) % 5)
    ref = io.BytesIO()
    Image.fromarray(arr.astype(np.uint8)).save(ref, format="PNG")          # uest_seg_multi_os.py:929-931
    a = np.array(Image.open(io.BytesIO(ref.getvalue())))
    b = np.array(Image.open(io.BytesIO(encode_png_gray8(arr.astype(np.uint8)))))
    assert np.array_equal(a, b)


def test_png_rejects_wrong_input():
    with pytest.raises(ValueError):
        encode_png_gray8(np.zeros((4, 4), dtype=np.int64))
    with pytest.raises(ValueError):
        encode_png_gray8(np.zeros((2, 4, 4), dtype=np.uint8))
