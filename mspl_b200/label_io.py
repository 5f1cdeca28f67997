"""Label-map output: uint8 maps -> mode-L PNG files + the ``tgt_train.lst`` CSV, off the critical path.

The reference writes one PNG per image synchronously inside its generation loop
(``Image.fromarray(amax_output.astype(np.uint8)).save(...)``, uest_seg_multi_os.py:929-931) and the dataset reads them
back with ``Image.open`` (data_loader/segmentation/greenhouse.py:232-234).  After the fused kernels that encode is the
remaining per-image serial cost, so here

  * the device->host copy of a batch of label maps goes to a pinned staging buffer on a side stream, and
  * PNG encoding + file writes run on a small thread pool (zlib releases the GIL), while the GPU fuses the next batch.

The files are ordinary 8-bit greyscale PNGs (filter 0, one IDAT); any PNG reader, PIL included, decodes them to the very
same array the reference's writer would have produced.
"""
import struct
import threading
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_PNG_SIGNATURE = b"\x89PNG\r\n\x1a\n"


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png_gray8(arr, level=1):
    """Encode a (H, W) uint8 array as an 8-bit greyscale PNG (colour type 0, no interlace, filter type 0 on every row)."""
    arr = np.ascontiguousarray(arr)
    if arr.dtype != np.uint8 or arr.ndim != 2:
        raise ValueError("expected a 2-D uint8 array, got %s %s" % (arr.dtype, arr.shape))
    h, w = arr.shape
    raw = np.zeros((h, w + 1), dtype=np.uint8)      # leading filter byte 0 (None) per scanline
    raw[:, 1:] = arr
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)
    return _PNG_SIGNATURE + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b"")


class LabelWriter:
    """Asynchronous sink for batches of label maps living on the GPU.

    ``submit(label_u8, paths)`` enqueues a device->host copy on a private stream (ordered after the producer stream) and
    returns immediately; worker threads wait for the copy, encode and write.  ``close()`` drains everything and re-raises
    the first worker error."""

    def __init__(self, device, workers=8, slots=4, level=1):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.level = level
        self.slots = threading.Semaphore(slots)      # bounds pinned staging memory
        self.futures = []

    def submit(self, label_u8, paths):
        if label_u8.dtype != torch.uint8 or not label_u8.is_cuda:
            raise ValueError("label maps must be uint8 CUDA tensors")
        if len(paths) != label_u8.shape[0]:
            raise ValueError("one path per label map")
        self.slots.acquire()
        host = torch.empty(label_u8.shape, dtype=torch.uint8, pin_memory=True)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            host.copy_(label_u8, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        label_u8.record_stream(self.stream)
        state = {"left": len(paths), "lock": threading.Lock()}
        if not paths:
            self.slots.release()
        for i, path in enumerate(paths):      # one task per image: a batch's maps are encoded in parallel
            self.futures.append(self.pool.submit(self._write_one, host, i, done, path, state))

    def _write_one(self, host, index, done, path, state):
        try:
            done.synchronize()
            with open(path, "wb") as f:
                f.write(encode_png_gray8(host[index].numpy(), self.level))
        finally:
            with state["lock"]:
                state["left"] -= 1
                last = state["left"] == 0
            if last:
                self.slots.release()

    def close(self):
        try:
            for fut in self.futures:
                fut.result()
        finally:
            self.futures = []
            self.pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
