"""Criteo "DAC" batch loader with device-side marshalling.

Reference: `DACRecord` (src/data/criteo.jl:91-95), `load(::DAC, path)` (:113-117, an mmap of the
preprocessed record file), `DACLoader` / `load!` (:284-344).  The reference transposes every batch on
CPU threads; here a batch's raw records go to the GPU as one contiguous copy from pinned memory and
are unpacked by `dlrmb_dac_unpack` into labels [B], dense [B][13], sparse [26][B][1], double-buffered
on a copy stream so the host-to-device transfer of batch i+1 overlaps step i.
"""
from __future__ import annotations

from typing import Iterator, Tuple

import numpy as np
import torch

from . import _lib

# struct DACRecord: label::Int32, continuous::NTuple{13,Float32}, categorical::NTuple{26,UInt32}
DAC_DTYPE = np.dtype([("label", "<i4"), ("continuous", "<f4", (13,)), ("categorical", "<u4", (26,))])
assert DAC_DTYPE.itemsize == 160


def load(path: str, writable: bool = False) -> np.ndarray:
    """``load(DAC(), path)``: memory-map a preprocessed record file."""
    return np.memmap(path, dtype=DAC_DTYPE, mode="r+" if writable else "r")


class DACLoader:
    """``DACLoader(dataset, batchsize)`` (src/data/criteo.jl:312-344): iterates whole batches and
    yields device tensors ``(labels [B] f32, dense [B][13] f32, sparse [26][B][1] int32)``.

    Two pinned staging buffers and two device record buffers; the copy of the next batch is issued
    on a side stream while the caller works on the current one.
    """

    def __init__(self, dataset: np.ndarray, batchsize: int, device=0, idx_base: int = 1):
        """``idx_base``: base of the categorical ids in ``dataset``.  Files written by the reference's
        preprocessing are 1-based (``reindex!`` assigns ``length(dict) + 1``, src/data/criteo.jl:249-253),
        hence the default; the ids are passed through unchanged and the base travels with the loader
        (``train`` adopts it, ``DLRMModel(..., idx_base=loader.idx_base)``)."""
        if idx_base not in (0, 1):
            raise ValueError("idx_base must be 0 or 1")
        self.idx_base = int(idx_base)
        if dataset.dtype != DAC_DTYPE:
            raise TypeError("dataset must be an array of DACRecord (loader.DAC_DTYPE)")
        self.dataset = dataset
        self.batchsize = int(batchsize)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DLRMB200Error(_lib.EINVAL, "DACLoader unpacks batches on the GPU: a CUDA device is required")
        self._lib = _lib.load()
        nbytes = self.batchsize * DAC_DTYPE.itemsize
        self._pinned = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._raw = [torch.empty(nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(self.device)
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]

    def __len__(self) -> int:
        return len(self.dataset) // self.batchsize

    def _stage(self, i: int) -> None:
        """Host copy into pinned memory + async H2D + unpack, all for batch i, on the copy stream."""
        slot = i & 1
        B = self.batchsize
        rec = self.dataset[i * B:(i + 1) * B]
        self._consumed[slot].synchronize()          # the buffers of batch i-2 are free again
        self._pinned[slot].numpy()[:] = np.frombuffer(rec.tobytes(), dtype=np.uint8)
        with torch.cuda.stream(self._copy_stream):
            self._raw[slot].copy_(self._pinned[slot], non_blocking=True)
            labels = torch.empty(B, dtype=torch.float32, device=self.device)
            dense = torch.empty((B, 13), dtype=torch.float32, device=self.device)
            sparse = torch.empty((26, B, 1), dtype=torch.int32, device=self.device)
            _lib.check(self._lib.dlrmb_dac_unpack(
                self.device.index or 0, self._raw[slot].data_ptr(), B, labels.data_ptr(), dense.data_ptr(),
                sparse.data_ptr(), self._copy_stream.cuda_stream))
            self._ready[slot].record(self._copy_stream)
        self._staged = (labels, dense, sparse)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        n = len(self)
        if n == 0:
            return
        for ev in self._consumed:
            ev.record()
        self._stage(0)
        for i in range(n):
            slot = i & 1
            batch = self._staged
            if i + 1 < n:
                self._stage(i + 1)                  # overlaps the caller's work on batch i
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[slot])
            for t in batch:
                t.record_stream(cur)
            yield batch
            self._consumed[slot].record(cur)
