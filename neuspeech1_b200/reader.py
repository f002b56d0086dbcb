"""On-device input pipeline: `.npy` recording -> channel selection / channel padding -> pinned staging
-> asynchronous copy into one of TWO persistent device batches -> the augmentation pass pads to 30 s, casts and lays the batch
out channels-last on the device.

Reference behaviour restated: `utils/reader.py:253-303` (`np.load`; Schoffelen recordings keep rows 28:301, Gwilliams rows :208,
anything else rows :modal_ch; fewer channels than `modal_ch` are zero-padded at the END, `:508-516`), `:496-506` (crop to
30 s x 200 Hz; the zero tail is written by `ns_aug_pass`, not on the host), and the collator `utils/data_utils.py:185-221`
(labels right-padded with -100; a leading BOS shared by every row is cut).  Tokenisation stays with the caller: items carry
token ids.

The reference ships padded fp32 batches (5 MB per sample at C = 208) from 16 numpy workers; here the host only touches the
UNPADDED samples (mean length ~2700 of 6000), and because batches land in two fixed device buffers the training step can be
replayed as a CUDA graph (`engine.train_step(use_graph=True)` keys its graphs on the buffer addresses).
"""
from __future__ import annotations

import queue
import threading
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .augment_eeg import BatchAugmenter


def select_channels(sample: np.ndarray, path: str, modal_ch: int) -> np.ndarray:
    """(C_file, n) -> (modal_ch, n): dataset-specific row window, then zero rows appended up to modal_ch (reader.py:270-282)."""
    if "schoffelen" in path:
        sample = sample[28:301]
    elif "gwilliams" in path:
        sample = sample[:208]
    else:
        sample = sample[:modal_ch]
    if sample.shape[0] > modal_ch:
        raise ValueError(f"{path}: {sample.shape[0]} channels selected but the stem takes {modal_ch}")
    if sample.shape[0] < modal_ch:
        sample = np.pad(sample, ((0, modal_ch - sample.shape[0]), (0, 0)))
    return sample


def collate_labels(labels: Sequence[Sequence[int]], bos_token_id: Optional[int] = None) -> torch.Tensor:
    """Right-pad with -100; drop the first column when every row starts with BOS (data_utils.py:198-219)."""
    L = max(len(l) for l in labels)
    out = torch.full((len(labels), L), -100, dtype=torch.long)
    for i, l in enumerate(labels):
        out[i, :len(l)] = torch.as_tensor(list(l), dtype=torch.long)
    if bos_token_id is not None and bool((out[:, 0] == bos_token_id).all()):
        out = out[:, 1:]
    return out


class DeviceBatchLoader:
    """Iterate `(input_features, labels, aug)` batches that live on `device`.

    items: sequence of dicts `{"path": <.npy file> | "array": (C, n) ndarray, "labels": [token ids]}`.
    `input_features` is one of two persistent (B, modal_ch, max_samples) fp32 device buffers holding the unpadded samples (rows
    beyond a sample's length are stale: `aug["n"]` carries the lengths and `ns_aug_pass` writes the zero tail), `labels` one of two
    persistent (B, max_label_len) int64 buffers (-100 padded; longer label rows are truncated, the reference filters them by
    `max_label_length`), `aug` the keyword arguments of `engine.train_step(..., aug=aug)` (augmentation decisions drawn per sample
    with the reference's RNG calls when `augment_configs` is given), written into persistent per-slot tensors as well: every
    address a training step sees repeats every second batch, which is what CUDA-graph replay of the step needs.
    A background thread reads and stages batch i+1 while batch i trains; the device copy runs on its own stream."""

    def __init__(self, items: Sequence[Dict], batch_size: int, modal_ch: int, device, augment_configs: Optional[Dict] = None,
                 max_duration: float = 30.0, sample_rate: int = 200, max_label_len: int = 64, bos_token_id: Optional[int] = None,
                 order: Optional[Iterable[int]] = None, drop_last: bool = True, train: bool = True):
        self.items, self.B, self.C = items, batch_size, modal_ch
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.T = int(max_duration * sample_rate)
        self.Lmax = max_label_len
        self.bos = bos_token_id
        self.order = list(order) if order is not None else list(range(len(items)))
        self.drop_last = drop_last
        self.aug = BatchAugmenter(augment_configs or {}, max_duration=max_duration, sample_rate=sample_rate, train=train)
        self.x_dev = [torch.zeros(batch_size, modal_ch, self.T, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.y_dev = [torch.full((batch_size, max_label_len), -100, dtype=torch.long, device=self.device) for _ in range(2)]
        self.x_host = [torch.zeros(batch_size, modal_ch, self.T, dtype=torch.float32, pin_memory=self.cuda) for _ in range(2)]
        self.aug_static = [self.aug.static_buffers(batch_size, modal_ch, self.device) for _ in range(2)]
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.free = [torch.cuda.Event() for _ in range(2)] if self.cuda else None     # device side: the step that read the slot was launched
        self.h2d_done = [torch.cuda.Event() for _ in range(2)] if self.cuda else None # host side: the pinned slot may be refilled
        # host-side handshake per pinned slot: the consumer hands the slot back only AFTER it has enqueued the copy out of it and
        # recorded h2d_done (an event that was never recorded, or was recorded for an older batch, synchronises at once)
        self.host_free = [threading.Semaphore(1) for _ in range(2)]
        self._consumer_delay = 0.0                     # tests: sleep this long between q.get() and the copy (widens the old race)

    def __len__(self) -> int:
        n = len(self.order)
        return n // self.B if self.drop_last else (n + self.B - 1) // self.B

    def _stage(self, idxs: List[int], slot: int):
        """Host side of one batch (worker thread): read, select channels, crop to 30 s, copy into the pinned buffer."""
        lens, labels = [], []
        xh = self.x_host[slot]
        self.host_free[slot].acquire()                  # the consumer has issued the copy of the batch staged here two batches ago
        if self.cuda:
            self.h2d_done[slot].synchronize()           # ... and that copy has left the pinned buffer
        for b, i in enumerate(idxs):
            it = self.items[i]
            s = it["array"] if "array" in it else np.load(it["path"])
            s = select_channels(np.asarray(s), it.get("path", ""), self.C)[:, :self.T]
            n = s.shape[1]
            xh[b, :, :n] = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32))
            lens.append(n); labels.append(list(it["labels"])[:self.Lmax])
        return lens, collate_labels(labels, self.bos)

    def release(self, slot: int):
        """The work that reads `slot` has been launched on the current stream: its device buffers may be overwritten after it.
        Called automatically when the next batch is requested."""
        if self.cuda:
            self.free[slot].record(torch.cuda.current_stream(self.device))

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, dict, int]]:
        batches = [self.order[k:k + self.B] for k in range(0, len(self.order), self.B)]
        if self.drop_last:
            batches = [b for b in batches if len(b) == self.B]
        q: "queue.Queue" = queue.Queue(maxsize=1)

        def worker():
            for k, idxs in enumerate(batches):
                q.put((k, idxs, self._stage(idxs, k & 1)))
            q.put(None)

        for sem in self.host_free:                      # a previous, abandoned iteration may have left a slot taken
            while sem.acquire(blocking=False):
                pass
            sem.release()
        th = threading.Thread(target=worker, daemon=True)
        th.start()
        prev_slot = None
        while True:
            got = q.get()
            if self._consumer_delay:
                import time
                time.sleep(self._consumer_delay)
            if prev_slot is not None:
                self.release(prev_slot)                 # the consumer launched its step on the previous batch before asking again
            if got is None:
                break
            k, idxs, (lens, labels) = got
            slot = k & 1
            nb = len(idxs)
            xd, yd = self.x_dev[slot], self.y_dev[slot]
            lab = torch.full((nb, self.Lmax), -100, dtype=torch.long, pin_memory=self.cuda)
            lab[:, :labels.shape[1]] = labels
            shapes = [(self.C, n) for n in lens]
            if self.cuda:
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(self.free[slot])
                    xd[:nb].copy_(self.x_host[slot][:nb], non_blocking=True)      # whole rows: one contiguous pinned -> device copy
                    yd[:nb].copy_(lab, non_blocking=True)
                    plan = self.aug.plan(shapes, self.device, static=self.aug_static[slot], x=xd[:nb])
                    self.h2d_done[slot].record(self.stream)
                torch.cuda.current_stream(self.device).wait_stream(self.stream)
            else:
                xd[:nb].copy_(self.x_host[slot][:nb])
                yd[:nb].copy_(lab)
                plan = self.aug.plan(shapes, self.device, static=self.aug_static[slot], x=xd[:nb])
            self.host_free[slot].release()              # copy enqueued, h2d_done recorded: the worker may wait on it and refill
            # fixed-shape views of the persistent buffers: the label width is the widest batch the loader may ever yield
            prev_slot = slot
            yield xd[:nb], yd[:nb], plan, slot
        th.join()


class SampleStore:
    """The training set's recordings, UNPADDED, resident in HBM.

    The reference re-reads every `.npy` each epoch, pads it to 30 s in fp32 and ships 5 MB per sample over PCIe
    (`utils/reader.py:253-303,496-516`).  Here every recording is read ONCE, cut to its channel window and to 30 s, rounded to the
    compute dtype the stem consumes anyway, and packed back to back into one flat device tensor: channel row c of item i starts at
    element `off[i] + c * ld[i]` and holds `n[i]` samples (`ld` = n rounded up to 8 so that rows start on 16-byte boundaries).
    A batch is then three small integer vectors (`src_off`, `src_ld`, `n`): `ns_aug_pass` gathers, augments, pads, and lays the
    batch out channels-last straight from the store.  Gwilliams at C = 208 (mean length ~2700 samples): ~1.1 MB per recording in
    bf16, i.e. 100 k recordings fit in 112 GB of the 180 GB.
    """

    def __init__(self, items: Sequence[Dict], modal_ch: int, device, dtype: torch.dtype = torch.bfloat16,
                 max_duration: float = 30.0, sample_rate: int = 200, chunk_bytes: int = 256 << 20):
        self.C, self.T = modal_ch, int(max_duration * sample_rate)
        self.device, self.dtype = torch.device(device), dtype
        cuda = self.device.type == "cuda"
        arrays, off, total = [], [], 0
        self.n = np.zeros(len(items), dtype=np.int32)
        self.ld = np.zeros(len(items), dtype=np.int32)
        for i, it in enumerate(items):
            s = it["array"] if "array" in it else np.load(it["path"])
            s = select_channels(np.asarray(s), it.get("path", ""), modal_ch)[:, :self.T]
            n = s.shape[1]
            ld = (n + 7) // 8 * 8
            self.n[i], self.ld[i] = n, ld
            off.append(total)
            total += modal_ch * ld
            arrays.append(s)
        self.off = np.asarray(off, dtype=np.int64)
        self.flat = torch.zeros(max(total, 8), dtype=dtype, device=self.device)
        # upload in pinned chunks of whole recordings (conversion to `dtype` on the host: half the PCIe bytes for bf16)
        esz = self.flat.element_size()
        cap = max(chunk_bytes // esz, int((self.ld.astype(np.int64) * modal_ch).max()) if len(items) else 8)
        stage = torch.zeros(cap, dtype=dtype, pin_memory=cuda)
        start, fill = 0, 0
        for i, s in enumerate(arrays):
            sz = modal_ch * int(self.ld[i])
            if fill + sz > cap:
                self.flat[start:start + fill].copy_(stage[:fill], non_blocking=False)
                start, fill = start + fill, 0
            v = stage[fill:fill + sz].view(modal_ch, int(self.ld[i]))
            v[:, :s.shape[1]] = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)).to(dtype)
            v[:, s.shape[1]:] = 0
            fill += sz
        if fill:
            self.flat[start:start + fill].copy_(stage[:fill], non_blocking=False)
        self.bytes = total * esz

    def __len__(self) -> int:
        return len(self.n)

    def batch_tables(self, idxs: Sequence[int]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        idxs = np.asarray(idxs, dtype=np.int64)
        return self.off[idxs], self.ld[idxs], self.n[idxs]


class ResidentBatchLoader:
    """Batches over a `SampleStore`: yields `(store.flat, labels, aug, slot)` where `aug` carries the gather tables
    (`src_off`, `src_ld`, `n`) and the augmentation decisions, all in two sets of persistent device tensors (so
    `engine.train_step(use_graph=True)` replays one CUDA graph per slot).  Per batch the host moves 3 small integer vectors and
    the labels; the samples never leave HBM."""

    def __init__(self, store: SampleStore, labels: Sequence[Sequence[int]], batch_size: int, augment_configs: Optional[Dict] = None,
                 max_label_len: int = 64, bos_token_id: Optional[int] = None, order: Optional[Iterable[int]] = None,
                 drop_last: bool = True, train: bool = True, sample_rate: int = 200):
        if len(labels) != len(store):
            raise ValueError("one label row per stored recording")
        self.store, self.labels, self.B = store, labels, batch_size
        self.device, self.cuda = store.device, store.device.type == "cuda"
        self.Lmax, self.bos = max_label_len, bos_token_id
        self.order = list(order) if order is not None else list(range(len(store)))
        self.drop_last = drop_last
        self.aug = BatchAugmenter(augment_configs or {}, max_duration=store.T / sample_rate, sample_rate=sample_rate, train=train)
        dev = self.device
        self.y_dev = [torch.full((batch_size, max_label_len), -100, dtype=torch.long, device=dev) for _ in range(2)]
        self.off_dev = [torch.zeros(batch_size, dtype=torch.long, device=dev) for _ in range(2)]
        self.ld_dev = [torch.zeros(batch_size, dtype=torch.int32, device=dev) for _ in range(2)]
        self.aug_static = [self.aug.static_buffers(batch_size, store.C, dev) for _ in range(2)]

    def __len__(self) -> int:
        n = len(self.order)
        return n // self.B if self.drop_last else (n + self.B - 1) // self.B

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, dict, int]]:
        batches = [self.order[k:k + self.B] for k in range(0, len(self.order), self.B)]
        if self.drop_last:
            batches = [b for b in batches if len(b) == self.B]
        for k, idxs in enumerate(batches):
            slot, nb = k & 1, len(idxs)
            off, ld, n = self.store.batch_tables(idxs)
            lab = torch.full((nb, self.Lmax), -100, dtype=torch.long, pin_memory=self.cuda)
            rows = collate_labels([list(self.labels[i])[:self.Lmax] for i in idxs], self.bos)
            lab[:, :rows.shape[1]] = rows
            offh = torch.from_numpy(off.copy())
            ldh = torch.from_numpy(ld.copy())
            if self.cuda:
                offh, ldh = offh.pin_memory(), ldh.pin_memory()
            # stream order protects the per-slot tables: these copies queue behind the step that last read the slot
            self.off_dev[slot][:nb].copy_(offh, non_blocking=True)
            self.ld_dev[slot][:nb].copy_(ldh, non_blocking=True)
            self.y_dev[slot][:nb].copy_(lab, non_blocking=True)
            src = dict(src_off=self.off_dev[slot][:nb], src_ld=self.ld_dev[slot][:nb])
            plan = self.aug.plan([(self.store.C, int(v)) for v in n], self.device, static=self.aug_static[slot], x=self.store.flat, src=src)
            plan.update(src)
            yield self.store.flat, self.y_dev[slot][:nb], plan, slot
