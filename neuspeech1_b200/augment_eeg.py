"""EEG augmentation with the reference's generator names, applied on the device in ONE pass over the batch.

Reference: utils/augment_eeg.py (mask generators, shift_data), utils/reader.py:552-594 (order noise -> mask -> taylor),
:456-458 + :403-411 (shift), :496-506 (crop / zero-pad to 30 s), utils/utils.py:33-60 (noise; returns 2*signal + noise).
The random DECISIONS are drawn on the host with the same torch / numpy RNG calls, in the same order, as the reference, at
grid resolution; the (B, C, T) arithmetic -- expand the grid, scale, zero edges, shift, pad, cast, channels-last layout --
is `ns_aug_pass` (one read of x, one write of y).  Gaussian noise values come from a device Philox stream (statistical
parity only; every other augmentation is bit-exact, see tests/test_gpu_kernels.py::test_aug_pass_matches_oracle).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


def random_discrete_only_mask(signal_shape, unit=(1, 40), prob=0.5):
    """Same draw as utils/augment_eeg.py:15-26 (keep = rand >= prob on a ceil(C/uc) x ceil(T/ut) grid, repeat-interleaved)."""
    length = int(np.ceil(signal_shape[1] / unit[1]))
    channel_num = int(np.ceil(signal_shape[0] / unit[0]))
    pre = (torch.rand(channel_num, length) >= prob).to(torch.float32)
    pre = torch.repeat_interleave(pre, int(np.ceil(signal_shape[0] / channel_num)), dim=0)
    return torch.repeat_interleave(pre, int(np.ceil(signal_shape[1] / length)), dim=1)[:signal_shape[0], :signal_shape[1]]


def shift_data(eeg, shift):
    return np.pad(eeg, [[0, 0], [shift, 0]])


class RandomShapeMasker:
    """utils/augment_eeg.py:81-98 (random_type 1 block, 2 time, 3 channel).  Unlike the reference the caller's `unit` list
    is not mutated."""

    def __init__(self, unit=(1, 40), mask_prob=0.25, random_type=1):
        self.unit, self.mask_prob, self.random_type = list(unit), mask_prob, random_type

    def effective_unit(self, signal_shape):
        unit = list(self.unit)
        if self.random_type == 2:
            unit[0] = signal_shape[0]
        elif self.random_type == 3:
            unit[1] = signal_shape[1]
        elif self.random_type != 1:
            raise NotImplementedError
        return unit

    def __call__(self, signal_shape):
        return random_discrete_only_mask(signal_shape, unit=self.effective_unit(signal_shape), prob=self.mask_prob)


@dataclass
class _Plan:
    n: int
    flags: int = 0
    snr_db: Optional[np.ndarray] = None
    grid: Optional[torch.Tensor] = None
    rep_c: int = 1
    rep_t: int = 1
    e0: int = 0
    e1: int = 0
    shift: int = 0


class BatchAugmenter:
    """`augment_configs` is the dict loaded from configs/augmentation1.json-style files (keys noise / mask / taylor / shift)."""

    def __init__(self, augment_configs: Dict, max_duration: float = 30.0, sample_rate: int = 200, train: bool = True):
        self.cfg = augment_configs or {}
        self.max_length = int(max_duration * sample_rate)
        self.sample_rate = sample_rate
        self.train = train

    def _draw(self, shape) -> _Plan:
        p = _Plan(n=int(shape[1]))
        for k, v in self.cfg.items():                                   # dict order, like reader.py:553
            if k == "noise" and torch.rand(1).item() < v["prob"]:
                p.flags |= 2
                p.snr_db = np.random.uniform(v["min_snr_dB"], v["max_snr_dB"], size=shape[0])
            if k == "mask" and torch.rand(1).item() < v["prob"]:
                m = RandomShapeMasker(**v["kwargs"])
                unit = m.effective_unit(shape)
                gl = int(np.ceil(shape[1] / unit[1])); gc = int(np.ceil(shape[0] / unit[0]))
                p.grid = (torch.rand(gc, gl) >= m.mask_prob).to(torch.uint8)
                p.rep_c = int(np.ceil(shape[0] / gc)); p.rep_t = int(np.ceil(shape[1] / gl))
                p.flags |= 1
            if k == "taylor" and torch.rand(1).item() < v["prob"]:
                p.e0 = int(np.random.randint(1, 10)); p.e1 = int(np.random.randint(1, 10))
        if self.train and "shift" in self.cfg and torch.rand(1).item() < self.cfg["shift"]["prob"]:
            max_shift = int(self.max_length - p.n - 0.5 * self.sample_rate)
            p.shift = int(np.random.randint(max_shift, size=None))
        return p

    def plan(self, shapes: Sequence[Sequence[int]], device, static: Optional[dict] = None, x: Optional[torch.Tensor] = None,
             src: Optional[dict] = None, seed: int = 0) -> dict:
        """Draw the batch's random decisions -> keyword arguments of ops.aug_pass / engine.encode(aug=...).

        x (+ src = {"src_off", "src_ld"} for a ragged sample store): the batch on the device.  Needed only when the configuration
        can draw gaussian noise: its per-channel sigma is a statistic of the samples (`ns_channel_meansq`).

        static: a dict made by `static_buffers(B, grid_cap, device)`.  The decisions are then written INTO those persistent
        device tensors and the returned tensors are views of them, so that the arguments keep their addresses from step to step
        (what `engine.train_step(use_graph=True)` keys its CUDA graphs on).  Without it fresh tensors are returned."""
        plans = [self._draw(s) for s in shapes]
        B = len(plans)
        has_grid = any(p.grid is not None for p in plans)
        pin = torch.cuda.is_available() and torch.device(device).type == "cuda"
        # All per-sample integers travel in ONE pinned staging buffer and one asynchronous copy (a pageable
        # torch.tensor(..., device=...) per field is a synchronous memcpy on the compute stream: it drains the GPU every step).
        rows = [[p.n for p in plans], [p.shift for p in plans], [p.e0 for p in plans], [p.e1 for p in plans], [p.flags for p in plans]]
        if has_grid or static is not None:
            rows += [[p.grid.shape[1] if p.grid is not None else 1 for p in plans], [p.rep_c for p in plans], [p.rep_t for p in plans]]
        host = torch.empty(len(rows), B, dtype=torch.int32, pin_memory=pin)
        host.copy_(torch.tensor(rows, dtype=torch.int32))
        if static is not None:
            devt = static["ints"][:, :B]
            devt.copy_(host, non_blocking=True)
        else:
            devt = host.to(device, non_blocking=True)
        kw = dict(n=devt[0], shift=devt[1], e0=devt[2], e1=devt[3], flags=devt[4])
        if has_grid or static is not None:
            gmax = max([p.grid.numel() for p in plans if p.grid is not None] + [1])
            if static is not None:
                cap = static["grid"].shape[1]
                if gmax > cap:
                    raise ValueError(f"mask grid of {gmax} cells exceeds the static buffer ({cap}); size it with static_buffers()")
                gmax = cap
            grid = torch.ones(B, gmax, dtype=torch.uint8, pin_memory=pin)
            for b, p in enumerate(plans):
                if p.grid is not None:
                    grid[b, :p.grid.numel()] = p.grid.reshape(-1)
            if static is not None:
                gdev = static["grid"][:B]
                gdev.copy_(grid, non_blocking=True)
            else:
                gdev = grid.to(device, non_blocking=True)
            kw.update(grid=gdev, grid_stride=gmax, gl=devt[5], rep_c=devt[6], rep_t=devt[7])
        self._plans = plans
        if any(p.flags & 2 for p in plans) and torch.device(device).type == "cuda":   # (host-only planning, as in the CPU tests, stops at the decisions)
            if x is None:
                raise ValueError("a noise augmentation was drawn: plan() needs the device batch `x` to measure the channel power")
            if static is not None:
                # the Philox seed is a by-value kernel argument: a replayed CUDA graph would repeat one noise pattern for ever
                raise NotImplementedError("gaussian-noise augmentation with graph-static plans (the seed is frozen by the capture); "
                                          "plan without `static` and step with use_graph=False")
            kw.update(self._noise_kwargs(x, kw["n"], plans, device, seed, src))
        return kw

    def _noise_kwargs(self, x, n, plans, device, seed, src=None) -> dict:
        from . import ops
        B, C = len(plans), max(len(p.snr_db) for p in plans if p.snr_db is not None)
        ms = torch.empty(B, C, dtype=torch.float32, device=device)
        ops.channel_meansq(x, n, ms, **(src or {}))
        snr = torch.zeros(B, C)
        for b, p in enumerate(plans):
            if p.snr_db is not None:
                snr[b] = torch.from_numpy(p.snr_db).float()
        return dict(sigma=torch.sqrt(ms / torch.pow(10.0, snr.to(device) / 10.0)), seed=seed)

    def static_buffers(self, B: int, n_channels: int, device) -> dict:
        """Persistent device tensors for `plan(..., static=...)`, sized for the finest mask grid the configuration can draw
        (ceil(C / unit_c) x ceil(max_length / unit_t) cells; a full row / column unit for random_type 2 / 3)."""
        cap = 1
        if "mask" in self.cfg:
            m = RandomShapeMasker(**self.cfg["mask"]["kwargs"])
            uc, ut = m.unit
            gc = 1 if m.random_type == 2 else int(np.ceil(n_channels / uc))
            gt = 1 if m.random_type == 3 else int(np.ceil(self.max_length / ut))
            cap = gc * gt
        return {"ints": torch.zeros(8, B, dtype=torch.int32, device=device), "grid": torch.ones(B, cap, dtype=torch.uint8, device=device)}

    def __call__(self, samples: List[np.ndarray], out: torch.Tensor, layout: int = 1, seed: int = 0) -> torch.Tensor:
        """samples: list of (C, n_b) arrays -> `out` ((B,T,Cp) channels-last for layout 1, (B,C,T) for layout 0) on the device."""
        from . import ops
        dev = out.device
        B, C = len(samples), samples[0].shape[0]
        Tin = max(s.shape[1] for s in samples)
        host = torch.zeros(B, C, Tin, dtype=torch.float32).pin_memory()
        for b, s in enumerate(samples):
            host[b, :, :s.shape[1]] = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32))
        x = host.to(dev, non_blocking=True)
        kw = self.plan([s.shape for s in samples], dev, x=x, seed=seed)
        return ops.aug_pass(x, out, layout, **kw)
