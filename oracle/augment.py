"""CPU oracle for the EEG augmentation pass  --  TEST INFRASTRUCTURE ONLY (see oracle/whisper_eeg.py header).

Restates, with the same RNG call sequence so results are bit-identical under the same seeds:
  * random_discrete_only_mask / RandomShapeMasker   utils/augment_eeg.py:15-26, :81-98
  * shift_data                                      utils/augment_eeg.py:54-56
  * add_gaussian_noise (returns 2*signal + noise!)  utils/utils.py:33-60
  * augment_audio order noise -> mask -> taylor     utils/reader.py:552-594
  * shift (after augment, in __getitem__)           utils/reader.py:456-458, :403-411
  * padding_sample crop / zero-pad to 6000          utils/reader.py:496-506
Pinned by tests/golden/augment_*.npz (made by oracle/make_golden.py running the reference's own utils/augment_eeg.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def grid_shape(signal_shape: Sequence[int], unit: Sequence[int]) -> Tuple[int, int, int, int]:
    """(grid_channels, grid_length, repeat_c, repeat_t) of random_discrete_only_mask."""
    gl = int(np.ceil(signal_shape[1] / unit[1]))
    gc = int(np.ceil(signal_shape[0] / unit[0]))
    return gc, gl, int(np.ceil(signal_shape[0] / gc)), int(np.ceil(signal_shape[1] / gl))


def random_grid(signal_shape, unit=(1, 40), prob=0.5) -> torch.Tensor:
    """The grid-resolution keep-mask (1 = keep): torch.rand(grid) >= prob.  Same RNG draw as the reference."""
    gc, gl, _, _ = grid_shape(signal_shape, unit)
    pre = torch.rand(gc, gl)
    return (pre >= prob).to(torch.float32)


def expand_grid(grid: torch.Tensor, signal_shape, unit) -> torch.Tensor:
    _, _, rc, rt = grid_shape(signal_shape, unit)
    m = torch.repeat_interleave(grid, rc, dim=0)
    return torch.repeat_interleave(m, rt, dim=1)[: signal_shape[0], : signal_shape[1]]


def random_discrete_only_mask(signal_shape, unit=(1, 40), prob=0.5) -> torch.Tensor:
    return expand_grid(random_grid(signal_shape, unit, prob), signal_shape, unit)


def effective_unit(signal_shape, unit, random_type: int) -> List[int]:
    unit = list(unit)
    if random_type == 2:      # time masking: one grid row spans all channels
        unit[0] = signal_shape[0]
    elif random_type == 3:    # channel masking
        unit[1] = signal_shape[1]
    elif random_type != 1:
        raise NotImplementedError
    return unit


def shape_mask(signal_shape, unit=(1, 40), mask_prob=0.25, random_type=1) -> torch.Tensor:
    return random_discrete_only_mask(signal_shape, effective_unit(signal_shape, unit, random_type), mask_prob)


def shift_data(eeg: np.ndarray, shift: int) -> np.ndarray:
    return np.pad(eeg, [[0, 0], [shift, 0]])


def add_gaussian_noise(signal: np.ndarray, snr_range) -> np.ndarray:
    ch, length = signal.shape
    snr = np.random.uniform(*snr_range, size=ch)
    noisy = np.zeros_like(signal)
    for i in range(ch):
        std = np.sqrt(np.mean(signal[i] ** 2) / (10 ** (snr[i] / 10)))
        noisy[i] = signal[i] + np.random.normal(scale=std, size=length)
    return signal + noisy


def padding_sample(sample: np.ndarray, max_length: int = 6000) -> np.ndarray:
    sample = sample[:, :max_length]
    return np.pad(sample, ((0, 0), (0, max_length - sample.shape[-1])))


@dataclass
class SamplePlan:
    """Everything random about one sample's augmentation, drawn on the host in the reference's order."""
    n: int                      # original length
    noise: bool = False
    snr_db: Optional[np.ndarray] = None
    grid: Optional[torch.Tensor] = None   # keep-grid (gc, gl) or None
    rep_c: int = 1
    rep_t: int = 1
    edge0: int = 0              # taylor: zero [0, edge0) and [n-edge1, n)
    edge1: int = 0
    shift: int = 0


def draw_plan(shape, cfg: dict, max_length: int = 6000, sample_rate: int = 200, train: bool = True) -> SamplePlan:
    """Draw one sample's random decisions with the reference's RNG call order (reader.py:552-594 then :456-458).
    Gaussian noise values themselves are not drawn here (they can only be matched statistically on a device)."""
    plan = SamplePlan(n=int(shape[1]))
    for k, v in cfg.items():
        if k == "noise" and torch.rand(1).item() < v["prob"]:
            plan.noise = True
            plan.snr_db = np.random.uniform(v["min_snr_dB"], v["max_snr_dB"], size=shape[0])
        if k == "mask" and torch.rand(1).item() < v["prob"]:
            kw = v["kwargs"]
            unit = effective_unit(shape, kw.get("unit", (1, 40)), kw.get("random_type", 1))
            plan.grid = random_grid(shape, unit, kw.get("mask_prob", 0.25))
            _, _, plan.rep_c, plan.rep_t = grid_shape(shape, unit)
        if k == "taylor" and torch.rand(1).item() < v["prob"]:
            plan.edge0 = int(np.random.randint(1, 10)); plan.edge1 = int(np.random.randint(1, 10))
    if train and "shift" in cfg and torch.rand(1).item() < cfg["shift"]["prob"]:
        max_shift = int(max_length - plan.n - 0.5 * sample_rate)
        plan.shift = int(np.random.randint(max_shift, size=None))
    return plan


def apply_plan(sample: np.ndarray, plan: SamplePlan, max_length: int = 6000, noise: Optional[np.ndarray] = None) -> np.ndarray:
    """Deterministic part: (C,n) -> (C,max_length) f32.  `noise` (C,n) unit-normal draws, if plan.noise."""
    x = sample.astype(np.float32)
    if plan.noise:
        std = np.sqrt(np.mean(x.astype(np.float64) ** 2, axis=1) / (10 ** (plan.snr_db / 10)))
        x = (2.0 * x + (noise * std[:, None])).astype(np.float32)
    if plan.grid is not None:
        m = torch.repeat_interleave(torch.repeat_interleave(plan.grid, plan.rep_c, 0), plan.rep_t, 1)
        x = x * m[: x.shape[0], : x.shape[1]].numpy()
    if plan.edge0 or plan.edge1:
        x = x.copy(); x[:, : plan.edge0] = 0
        if plan.edge1:
            x[:, -plan.edge1:] = 0
    if plan.shift:
        x = shift_data(x, plan.shift)
    return padding_sample(x, max_length).astype(np.float32)
