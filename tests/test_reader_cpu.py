"""Host logic of the on-device input pipeline (neuspeech1_b200/reader.py) on the CPU device: channel selection like
utils/reader.py:270-282 / :508-516, label collation like utils/data_utils.py:198-219, double-buffered batches with stable
addresses, lengths handed to the augmentation pass."""
import numpy as np
import pytest
import torch

from neuspeech1_b200.reader import DeviceBatchLoader, collate_labels, select_channels


def test_select_channels_follows_the_reference_windows():
    s = np.arange(320 * 5, dtype=np.float32).reshape(320, 5)
    assert np.array_equal(select_channels(s, "/data/schoffelen/sub-A2002/x.npy", 273), s[28:301])
    assert np.array_equal(select_channels(s, "/data/gwilliams2023/x.npy", 208), s[:208])
    assert np.array_equal(select_channels(s, "/data/other/x.npy", 64), s[:64])
    padded = select_channels(s, "/data/gwilliams2023/x.npy", 273)            # combined-dataset run: zero rows appended
    assert padded.shape == (273, 5) and np.array_equal(padded[:208], s[:208]) and not padded[208:].any()
    with pytest.raises(ValueError):
        select_channels(s, "/data/schoffelen/x.npy", 208)                    # 273 selected rows do not fit a 208-channel stem


def test_collate_labels():
    out = collate_labels([[7, 1, 2, 3], [7, 5]], bos_token_id=7)
    assert out.tolist() == [[1, 2, 3], [5, -100, -100]]
    out = collate_labels([[7, 1, 2, 3], [8, 5]], bos_token_id=7)             # not every row starts with BOS: nothing is cut
    assert out.tolist() == [[7, 1, 2, 3], [8, 5, -100, -100]]


def test_loader_double_buffers_and_lengths():
    rng = np.random.RandomState(0)
    items = []
    for i in range(5):
        n = int(rng.randint(50, 400))
        items.append({"array": rng.randn(12, n).astype(np.float32), "path": f"/x/other/{i}.npy", "labels": list(range(3 + i))})
    cfg = {"mask": {"prob": 1.0, "kwargs": {"unit": [1, 40], "mask_prob": 0.25, "random_type": 1}}}
    ld = DeviceBatchLoader(items, batch_size=2, modal_ch=16, device="cpu", augment_configs=cfg, max_duration=2.0, sample_rate=200,
                           max_label_len=8)
    assert len(ld) == 2
    seen, ptrs = [], []
    for k, (x, y, aug, slot) in enumerate(ld):
        assert slot == k % 2 and x.shape == (2, 16, 400) and y.shape == (2, 8)
        for b in range(2):
            it = items[2 * k + b]
            n = it["array"].shape[1]
            assert int(aug["n"][b]) == n
            assert torch.equal(x[b, :12, :n], torch.from_numpy(it["array"])) and not x[b, 12:, :n].any()
            assert y[b, :len(it["labels"])].tolist() == it["labels"] and bool((y[b, len(it["labels"]):] == -100).all())
        assert aug["grid"].shape[1] == aug["grid_stride"] == 16 * 10 and int(aug["flags"][0]) == 1
        seen.append(k); ptrs.append((x.data_ptr(), y.data_ptr(), aug["n"].data_ptr(), aug["grid"].data_ptr()))
    assert seen == [0, 1] and ptrs[0] != ptrs[1]
    ptrs2 = [(x.data_ptr(), y.data_ptr(), aug["n"].data_ptr(), aug["grid"].data_ptr()) for x, y, aug, _ in ld]
    assert ptrs2 == ptrs                                    # a second epoch lands in the same buffers


def test_loader_slow_consumer_never_sees_a_later_batch():
    """Stress of the pinned-slot handshake: with a consumer that dawdles between taking a batch off the queue and copying it, the
    worker used to refill the slot with batch k+2 first (batch k then carried batch k+2's samples with batch k's labels)."""
    rng = np.random.RandomState(1)
    items = [{"array": np.full((4, 20 + i), float(i + 1), dtype=np.float32), "path": f"/x/other/{i}.npy", "labels": [i]}
             for i in range(16)]
    ld = DeviceBatchLoader(items, batch_size=2, modal_ch=4, device="cpu", augment_configs={}, max_duration=0.5, sample_rate=200,
                           max_label_len=2)
    ld._consumer_delay = 0.05
    for epoch in range(2):
        for k, (x, y, aug, slot) in enumerate(ld):
            for b in range(2):
                i = 2 * k + b
                assert float(x[b, 0, 0]) == float(i + 1) and int(y[b, 0]) == i and int(aug["n"][b]) == 20 + i, (epoch, k, b)


def test_sample_store_packs_unpadded_rows_and_loader_hands_out_tables():
    """reader.SampleStore / ResidentBatchLoader on the CPU device: every recording sits unpadded in one flat tensor (rows start on
    multiples of 8 elements), a batch is (src_off, src_ld, n) + labels in per-slot persistent tensors."""
    from neuspeech1_b200.reader import ResidentBatchLoader, SampleStore
    rng = np.random.RandomState(0)
    items = []
    for i in range(6):
        n = int(rng.randint(30, 90))
        items.append({"array": rng.randn(12, n).astype(np.float32), "path": f"/x/other/{i}.npy", "labels": list(range(2 + i))})
    st = SampleStore(items, modal_ch=16, device="cpu", dtype=torch.float32, max_duration=0.4, sample_rate=200, chunk_bytes=16 * 88 * 4 * 2)
    assert len(st) == 6 and st.flat.numel() == int((st.ld.astype(np.int64) * 16).sum())
    for i, it in enumerate(items):
        n = min(it["array"].shape[1], 80)                                   # cut to max_duration * sample_rate
        assert st.n[i] == n and st.ld[i] % 8 == 0 and st.off[i] % 8 == 0
        rows = st.flat[st.off[i]: st.off[i] + 16 * st.ld[i]].view(16, st.ld[i])
        assert torch.equal(rows[:12, :n], torch.from_numpy(it["array"][:, :n])) and not rows[12:].any() and not rows[:, n:].any()
    ld = ResidentBatchLoader(st, [it["labels"] for it in items], batch_size=2, max_label_len=8)
    assert len(ld) == 3
    ptrs = []
    for k, (x, y, aug, slot) in enumerate(ld):
        assert x.data_ptr() == st.flat.data_ptr() and slot == k % 2
        for b in range(2):
            i = 2 * k + b
            assert int(aug["src_off"][b]) == st.off[i] and int(aug["src_ld"][b]) == st.ld[i] and int(aug["n"][b]) == st.n[i]
            assert y[b, :len(items[i]["labels"])].tolist() == items[i]["labels"]
        ptrs.append((y.data_ptr(), aug["src_off"].data_ptr(), aug["n"].data_ptr()))
    assert ptrs[0] == ptrs[2] and ptrs[0] != ptrs[1]
    st16 = SampleStore(items, modal_ch=16, device="cpu", dtype=torch.bfloat16, max_duration=0.4, sample_rate=200)
    assert st16.bytes * 2 == st.bytes and torch.equal(st16.flat.float(), st.flat.to(torch.bfloat16).float())
