"""The beam-search loop (neuspeech1_b200/generation.py) against stock `transformers` generate on the CPU.

`evaluation.py:370-385` decodes with num_beams=5, repetition_penalty=5.0, no_repeat_ngram_size=2.  Here the loop runs over the
ORACLE's decoder as step function (fp32, KV cache gathered per step like `_reorder_cache`) and must return the token ids stock
HF returns for the same weights and inputs: random-init tiny Whisper + EEG stem, small and large vocabularies (EOS reached
early / never), beam widths 3 and 5, with and without a decoder prompt, two penalty settings.  HF's Whisper wrapper returns
the generated suffix and drops a final EOS before padding; pad == eos in this model family, so rows are compared after
stripping trailing pad tokens."""
import warnings

import pytest
import torch

from neuspeech1_b200.generation import apply_no_repeat_ngram, apply_repetition_penalty, beam_search
from oracle import whisper_eeg as O
from oracle.hf_bridge import build_hf


def _strip(row, pad):
    r = row.tolist()
    while r and r[-1] == pad:
        r.pop()
    return r


def _oracle_step_fns(dims, P, enc, K):
    state = {}

    def step_fn(tokens, pos):
        if pos == 0:
            state["enc"] = enc.repeat_interleave(K, 0)
            state["past"] = None
        y, state["past"] = O.decoder(tokens, state["enc"], P, dims, state["past"])
        return y[:, -1] @ P["model.decoder.embed_tokens.weight"].t()

    def reorder_fn(idx):
        state["past"] = [tuple(t.index_select(0, idx) for t in layer) if isinstance(layer, (tuple, list)) else layer.index_select(0, idx)
                         for layer in state["past"]]

    return step_fn, reorder_fn


@pytest.mark.parametrize("seed", [0, 1, 3, 5, 6, 9])
def test_beam_search_matches_transformers(seed):
    import transformers
    transformers.logging.set_verbosity_error()
    V = 40 if seed % 2 else 300                      # small vocabulary: EOS is reached early and hypotheses finish
    dims = O.Dims(d_model=64, enc_layers=1, dec_layers=2, enc_heads=2, dec_heads=2, enc_ffn=128, dec_ffn=128, vocab=V,
                  max_source_positions=40, max_target_positions=24, eeg_ch=6, pad_token_id=V - 3, eos_token_id=V - 3,
                  decoder_start_token_id=V - 2, begin_suppress_tokens=(20, V - 4), lora_r=4, lora_alpha=8)
    P = O.init_params(dims, seed=seed, std=0.5)
    g = torch.Generator().manual_seed(seed)
    B = 4
    x = torch.randn(B, dims.eeg_ch, dims.T, generator=g) * 2
    K = 5 if seed % 3 else 3
    pen, ngram = (5.0, 2) if seed % 4 else (1.3, 3)
    L0 = 1 if seed % 2 == 0 else 3
    max_length = 20
    prompt = torch.full((B, 1), dims.decoder_start_token_id, dtype=torch.long)
    if L0 > 1:
        prompt = torch.cat([prompt, torch.randint(0, V - 5, (B, L0 - 1), generator=g)], dim=1)
    m = build_hf(dims, P)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kw = dict(decoder_input_ids=prompt) if L0 > 1 else {}
        ref = m.generate(x, do_sample=False, num_beams=K, repetition_penalty=pen, no_repeat_ngram_size=ngram,
                         max_length=max_length, **kw)
    enc = O.encoder(x, P, dims, None)
    step_fn, reorder_fn = _oracle_step_fns(dims, P, enc, K)
    out = beam_search(step_fn, reorder_fn, prompt, K, max_length, dims.vocab, dims.eos_token_id, dims.pad_token_id,
                      dims.begin_suppress_tokens, pen, ngram)
    assert torch.equal(out[:, :L0], prompt)
    for b in range(B):
        assert _strip(out[b, L0:], dims.pad_token_id) == _strip(ref[b], dims.pad_token_id), (b, out[b], ref[b])
    assert len({tuple(r.tolist()) for r in ref}) > 1           # the inputs matter: not one sequence for every sample


@pytest.mark.parametrize("case", [(2.0, 4, 18), (0.5, 5, 18), (1.0, 2, 12), (0.0, 5, 16), (1.0, 5, 6), (1.5, 3, 24)])
def test_beam_search_length_penalties_match_transformers(case):
    """length_penalty (finished-hypothesis score = sum-logprob / generated_length ** lp, also in the stop heuristic), narrow and
    wide beams, a length limit a few tokens after the prompt."""
    import transformers
    transformers.logging.set_verbosity_error()
    lp, K, max_length = case
    V = 40
    dims = O.Dims(d_model=64, enc_layers=1, dec_layers=2, enc_heads=2, dec_heads=2, enc_ffn=128, dec_ffn=128, vocab=V,
                  max_source_positions=40, max_target_positions=24, eeg_ch=6, pad_token_id=V - 3, eos_token_id=V - 3,
                  decoder_start_token_id=V - 2, begin_suppress_tokens=(20, V - 4), lora_r=4, lora_alpha=8)
    seed = int(lp * 10) + K
    P = O.init_params(dims, seed=seed + 20, std=0.5)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(4, dims.eeg_ch, dims.T, generator=g) * 2
    prompt = torch.full((4, 1), dims.decoder_start_token_id, dtype=torch.long)
    m = build_hf(dims, P)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = m.generate(x, do_sample=False, num_beams=K, repetition_penalty=5.0, no_repeat_ngram_size=2, max_length=max_length,
                         length_penalty=lp)
    step_fn, reorder_fn = _oracle_step_fns(dims, P, O.encoder(x, P, dims, None), K)
    out = beam_search(step_fn, reorder_fn, prompt, K, max_length, dims.vocab, dims.eos_token_id, dims.pad_token_id,
                      dims.begin_suppress_tokens, 5.0, 2, length_penalty=lp)
    for b in range(4):
        assert _strip(out[b, 1:], dims.pad_token_id) == _strip(ref[b], dims.pad_token_id), (b, out[b], ref[b])


def test_logit_processors():
    scores = torch.log_softmax(torch.arange(12, dtype=torch.float32).view(2, 6), dim=-1)
    seqs = torch.tensor([[1, 2, 1], [0, 0, 3]])
    out = apply_repetition_penalty(scores, seqs, 2.0)
    assert torch.allclose(out[0, 1], scores[0, 1] * 2) and torch.allclose(out[0, 2], scores[0, 2] * 2) and out[0, 3] == scores[0, 3]
    assert torch.allclose(out[1, 0], scores[1, 0] * 2) and torch.allclose(out[1, 3], scores[1, 3] * 2)
    pos = torch.tensor([[0.5, -0.5]])
    assert torch.allclose(apply_repetition_penalty(pos, torch.tensor([[0, 1]]), 2.0), torch.tensor([[0.25, -1.0]]))
    # bigram ban: row 0 ends with 1, "1 2" was seen -> 2 is banned; row 1 ends with 3, never seen before -> nothing banned
    out = apply_no_repeat_ngram(scores, seqs, 2)
    assert out[0, 2] == float("-inf") and torch.isfinite(out[0, [0, 1, 3, 4, 5]]).all() and torch.isfinite(out[1]).all()
    # trigram ban needs the last two tokens to match a window
    seqs3 = torch.tensor([[4, 5, 0, 4, 5], [1, 2, 3, 4, 5]])
    out = apply_no_repeat_ngram(scores, seqs3, 3)
    assert out[0, 0] == float("-inf") and torch.isfinite(out[1]).all()
    assert torch.equal(apply_no_repeat_ngram(scores, seqs[:, :1], 3), scores)      # shorter than an n-gram: untouched


@pytest.mark.parametrize("seed", [2, 7])
def test_beam_search_sequence_bias_matches_transformers(seed):
    """evaluation.py:339-343,380: `generate(..., sequence_bias={(token ids): bias})` (GetSequenceBias, bias -1.0 on phrases of the
    training set).  Single-token and multi-token entries, applied before the repetition penalty like HF's processor list."""
    import transformers
    transformers.logging.set_verbosity_error()
    V = 60
    dims = O.Dims(d_model=64, enc_layers=1, dec_layers=2, enc_heads=2, dec_heads=2, enc_ffn=128, dec_ffn=128, vocab=V,
                  max_source_positions=40, max_target_positions=24, eeg_ch=6, pad_token_id=V - 3, eos_token_id=V - 3,
                  decoder_start_token_id=V - 2, begin_suppress_tokens=(20, V - 4), lora_r=4, lora_alpha=8)
    P = O.init_params(dims, seed=seed, std=0.5)
    g = torch.Generator().manual_seed(seed)
    B, K, max_length = 4, 5, 16
    x = torch.randn(B, dims.eeg_ch, dims.T, generator=g) * 2
    prompt = torch.full((B, 1), dims.decoder_start_token_id, dtype=torch.long)
    m = build_hf(dims, P)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plain = m.generate(x, do_sample=False, num_beams=K, repetition_penalty=5.0, no_repeat_ngram_size=2, max_length=max_length)
    # bias against what the unbiased search produced: its first token everywhere, and its first bigram / trigram of sample 0
    bias = {(int(plain[0, 0]),): -3.0, (int(plain[1, 0]), int(plain[1, 1])): -4.0, (int(plain[2, 0]), int(plain[2, 1]), int(plain[2, 2])): -4.0,
            (7,): 1.5}
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = m.generate(x, do_sample=False, num_beams=K, repetition_penalty=5.0, no_repeat_ngram_size=2, max_length=max_length,
                         sequence_bias=bias)
    assert not torch.equal(ref[:, :plain.shape[1]] if ref.shape[1] >= plain.shape[1] else ref, plain[:, :ref.shape[1]])   # the bias matters
    step_fn, reorder_fn = _oracle_step_fns(dims, P, O.encoder(x, P, dims, None), K)
    out = beam_search(step_fn, reorder_fn, prompt, K, max_length, dims.vocab, dims.eos_token_id, dims.pad_token_id,
                      dims.begin_suppress_tokens, 5.0, 2, sequence_bias=bias)
    for b in range(B):
        assert _strip(out[b, 1:], dims.pad_token_id) == _strip(ref[b], dims.pad_token_id), (b, out[b], ref[b])
