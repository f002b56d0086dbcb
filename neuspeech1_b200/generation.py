"""Beam-search decoding loop of `evaluation.py:370-385` (`num_beams=5, repetition_penalty=5.0, no_repeat_ngram_size=2`).

The loop is model-agnostic host logic over torch tensors on whatever device the logits live on: the caller supplies
`step_fn(tokens (B*K, Lq), pos) -> logits (B*K, V)` (one decoder pass with its KV cache) and `reorder_fn(beam_idx (B*K,))`
(the cache gather of `utils/load_model.py:1353-1360::_reorder_cache`).  It restates the behaviour of the generation loop the
reference inherits from `transformers` (GenerationMixin beam search with the default `early_stopping=False`,
`length_penalty=1.0`, one EOS id, `num_return_sequences=1`):

  * scores are log-softmax of the logits; the processors act on them in HF's order: sequence bias (evaluation.py:339-343,
    `sequence_bias={(token ids): bias}`), repetition penalty (a seen token's score s becomes s*penalty if s < 0 else
    s/penalty), no-repeat-n-gram ban, begin-suppress at the first generated position
  * per sample the best 2K of the K*V continuations are kept; the K best that did NOT just stop continue, the ones among the
    first K that did stop (EOS, or the length limit) become finished hypotheses scored by sum-logprob / generated_length**lp
  * a sample is done when its best running score, normalised by the current generated length, can no longer beat its worst
    finished hypothesis (only once it has K finished ones); the loop ends when every sample is done or nothing can continue

`tests/test_generation_cpu.py` holds this loop to stock `transformers` `generate` on the CPU (identical ids in fp32, the
oracle's decoder as step function); `tests/test_gpu_model.py` then runs it over the B200 decoder step.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

NEG = -1.0e9


def apply_repetition_penalty(scores: torch.Tensor, seqs: torch.Tensor, penalty: float) -> torch.Tensor:
    """scores (N, V), seqs (N, t): every token already in a row's sequence is penalised (prompt included, like HF)."""
    if penalty == 1.0:
        return scores
    s = torch.gather(scores, 1, seqs)
    s = torch.where(s < 0, s * penalty, s / penalty)
    return scores.scatter(1, seqs, s)


def apply_no_repeat_ngram(scores: torch.Tensor, seqs: torch.Tensor, n: int) -> torch.Tensor:
    """Ban every token that would complete an n-gram already present in its row (vectorised over rows)."""
    t = seqs.shape[1]
    if n <= 0 or t + 1 < n:
        return scores
    if n == 1:
        return scores.scatter(1, seqs, float("-inf"))
    # windows of n-1 tokens starting at i (i + n - 1 < t) whose continuation seqs[:, i + n - 1] exists, compared with the last
    # n-1 tokens of the row
    tail = seqs[:, t - (n - 1):]                                   # (N, n-1)
    nwin = t - (n - 1)
    if nwin <= 0:
        return scores
    match = torch.ones(seqs.shape[0], nwin, dtype=torch.bool, device=seqs.device)
    for j in range(n - 1):
        match &= seqs[:, j:j + nwin] == tail[:, j:j + 1]
    banned = seqs[:, n - 1:n - 1 + nwin]                           # continuation of window i
    rows = match.nonzero(as_tuple=True)
    if rows[0].numel():
        scores = scores.clone()
        scores[rows[0], banned[rows[0], rows[1]]] = float("-inf")
    return scores


def apply_sequence_bias(scores: torch.Tensor, seqs: torch.Tensor, sequence_bias) -> torch.Tensor:
    """transformers SequenceBiasLogitsProcessor: `sequence_bias` maps token-id tuples to a bias; a length-1 entry biases its
    token everywhere, a longer one biases its last token in the rows whose latest tokens equal its prefix."""
    if not sequence_bias:
        return scores
    scores = scores.clone()
    t = seqs.shape[1]
    for ids, bias in sequence_bias.items():
        ids = tuple(int(i) for i in ids)
        if any(i >= scores.shape[1] for i in ids):
            raise ValueError(f"sequence_bias token out of the vocabulary: {ids}")
        if len(ids) == 1:
            scores[:, ids[0]] += float(bias)
        elif len(ids) <= t:
            pre = torch.tensor(ids[:-1], dtype=seqs.dtype, device=seqs.device)
            rows = (seqs[:, t - len(pre):] == pre).all(dim=1)
            scores[:, ids[-1]] += rows.to(scores.dtype) * float(bias)
    return scores


def torch_scorer(vocab: int, num_beams: int, begin_suppress_tokens: Sequence[int], repetition_penalty: float, no_repeat_ngram_size: int,
                 sequence_bias=None):
    """The scoring step as plain tensor ops (CPU tests, sequence_bias): (logits (B*K, >=V), seqs (B*K, t), run_score (B, K),
    first) -> (top_score, src_beam, tok), each (B, 2K)."""
    K = num_beams

    def score(logits, flat, run_score, first):
        B = run_score.shape[0]
        lp = torch.log_softmax(logits[:, :vocab].to(torch.float32), dim=-1)
        lp = apply_sequence_bias(lp, flat, sequence_bias)
        lp = apply_repetition_penalty(lp, flat, repetition_penalty)
        lp = apply_no_repeat_ngram(lp, flat, no_repeat_ngram_size)
        if first and len(begin_suppress_tokens):
            lp = lp.index_fill(1, torch.tensor(list(begin_suppress_tokens), dtype=torch.long, device=lp.device), float("-inf"))
        acc = (lp.view(B, K, vocab) + run_score[:, :, None]).view(B, K * vocab)
        top_score, top_idx = torch.topk(acc, 2 * K, dim=1)
        return top_score, top_idx // vocab, top_idx % vocab

    return score


@torch.no_grad()
def beam_search(step_fn: Callable[[torch.Tensor, int], torch.Tensor], reorder_fn: Callable[[torch.Tensor], None],
                prompt: torch.Tensor, num_beams: int, max_length: int, vocab: int, eos_token_id: int, pad_token_id: int,
                begin_suppress_tokens: Sequence[int] = (), repetition_penalty: float = 1.0, no_repeat_ngram_size: int = 0,
                length_penalty: float = 1.0, sequence_bias=None, scorer=None) -> torch.Tensor:
    """-> best hypothesis per sample, (B, <= max_length) int64 including the prompt, padded with `pad_token_id`.
    `prompt` is (B, L0); the first call of step_fn receives the prompt repeated K times per sample (rows b*K + k).
    `scorer(logits, seqs, run_score, first) -> (top_score, src_beam, tok)` is the vocabulary-sized part of a step: the B200
    engine passes its fused kernel (ns_beam_row_topk), the default is `torch_scorer`."""
    if scorer is None:
        scorer = torch_scorer(vocab, num_beams, begin_suppress_tokens, repetition_penalty, no_repeat_ngram_size, sequence_bias)
    dev = prompt.device
    B, L0 = prompt.shape
    K = num_beams
    if max_length <= L0:
        return prompt.clone()
    run_seq = torch.full((B, K, max_length), pad_token_id, dtype=torch.long, device=dev)
    run_seq[:, :, :L0] = prompt[:, None, :]
    fin_seq = run_seq.clone()
    run_score = torch.zeros(B, K, dtype=torch.float32, device=dev)
    run_score[:, 1:] = NEG                                          # all K beams start identical: only the first one counts
    fin_score = torch.full((B, K), NEG, dtype=torch.float32, device=dev)
    fin_len = torch.full((B, K), L0, dtype=torch.long, device=dev)
    is_fin = torch.zeros(B, K, dtype=torch.bool, device=dev)
    can_improve = torch.ones(B, 1, dtype=torch.bool, device=dev)
    first_k = torch.zeros(2 * K, dtype=torch.bool, device=dev)
    first_k[:K] = True
    batch_off = (torch.arange(B, device=dev) * K)[:, None]

    def gather(x, idx):                                             # x (B, n, ...), idx (B, m) -> (B, m, ...)
        while idx.dim() < x.dim():
            idx = idx.unsqueeze(-1)
        return torch.gather(x, 1, idx.expand(-1, -1, *x.shape[2:]))

    cur = L0
    while True:
        flat = run_seq[:, :, :cur].reshape(B * K, cur)
        tokens = flat if cur == L0 else flat[:, -1:]
        logits = step_fn(tokens, 0 if cur == L0 else cur - 1)
        top_score, src_beam, tok = scorer(logits, flat, run_score, cur == L0)
        cand = gather(run_seq, src_beam)
        cand[:, :, cur] = tok
        stop = (tok == eos_token_id) | (cur + 1 >= max_length)      # EOS, or the length limit reached by this token
        # ---- the K best candidates that go on
        go_score = top_score + stop.to(torch.float32) * NEG
        keep = torch.topk(go_score, K, dim=1)[1]
        run_seq = gather(cand, keep)
        run_score = gather(go_score, keep)
        beam_idx = (gather(src_beam, keep) + batch_off).reshape(-1)
        # ---- finished hypotheses: only candidates among the first K count
        just = stop & first_k[None, :]
        fs = top_score / float(cur + 1 - L0) ** length_penalty
        fs = fs + (~can_improve).to(torch.float32) * NEG + (~just).to(torch.float32) * NEG
        m_seq = torch.cat((fin_seq, cand), dim=1)
        m_score = torch.cat((fin_score, fs), dim=1)
        m_len = torch.cat((fin_len, torch.full((B, 2 * K), cur + 1, dtype=torch.long, device=dev)), dim=1)
        m_fin = torch.cat((is_fin, just), dim=1)
        best = torch.topk(m_score, K, dim=1)[1]
        fin_seq, fin_score, fin_len, is_fin = gather(m_seq, best), gather(m_score, best), gather(m_len, best), gather(m_fin, best)
        cur += 1
        # ---- can the running beams still beat the worst finished hypothesis?
        best_possible = run_score[:, :1] / float(cur - L0) ** length_penalty
        worst_fin = torch.where(is_fin, fin_score.min(dim=1, keepdim=True)[0], torch.full_like(fin_score, NEG))
        can_improve = can_improve & (best_possible > worst_fin).any(dim=-1, keepdim=True)
        if not bool(can_improve.any()) or bool(stop.all()):
            break
        reorder_fn(beam_idx)
    out_len = int(fin_len[:, 0].max())
    out = fin_seq[:, 0, :out_len].clone()
    pos = torch.arange(out_len, device=dev)[None, :]
    return torch.where(pos < fin_len[:, :1], out, torch.full_like(out, pad_token_id))
