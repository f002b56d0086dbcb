"""The drop-in nn.Module (neuspeech1_b200.load_model) on the GPU: autograd path (what HF Trainer drives), external torch
optimizer on the aliased parameters, generate, stem swap, no-grad evaluation."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import whisper_eeg as O

if torch.cuda.is_available():
    from neuspeech1_b200.engine import ModelDims
    from neuspeech1_b200.load_model import WhisperForConditionalGeneration
    from neuspeech1_b200.model_utils import projection_module
    DEV = torch.device("cuda")


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_autograd_path_matches_oracle_and_external_optimizer_works():
    dims = O.TINY
    P = O.init_params(dims, seed=0); lora = O.init_lora(dims, seed=1, b_std=0.05)
    x, labels = O.synthetic_batch(dims, B=2, L=8, seed=1)
    loss_ref, grads_ref, enc_ref = O.grads(x, labels, P, dims, lora)
    m = WhisperForConditionalGeneration(ModelDims.from_any(dims), P, lora, dtype=torch.float32, device=DEV)
    m.train()
    out = m(input_features=x.to(DEV), labels=labels.to(DEV))
    assert abs(float(out.loss) - float(loss_ref)) < 1e-3 * float(loss_ref)
    assert float(out["loss"]) == float(out.loss)
    assert out.logits.shape == (2, 8, dims.vocab)
    (out.loss * 0.5).backward()
    named = dict(m.named_parameters())
    for k, g in grads_ref.items():
        p = named[k]
        assert p.grad is not None, k
        assert rel(p.grad, 0.5 * g) < 2e-3, k
    frozen = [k for k, p in named.items() if not p.requires_grad]
    assert all(named[k].grad is None for k in frozen)
    # external optimizer on the aliased parameters changes what the engine computes
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-2, weight_decay=0.0)
    opt.step(); opt.zero_grad(set_to_none=True)
    out2 = m(input_features=x.to(DEV), labels=labels.to(DEV))
    assert float(out2.loss) < float(out.loss)
    # oracle takes the same (unclipped) AdamW step
    st = O.AdamWState()
    both = {**lora, **{k: P[k] for k in grads_ref if k in P}}
    O.clip_and_adamw(both, {k: 0.5 * g for k, g in grads_ref.items()}, st, lr=1e-2, max_norm=1e9)
    loss_ref2, _, _ = O.grads(x, labels, P, dims, lora)
    assert abs(float(out2.loss) - float(loss_ref2)) < 2e-3 * float(loss_ref2)


def test_eval_generate_and_stem_swap():
    dims = O.TINY
    P = O.init_params(dims, seed=0)
    x, labels = O.synthetic_batch(dims, B=2, L=8, seed=1)
    m = WhisperForConditionalGeneration(ModelDims.from_any(dims), P, None, dtype=torch.float32, device=DEV)
    m.eval()
    with torch.no_grad():
        out = m(input_features=x.to(DEV), decoder_input_ids=O.shift_tokens_right(labels, dims.pad_token_id, dims.decoder_start_token_id).to(DEV))
    _, logits_ref, _ = O.forward_loss(x, labels, P, dims)
    assert out.loss is None and rel(out.logits, logits_ref) < 1e-3                      # evaluation.py:392-395 path
    ids = m.generate(x.to(DEV), do_sample=False, num_beams=1, max_new_tokens=10)
    ids_ref = O.greedy_decode(x, P, dims, max_length=11)
    assert torch.equal(ids.cpu(), ids_ref)
    enc = m.get_encoder()(x.to(DEV)).last_hidden_state
    assert rel(enc, O.encoder(x, P, dims)) < 1e-3
    # cross-dataset stem swap (finetune.py:150-163): new channel count, transformer weights kept
    new_ch = 21
    stem = projection_module("base", meg_ch=new_ch, d_model=dims.d_model)
    m.model.encoder.set_input_embeddings(stem)
    dims2 = O.Dims(**{**dims.__dict__, "eeg_ch": new_ch})
    P2 = dict(P)
    for k, v in stem.state_dict().items():
        P2["model.encoder.conv1." + k] = v.detach().clone()
    x2, _ = O.synthetic_batch(dims2, B=2, L=8, seed=2)
    out = m(input_features=x2.to(DEV), labels=labels.to(DEV))
    loss_ref, _, enc_ref = O.forward_loss(x2, labels, P2, dims2)
    assert rel(out.encoder_last_hidden_state, enc_ref) < 1e-3
    assert abs(float(out.loss) - float(loss_ref)) < 1e-3 * float(loss_ref)


def test_fused_training_step_reduces_loss_bf16():
    dims = O.TINY
    P = O.init_params(dims, seed=0); lora = O.init_lora(dims, seed=1, b_std=0.0)
    x, labels = O.synthetic_batch(dims, B=4, L=8, seed=3)
    m = WhisperForConditionalGeneration(ModelDims.from_any(dims), P, lora, dtype=torch.bfloat16, device=DEV)
    losses = [float(m.training_step(x.to(DEV), labels.to(DEV), lr=2e-3).loss) for _ in range(8)]
    assert losses[-1] < losses[0] - 0.05, losses


def test_trainer_fit_follows_hf_schedule_and_saves_best_adapter(tmp_path):
    """neuspeech1_b200.trainer.Trainer (finetune.py:231-282 without HF): linear warm-up/decay in HF's order (the first optimizer
    step of a warm-up runs with lr = 0), clip + AdamW, periodic evaluation with the LoRA dropout off, best-eval adapter
    checkpoints in PEFT's layout (utils/callback.py:11-22).  Weights after 6 steps == the oracle stepping with the same
    learning rates and dropout masks."""
    import os
    from neuspeech1_b200.parallel import linear_warmup_decay
    from neuspeech1_b200.trainer import Trainer
    dims = O.TINY
    P = O.init_params(dims, seed=0); lora = O.init_lora(dims, seed=1, b_std=0.05)
    m = WhisperForConditionalGeneration(ModelDims.from_any(dims), P, lora, dtype=torch.float32, device=DEV, lora_dropout=0.05)
    m.engine.set_dropout_seed(31)
    batches = []
    for k in range(3):
        x, labels = O.synthetic_batch(dims, B=2, L=6, seed=20 + k)
        batches.append({"input_features": x.to(DEV), "labels": labels.to(DEV)})
    logs = []
    tr = Trainer(m, lr=1e-3, warmup_steps=2, output_dir=str(tmp_path), eval_steps=3, logging_steps=1, log=logs.append)
    tr.fit(batches, epochs=2, eval_loader=batches[:1])
    assert tr.step == 6 and tr.total_steps == 6
    assert [linear_warmup_decay(k, 1e-3, 2, 6) for k in range(3)] == [0.0, 5e-4, 1e-3]
    Pc = {k: v.clone() for k, v in P.items()}; lc = {k: v.clone() for k, v in lora.items()}
    st = O.AdamWState()
    seed = 31
    for k in range(6):
        b = batches[k % 3]
        seed = O.next_dropout_seed(seed)
        lc["__dropout__"] = (0.05, seed)
        O.train_step(b["input_features"].cpu(), b["labels"].cpu(), Pc, dims, lc, st, lr=linear_warmup_decay(k, 1e-3, 2, 6))
    for name in O.trainable_names(Pc, lc):
        ref = lc[name] if name in lc else Pc[name]
        assert rel(m.engine.trainable(name), ref) < 2e-3, name
    lc.pop("__dropout__")
    with torch.no_grad():
        ev_ref, _, _ = O.forward_loss(batches[0]["input_features"].cpu(), batches[0]["labels"].cpu(), Pc, dims, lc)
    assert abs(tr.evaluate(batches[:1]) - float(ev_ref)) < 1e-3 * float(ev_ref)          # evaluation runs without dropout
    assert m.training                                                                    # ... and restores train mode
    assert any("eval_loss" in l for l in logs)
    for ck in ("checkpoint-final",):
        files = set(os.listdir(tmp_path / ck))
        assert {"adapter_model.safetensors", "adapter_config.json"} <= files
    assert any(d.startswith("checkpoint-3") or d.startswith("checkpoint-6") for d in os.listdir(tmp_path))


def test_generate_contract_beams_bias_and_early_stop():
    """model.generate as evaluation.py:370-385 calls it: beam 5 + penalties (+ sequence_bias, :339-343) return what the tensor-op
    loop over the oracle decoder returns (that loop is pinned to transformers on the CPU); greedy stops early once every row
    has emitted EOS; `_reorder_cache` is there for external loops (utils/load_model.py:1353-1360)."""
    from neuspeech1_b200.generation import beam_search
    dims = O.Dims(d_model=256, enc_layers=2, dec_layers=2, enc_heads=4, dec_heads=4, enc_ffn=512, dec_ffn=512, vocab=120,
                  max_source_positions=160, max_target_positions=32, eeg_ch=24, pad_token_id=117, eos_token_id=117,
                  decoder_start_token_id=118, begin_suppress_tokens=(20, 116), lora_r=32, lora_alpha=64)
    P = O.init_params(dims, seed=0, std=0.12)
    x, _ = O.synthetic_batch(dims, B=3, L=8, seed=1)
    m = WhisperForConditionalGeneration(ModelDims.from_any(dims), P, None, dtype=torch.float32, device=DEV)
    m.eval()
    enc = O.encoder(x, P, dims, None)
    K = 5

    def ref_ids(bias):
        state = {}

        def step_fn(tokens, pos):
            if pos == 0:
                state["enc"] = enc.repeat_interleave(K, 0); state["past"] = None
            y, state["past"] = O.decoder(tokens, state["enc"], P, dims, state["past"])
            return y[:, -1] @ P["model.decoder.embed_tokens.weight"].t()

        def reorder_fn(idx):
            state["past"] = [[t.index_select(0, idx) for t in layer] for layer in state["past"]]

        prompt = torch.full((3, 1), dims.decoder_start_token_id, dtype=torch.long)
        return beam_search(step_fn, reorder_fn, prompt, K, 20, dims.vocab, dims.eos_token_id, dims.pad_token_id,
                           dims.begin_suppress_tokens, 5.0, 2, sequence_bias=bias)[:, 1:]

    plain = m.generate(x.to(DEV), num_beams=K, repetition_penalty=5.0, no_repeat_ngram_size=2, max_length=20).cpu()
    assert torch.equal(plain, ref_ids(None))
    bias = {(int(plain[0, 0]),): -5.0, (int(plain[1, 0]), int(plain[1, 1])): -5.0}
    biased = m.generate(x.to(DEV), num_beams=K, repetition_penalty=5.0, no_repeat_ngram_size=2, max_length=20, sequence_bias=bias).cpu()
    assert torch.equal(biased, ref_ids(bias)) and not torch.equal(biased[:, :plain.shape[1]], plain[:, :biased.shape[1]])
    # greedy: with EOS made overwhelmingly likely after the first token every row finishes at once and the loop stops early
    g1 = m.generate(x.to(DEV), max_length=dims.max_target_positions).cpu()
    assert g1.shape == (3, dims.max_target_positions - 1)
    ref = O.greedy_decode(x, P, dims, max_length=dims.max_target_positions)
    assert torch.equal(g1[:, :ref.shape[1]], ref) and bool((g1[:, ref.shape[1]:] == dims.pad_token_id).all())
    past = ((torch.arange(6.).view(3, 2), torch.ones(3, 2)),)
    out = m._reorder_cache(past, torch.tensor([2, 0, 0]))
    assert torch.equal(out[0][0], past[0][0][[2, 0, 0]])
