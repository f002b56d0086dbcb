"""Per-tensor bf16 gradient error of the CUDA path against the fp32 CPU oracle, next to the error of the SAME oracle code run
by torch on the GPU with bf16 storage / fp32 accumulation (the floor the number format itself sets).  Developer tool (imports
oracle/): python tools/grad_err_report.py [TINY MID BASE] -> gpurun_out/grad_err_<name>.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import whisper_eeg as O
from neuspeech1_b200.engine import ModelDims, WhisperEEGEngine

DEV = torch.device("cuda")
MID = O.Dims(d_model=256, enc_layers=2, dec_layers=2, enc_heads=4, dec_heads=4, enc_ffn=512, dec_ffn=512, vocab=2000,
             max_source_positions=160, max_target_positions=48, eeg_ch=24, pad_token_id=1997, eos_token_id=1997,
             decoder_start_token_id=1998, begin_suppress_tokens=(220, 1996), lora_r=32, lora_alpha=64)


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def torch_bf16_grads(x, labels, P, dims, lora):
    """The oracle's own forward, on the GPU, every tensor stored in bf16 (matmuls accumulate in fp32 inside cuBLAS)."""
    Pb = {k: v.to(DEV, torch.bfloat16) for k, v in P.items()}
    Lb = {k: (v.to(DEV, torch.bfloat16) if torch.is_tensor(v) else v) for k, v in lora.items()}
    loss, g, enc = O.grads(x.to(DEV, torch.bfloat16), labels.to(DEV), Pb, dims, Lb)
    return loss.float().cpu(), {k: v.float().cpu() for k, v in g.items()}, enc.float().cpu()


def main():
    names = sys.argv[1:] or ["TINY", "MID", "BASE"]
    os.makedirs("gpurun_out", exist_ok=True)
    for nm in names:
        dims = {"TINY": O.TINY, "MID": MID, "BASE": O.WHISPER_BASE}[nm]
        B, L = (2, 32) if nm == "BASE" else (3, 8)
        P = O.init_params(dims, seed=0)
        lora = O.init_lora(dims, seed=1, b_std=0.05)
        x, labels = O.synthetic_batch(dims, B=B, L=L, seed=1)
        loss_ref, g_ref, enc_ref = O.grads(x, labels, P, dims, lora)
        eng = WhisperEEGEngine(ModelDims.from_any(dims), P, lora, dtype=torch.bfloat16, device=DEV)
        loss, _, enc = eng.forward_loss(x.to(DEV), labels.to(DEV))
        eng.backward()
        loss_t, g_t, enc_t = torch_bf16_grads(x, labels, P, dims, lora)
        rows = {k: (rel(eng.trainable_grad(k), g_ref[k]), rel(g_t[k], g_ref[k])) for k in g_ref}
        out = {"shape": nm, "B": B, "loss": [float(loss), float(loss_t), float(loss_ref)],
               "enc_rel": [rel(enc, enc_ref), rel(enc_t, enc_ref)], "grads": rows}
        worst = sorted(rows.items(), key=lambda kv: -kv[1][0])[:8]
        print(nm, "loss ours/torch-bf16/fp32", out["loss"], "enc rel ours/torch", out["enc_rel"])
        print("  max ours %.4f  max torch-bf16 %.4f  median ours %.4f  median torch %.4f" % (
            max(v[0] for v in rows.values()), max(v[1] for v in rows.values()),
            sorted(v[0] for v in rows.values())[len(rows) // 2], sorted(v[1] for v in rows.values())[len(rows) // 2]))
        for k, v in worst:
            print("   %-70s ours %.4f torch-bf16 %.4f" % (k, v[0], v[1]))
        json.dump(out, open(f"gpurun_out/grad_err_{nm}.json", "w"), indent=1)
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
