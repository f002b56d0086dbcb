"""Minimal fused training loop for the hot path: what `Seq2SeqTrainer.train()` does for this model (finetune.py:231-282) --
batches -> augmentation pass -> forward/loss/backward -> gradient all-reduce -> clip + AdamW with linear warm-up/decay ->
periodic evaluation loss, best-eval adapter checkpoints (utils/callback.py:11-22).  The model also works under the stock HF
Trainer through its autograd path; this loop is the zero-overhead variant bench.py measures."""
from __future__ import annotations

import os
from typing import Callable, Iterable, Optional

import torch

from .lora import save_adapter
from .parallel import DataParallel, linear_warmup_decay


class Trainer:
    def __init__(self, model, lr: float = 1e-3, warmup_steps: int = 500, total_steps: Optional[int] = None, max_grad_norm: float = 1.0,
                 weight_decay: float = 0.0, output_dir: Optional[str] = None, eval_steps: int = 1000, logging_steps: int = 100,
                 dp: Optional[DataParallel] = None, log: Callable[[str], None] = print):
        """total_steps: length of the linear schedule (HF: num_train_epochs * steps per epoch).  `fit` derives it from the loader
        when it is None; `training_step` on its own needs it.  Augmentation decisions travel with the batches (`batch["aug"]`,
        produced by neuspeech1_b200.reader.DeviceBatchLoader / augment_eeg.BatchAugmenter), not with the trainer."""
        self.model, self.lr, self.warmup_steps, self.total_steps = model, lr, warmup_steps, total_steps
        self.max_grad_norm, self.weight_decay = max_grad_norm, weight_decay
        self.output_dir, self.eval_steps, self.logging_steps = output_dir, eval_steps, logging_steps
        self.dp = dp or DataParallel(device=model.device)
        self.log = log if self.dp.rank == 0 else (lambda s: None)
        self.step = 0
        self.best_eval = float("inf")

    def training_step(self, input_features: torch.Tensor, labels: torch.Tensor, aug: Optional[dict] = None) -> torch.Tensor:
        eng = self.model.engine
        if self.total_steps is None:
            raise ValueError("Trainer.training_step needs total_steps (the length of the linear schedule); fit() derives it")
        # HF order: optimizer.step() for step k runs with lambda(k) and lr_scheduler.step() comes after it, so the first step
        # of a warm-up schedule has lr = 0 (trainer.py:1934-1936, get_linear_schedule_with_warmup)
        lr = linear_warmup_decay(self.step, self.lr, self.warmup_steps, self.total_steps)
        eng.training = True
        eng._advance_seed()
        eng.pack_trainable()
        loss, _, _ = eng.forward_loss(input_features, labels, aug=aug, save=True, ce_grad_scale=1.0)
        eng.backward()
        if self.dp.world > 1:
            self.dp.all_reduce_mean(eng.grad)
        eng.optimizer_step(lr, max_grad_norm=self.max_grad_norm, weight_decay=self.weight_decay)
        self.step += 1
        return loss

    @torch.no_grad()
    def evaluate(self, loader: Iterable) -> float:
        tot, n = 0.0, 0
        was_training = self.model.training
        self.model.eval()                       # LoRA-branch dropout off (HF evaluation_loop calls model.eval())
        try:
            for batch in loader:
                out = self.model(input_features=batch["input_features"], labels=batch["labels"])
                tot += float(out.loss); n += 1
        finally:
            self.model.train(was_training)
        return tot / max(n, 1)

    def fit(self, train_loader: Iterable, epochs: int = 1, eval_loader: Optional[Iterable] = None):
        if self.total_steps is None:
            self.total_steps = epochs * len(train_loader)
        for ep in range(epochs):
            for batch in train_loader:
                loss = self.training_step(batch["input_features"], batch["labels"], aug=batch.get("aug"))
                if self.step % self.logging_steps == 0:
                    self.log(f"step {self.step} loss {float(loss):.4f}")
                if eval_loader is not None and self.step % self.eval_steps == 0:
                    ev = self.evaluate(eval_loader)
                    self.log(f"step {self.step} eval_loss {ev:.4f}")
                    if ev <= self.best_eval and self.output_dir and self.dp.rank == 0:      # SavePeftModelCallback policy
                        self.best_eval = ev
                        save_adapter(self.model, os.path.join(self.output_dir, f"checkpoint-{self.step}"))
        if self.output_dir and self.dp.rank == 0:
            save_adapter(self.model, os.path.join(self.output_dir, "checkpoint-final"))
        return self
