"""The loss call of segmentor/trainer_contrastive.py:209-252, as a hook a trainer can call.

Only the hook point is replaced: warm-up gate, bank attach, loss, enqueue AFTER the loss
and BEFORE backward.  Data loading, the network, the optimizer and DDP stay with the caller."""
from __future__ import annotations

from typing import Optional

import torch

from .bank import MemoryBank, dequeue_and_enqueue
from .loss import get_seg_loss


class LossStepTimer:
    """CUDA-event timer of the loss step: replaces the wall-clock "Loss Time" meter of
    trainer_contrastive.py:225-255, which needs a device synchronisation to mean anything.  ``start()`` / ``stop()`` record
    events on the current stream and never block; ``read()`` returns the duration (ms) of the most recent step whose
    events have completed — i.e. it lags the training loop by a step or two instead of stalling it."""

    def __init__(self, depth: int = 4):
        self._ring = [None] * max(2, depth)
        self._i = 0
        self._open = None
        self.last_ms: Optional[float] = None

    def start(self) -> None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        self._open = ev

    def stop(self) -> None:
        if self._open is None:
            return
        self._open[1].record()
        self._ring[self._i % len(self._ring)] = self._open
        self._i += 1
        self._open = None

    def read(self) -> Optional[float]:
        for k in range(1, len(self._ring) + 1):              # newest first
            ev = self._ring[(self._i - k) % len(self._ring)]
            if ev is not None and ev[1].query():
                self.last_ms = ev[0].elapsed_time(ev[1])
                break
        return self.last_ms


class ContrastTrainerHook:
    def __init__(self, configer, bank: Optional[MemoryBank] = None, timer: Optional[LossStepTimer] = None):
        self.configer = configer
        self.timer = timer               # optional CUDA-event timing of loss_step (no synchronisation)
        self.pixel_loss = get_seg_loss(configer)
        self.with_contrast = configer.exists("contrast")
        self.contrast_warmup_iters = configer.get("contrast", "warmup_iters") if configer.exists("contrast", "warmup_iters") else 0
        # Q10: the key's presence (not its value) enables the bank (trainer_contrastive.py:78)
        self.with_memory = configer.exists("contrast", "with_memory")
        if self.with_memory:
            self.memory_size = configer.get("contrast", "memory_size")
            self.pixel_update_freq = configer.get("contrast", "pixel_update_freq")
        self.network_stride = configer.get("network", "stride")
        self.bank = bank
        self.rng = configer.get("contrast", "rng") if configer.exists("contrast", "rng") else "device"

    def loss_step(self, outputs: dict, targets: torch.Tensor, iters: int, distributed: bool = True) -> torch.Tensor:
        """outputs: the model's dict ({'seg','embed'[,'seg_aux','key','lb_key']}); returns the loss to backward."""
        with_embed = iters >= self.contrast_warmup_iters
        if self.timer is not None:
            self.timer.start()
        if self.with_contrast and self.with_memory and self.bank is not None:
            self.bank.attach(outputs)
        if distributed:
            loss = self.pixel_loss(outputs, targets, with_embed=with_embed)
        else:
            # the non-distributed branch of the reference calls the loss WITHOUT with_embed, so contrast is
            # weighted 0 there (trainer_contrastive.py:245)
            loss = self.pixel_loss(outputs, targets)
        if self.with_memory and self.bank is not None and "key" in outputs and "lb_key" in outputs:
            dequeue_and_enqueue(outputs["key"], outputs["lb_key"], self.bank.segment_queue,
                                self.bank.segment_queue_ptr, self.bank.pixel_queue, self.bank.pixel_queue_ptr,
                                network_stride=self.network_stride, memory_size=self.memory_size,
                                pixel_update_freq=self.pixel_update_freq, rng=self.rng, shadow=self.bank.shadow)
        if self.timer is not None:
            self.timer.stop()
        return loss
