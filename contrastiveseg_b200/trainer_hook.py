"""The loss call of segmentor/trainer_contrastive.py:209-252, as a hook a trainer can call.

Only the hook point is replaced: warm-up gate, bank attach, loss, enqueue AFTER the loss
and BEFORE backward.  Data loading, the network, the optimizer and DDP stay with the caller."""
from __future__ import annotations

from typing import Optional

import torch

from .bank import MemoryBank, dequeue_and_enqueue
from .loss import get_seg_loss


class ContrastTrainerHook:
    def __init__(self, configer, bank: Optional[MemoryBank] = None):
        self.configer = configer
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
        return loss
