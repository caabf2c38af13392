"""Bulk descriptor extraction from HOST memory (BASELINE config 3): the end-to-end call a user makes.

The reference's loops do `data.cuda()` -> `model(data)` -> `.cpu().numpy()` batch by batch
(hardnet/HardNet.py:453-461). `extract_descriptors` does the same job as a three-stage pipeline: pinned
host -> device copies on one stream, the B200 forward on a second, device -> pinned host on a third, with
double-buffered device staging, so PCIe transfers overlap the kernels.
"""
from __future__ import annotations

import torch


class DescriptorExtractor:
    def __init__(self, model, batch: int | None = None, out_dtype: torch.dtype = torch.float32, device=None,
                 in_dtype: torch.dtype = torch.float32):
        self.model = model
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if batch is None:
            # one conv-stack pass of the engine (128 patches per SM unless the model was built with chunk_patches): no
            # partial passes inside a batch and a short pipeline fill / drain
            batch = getattr(model, "_chunk_patches", 0) or 128 * torch.cuda.get_device_properties(self.device).multi_processor_count
        self.batch = int(batch)
        self.out_dtype = out_dtype
        with torch.cuda.device(self.device):
            self.copy_in = torch.cuda.Stream()
            self.compute = torch.cuda.Stream()
            self.copy_out = torch.cuda.Stream()
            self.d_in = [torch.empty((self.batch, 1, 32, 32), dtype=in_dtype, device=self.device) for _ in range(2)]
            self.d_out = [torch.empty((self.batch, 128), dtype=out_dtype, device=self.device) for _ in range(2)]
            self.in_ready = [torch.cuda.Event() for _ in range(2)]
            self.in_free = [torch.cuda.Event() for _ in range(2)]
            self.out_ready = [torch.cuda.Event() for _ in range(2)]
            self.out_free = [torch.cuda.Event() for _ in range(2)]

    @torch.no_grad()
    def __call__(self, patches_host: torch.Tensor, out_host: torch.Tensor | None = None) -> torch.Tensor:
        """patches_host: [N,1,32,32] CPU tensor (pinned for full speed); returns [N,128] on the host."""
        n = patches_host.size(0)
        if out_host is None:
            out_host = torch.empty((n, 128), dtype=self.out_dtype, pin_memory=True)
        eng_model = self.model
        with torch.cuda.device(self.device):
            for k, start in enumerate(range(0, n, self.batch)):
                slot = k & 1
                m = min(self.batch, n - start)
                with torch.cuda.stream(self.copy_in):
                    if k >= 2:
                        self.copy_in.wait_event(self.in_free[slot])
                    self.d_in[slot][:m].copy_(patches_host[start:start + m], non_blocking=True)
                    self.in_ready[slot].record(self.copy_in)
                with torch.cuda.stream(self.compute):
                    self.compute.wait_event(self.in_ready[slot])
                    if k >= 2:
                        self.compute.wait_event(self.out_free[slot])
                    eng_model(self.d_in[slot][:m], out=self.d_out[slot][:m])
                    self.in_free[slot].record(self.compute)
                    self.out_ready[slot].record(self.compute)
                with torch.cuda.stream(self.copy_out):
                    self.copy_out.wait_event(self.out_ready[slot])
                    out_host[start:start + m].copy_(self.d_out[slot][:m], non_blocking=True)
                    self.out_free[slot].record(self.copy_out)
            self.copy_out.synchronize()
        return out_host


def extract_descriptors(model, patches_host: torch.Tensor, batch: int | None = None, out_dtype=torch.float32) -> torch.Tensor:
    return DescriptorExtractor(model, batch=batch, out_dtype=out_dtype, in_dtype=patches_host.dtype)(patches_host)
