"""Data-parallel support for LSQ+ QAT: one flat gradient buffer, one all-reduce per step.

The reference has no distributed code; under DistributedDataParallel its
lazily created scale / shift parameters would be averaged like any other parameter.  Here the
only data that has to cross GPUs is tiny: every site's grad_scale / grad_shift (27 702 floats
for ResNet-50: 71 activation sites x 2 + 27 560 weight channels).  Activations shard by batch
and are never exchanged.

Design: the backward kernels write their reduced gradients STRAIGHT into slices of one flat
fp32 buffer (the C ABI takes any device pointer for gscale / gshift), so there is no gather or
copy step; one `all_reduce(SUM)` over that buffer (NCCL over NVLink/NVSwitch on GPUs, gloo in
the CPU tests) finishes the step, optionally on a side stream so it overlaps whatever follows.

Semantics: each rank scales by gs = grad_scaler / sqrt(numel_local * quant_max) with its LOCAL
numel, exactly as the reference op would inside DDP (csrc/ops/cuda/lsq_cuda.cu:124); the
all-reduced value is therefore sum_r grad(shard_r), optionally divided by the world size
(`average=True`, DDP's convention).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class FlatGradBuffer:
    """Flat fp32 buffer holding (grad_scale, grad_shift) of every site, addressed by name."""

    def __init__(self, slots: Sequence[Tuple[str, int]], device, dtype=torch.float32):
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        for name, n in slots:
            if name in self.offsets:
                raise ValueError(f"duplicate site name {name!r}")
            if n <= 0:
                raise ValueError(f"site {name!r} needs a positive channel count")
            self.offsets[name] = (off, n)
            off += 2 * n
        self.numel = off
        self.flat = torch.zeros(off, dtype=dtype, device=device)
        self._stream: Optional[torch.cuda.Stream] = None

    def gscale(self, name: str) -> torch.Tensor:
        off, n = self.offsets[name]
        return self.flat[off:off + n]

    def gshift(self, name: str) -> torch.Tensor:
        off, n = self.offsets[name]
        return self.flat[off + n:off + 2 * n]

    def views(self, name: str):
        return self.gscale(name), self.gshift(name)

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None, average: bool = False, side_stream: bool = False):
        """SUM-all-reduce the flat buffer.  Returns a handle with .wait(); with `side_stream`
        on CUDA the collective is enqueued on a private stream that first waits for the
        producer (current) stream, and .wait() makes the current stream wait for it."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return _Done()
        world = dist.get_world_size(group)
        if side_stream and self.flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=self.flat.device)
            cur = torch.cuda.current_stream(self.flat.device)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                if average:
                    self.flat.div_(world)
            return _StreamHandle(self._stream, self.flat.device)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(world)
        return _Done()

    def scatter_to_params(self, named_params: Dict[str, Tuple[torch.nn.Parameter, Optional[torch.nn.Parameter]]]):
        """Point `.grad` of each (scale, shift) parameter pair at its slice (no copy)."""
        for name, (scale, shift) in named_params.items():
            gs, gb = self.views(name)
            scale.grad = gs.view_as(scale)
            if shift is not None:
                shift.grad = gb.view_as(shift)


class _Done:
    def wait(self):
        return None


class _StreamHandle:
    def __init__(self, stream, device):
        self.stream, self.device = stream, device

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


def shard_batch(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of rank's slice of a batch of n (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def expected_allreduced(per_rank: List[torch.Tensor], average: bool = False) -> torch.Tensor:
    """Reference semantics of the all-reduced buffer: sum over ranks (/ world if average)."""
    out = torch.stack([t.double() for t in per_rank]).sum(0)
    return out / len(per_rank) if average else out
