"""Training-loop helpers of SURVEY.md §8(f) row 3 that need no autograd.

``update_ema`` keeps the reference's signature (engine_generation.py:29-39, called once per optimizer step from
``train_one_epoch``): ``targ.detach().mul_(rate).add_(src, alpha=1 - rate)`` for every parameter pair. The reference
issues two tiny elementwise launches per parameter (1 274 per step for the default denoiser); here the whole list is
ONE launch of ``rald_ema_update`` over a cached device table of pointers (rebuilt only when a tensor of the list moves),
bit-identical to ATen's arithmetic. There is no CPU path."""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import torch

from . import _lib

_TABLES: Dict[Tuple, Tuple[torch.Tensor, int, int, int]] = {}
_MAX_TABLES = 8


def _table(targets: List[torch.Tensor], sources: List[torch.Tensor]):
    key = (targets[0].device.index, tuple(t.data_ptr() for t in targets), tuple(s.data_ptr() for s in sources),
           tuple(t.numel() for t in targets))
    hit = _TABLES.get(key)
    if hit is not None:
        return hit
    chunk = int(_lib.lib().rald_ema_chunk_elems())
    n = len(targets)
    first, total = [], 0
    for t in targets:
        first.append(total)
        total += (t.numel() + chunk - 1) // chunk
    first.append(total)
    host = torch.tensor([t.data_ptr() for t in targets] + [s.data_ptr() for s in sources] +
                        [t.numel() for t in targets] + first, dtype=torch.int64)
    dev = host.to(targets[0].device)
    if len(_TABLES) >= _MAX_TABLES:
        _TABLES.pop(next(iter(_TABLES)))
    entry = (dev, n, total, sum(t.numel() for t in targets))
    _TABLES[key] = entry
    return entry


@torch.no_grad()
def update_ema(target_params: Iterable[torch.Tensor], source_params: Iterable[torch.Tensor], rate: float = 0.99) -> None:
    """Drop-in for engine_generation.py:29-39. All tensors fp32, contiguous, on one CUDA device."""
    targets = [t.detach() for t in target_params]
    sources = [s.detach() for s in source_params]
    if len(sources) < len(targets):      # zip() semantics of the reference: the shorter list wins
        targets = targets[:len(sources)]
    sources = sources[:len(targets)]
    if not targets:
        return
    dev = targets[0].device
    if dev.type != "cuda":
        raise _lib.RaldError("rald_b200 runs on CUDA devices only (no CPU fallback)")
    for t, s in zip(targets, sources):
        if t.dtype != torch.float32 or s.dtype != torch.float32:
            raise _lib.RaldError("update_ema: parameters must be fp32 (the checkpoint contract of the reference)")
        if t.device != dev or s.device != dev:
            raise _lib.RaldError("update_ema: all parameters must live on one device")
        if t.shape != s.shape:
            raise _lib.RaldError(f"update_ema: shape mismatch {tuple(t.shape)} vs {tuple(s.shape)}")
        if not (t.is_contiguous() and s.is_contiguous()):
            raise _lib.RaldError("update_ema: parameters must be contiguous")
    table, n, chunks, elems = _table(targets, sources)
    with torch.cuda.device(dev):
        _lib.call("rald_ema_update", table.data_ptr(), n, chunks, elems, float(rate), float(1 - rate), _lib.cur_stream())
    # The kernel writes through raw device pointers, which autograd's version counters do not see — but the runtimes
    # (DitRuntime / AeRuntime._signature) detect weight changes through exactly those counters. Bump them as the
    # in-place torch ops this replaces did, so that an EMA into a module's own parameters repacks its bf16 weights.
    torch.autograd.graph.increment_version(targets)
