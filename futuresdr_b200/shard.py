"""Multi-GPU sharding of one logical sample stream (new design -- the reference
is single-process, single-device).

One process per GPU.  The stream is cut into contiguous time ranges: at step ``t`` rank ``r``
of ``W`` owns samples ``[(t*W + r)*S, (t*W + r + 1)*S)``.  A FIR output needs the ``ntaps-1``
inputs before it, so the only data-path exchange is the *overlap region*: every rank
contributes the last ``ntaps-1`` samples of its chunk to one fixed-size NCCL all-gather
(``(ntaps-1)*8`` bytes per rank, e.g. 2 KiB for 256 taps) and copies its left neighbour's
tail into the halo slot in front of its own chunk (rank 0 takes rank ``W-1``'s tail of the
previous step; at the very start of the stream there is no history, so the first chunk yields
``S-(ntaps-1)`` outputs exactly like the reference, perf/fir/fir.rs:97).

The per-rank compute is the same C-ABI FIR plan as the single-GPU path; ``compute`` can be
replaced (tests run this file's logic on CPU over gloo with the oracle as the kernel).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist


class ShardedFir:
    def __init__(self, taps, chunk_items: int, sample_dtype=np.complex64, decim: int = 1,
                 device: Optional[torch.device] = None, group=None,
                 compute: Optional[Callable] = None, algo: int = 0):
        self.taps = np.ascontiguousarray(taps)
        self.ntaps = int(self.taps.size)
        self.S = int(chunk_items)
        self.decim = int(decim)
        # History a chunk needs so its first output continues the global output sequence without
        # gap or overlap: H = ceil((ntaps-1)/D)*D  (== ntaps-1 for D == 1).  The reference keeps
        # slice starts at multiples of D because it consumes n*D items per call
        # (decimating_fir.rs:94); H being a multiple of D preserves that phase across shards.
        self.halo = ((self.ntaps + self.decim - 2) // self.decim) * self.decim
        if self.S % self.decim:
            raise ValueError("chunk_items must be a multiple of decim so shard phases align")
        if self.S < self.halo:
            raise ValueError("chunk shorter than the FIR history")
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.tdtype = torch.complex64 if np.dtype(sample_dtype) == np.complex64 else torch.float32
        # [halo | chunk] contiguous so the kernel sees history + new samples as one slice
        self.xbuf = torch.zeros(self.halo + self.S, dtype=self.tdtype, device=self.device)
        self.tails = torch.zeros(self.world, max(self.halo, 1), dtype=self.tdtype, device=self.device)
        self.have_history = False          # becomes True after the first step of the stream
        self.step_index = 0
        if compute is None:
            from .filters import DecimatingFirFilter
            self._filter = DecimatingFirFilter(self.decim, self.taps, sample_dtype, algo=algo)
            compute = self._filter.filter
        self.compute = compute

    @property
    def chunk(self) -> torch.Tensor:
        """The rank's writable chunk (fill this with the step's samples)."""
        return self.xbuf[self.halo:]

    def exchange_halo(self):
        """All-gather of the overlap region; installs the left neighbour's tail as history."""
        if self.halo == 0:
            return
        my_tail = self.xbuf[self.S:]                      # last `halo` samples of [halo|chunk]
        if self.world > 1:
            # complex tensors travel as their (re, im) float view
            real = (lambda t: torch.view_as_real(t) if t.is_complex() else t)
            dist.all_gather_into_tensor(real(self.tails).reshape(-1), real(my_tail.contiguous()).reshape(-1),
                                        group=self.group)
        else:
            self.tails[0].copy_(my_tail)

    def step(self, out: torch.Tensor):
        """Filter this step's chunk (already written into ``self.chunk``).

        Returns (consumed, produced, status).  ``out`` must hold ``S // decim`` items.
        """
        # history for this chunk: left neighbour's tail of THIS step (r > 0), or rank W-1's
        # tail of the PREVIOUS step (r == 0), which `tails` still holds from the last exchange.
        if self.rank == 0:
            if self.have_history and self.halo:
                self.xbuf[:self.halo].copy_(self.tails[self.world - 1][:self.halo])
            self.exchange_halo()
        else:
            self.exchange_halo()
            if self.halo:
                self.xbuf[:self.halo].copy_(self.tails[self.rank - 1][:self.halo])
        first_of_stream = self.rank == 0 and not self.have_history
        src = self.xbuf[self.halo:] if first_of_stream else self.xbuf
        res = self.compute(src, out)
        self.have_history = True
        self.step_index += 1
        return res
