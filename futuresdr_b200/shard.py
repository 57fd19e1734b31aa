"""Multi-GPU sharding of one logical sample stream (new design -- the reference
is single-process, single-device).

One process per GPU.  The stream is cut into contiguous time ranges: at step ``t`` rank ``r``
of ``W`` owns samples ``[(t*W + r)*S, (t*W + r + 1)*S)``.  A FIR output needs the ``ntaps-1``
inputs before it, so the only data-path exchange is the *overlap region*: every rank
contributes the last ``H = ceil((ntaps-1)/D)*D`` samples of its chunk to one fixed-size NCCL
all-gather (2 KiB per rank for 256 taps) and installs its left neighbour's tail as history in
front of its own chunk (rank 0 takes rank ``W-1``'s tail of the previous step; at the very start
of the stream there is no history, so the first chunk yields ``S-(ntaps-1)`` outputs exactly like
the reference, perf/fir/fir.rs:97).

The exchange is latency-bound (~tens of microseconds), so it is taken off the critical path: the
all-gather is launched asynchronously, the outputs whose window lies entirely inside the rank's
own chunk (all but the first ``split/D``, split = H rounded up to an aligned slice) are computed meanwhile, and only a small second launch for
the first ``H/D`` outputs waits for the neighbour's tail.

The per-rank compute is the same C-ABI FIR plan as the single-GPU path; ``compute`` can be
replaced (tests run this file's logic on CPU over gloo with the oracle as the kernel).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class ShardedFir:
    def __init__(self, taps, chunk_items: int, sample_dtype=np.complex64, decim: int = 1,
                 device: Optional[torch.device] = None, group=None,
                 compute: Optional[Callable] = None, algo: int = 0, overlap: bool = True):
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
        # Split point of the overlapped step: outputs [0, split/D) need the neighbour's tail, the rest only
        # the rank's own chunk.  It is H rounded up so that both xbuf[split:] and out[split/D:] stay 16-byte
        # aligned -- an unaligned slice would push the FIR plan onto its scalar fallback kernel.
        isz = np.dtype(sample_dtype).itemsize
        q = self.decim * max(1, 16 // isz)
        self.split = ((self.halo + q - 1) // q) * q
        if self.S < self.split + self.ntaps:
            raise ValueError("chunk shorter than the FIR history")
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.tdtype = torch.complex64 if np.dtype(sample_dtype) == np.complex64 else torch.float32
        # [halo | chunk] contiguous so the kernel sees history + new samples as one slice
        self.xbuf = torch.zeros(self.halo + self.S, dtype=self.tdtype, device=self.device)
        self.tails = torch.zeros(self.world, max(self.halo, 1), dtype=self.tdtype, device=self.device)
        self.my_tail = torch.zeros(max(self.halo, 1), dtype=self.tdtype, device=self.device)
        self.prev_last_tail = torch.zeros(max(self.halo, 1), dtype=self.tdtype, device=self.device)
        self.have_history = False          # becomes True after the first step of the stream
        self.step_index = 0
        self.overlap = bool(overlap)
        if compute is None:
            from .filters import DecimatingFirFilter
            self._filter = DecimatingFirFilter(self.decim, self.taps, sample_dtype, algo=algo)
            compute = self._filter.filter
            # the split/D outputs behind the exchange are a few hundred items: the CUDA-core kernel has no
            # TMEM / tap-table prologue, so that second launch costs ~half of a tensor-kernel launch
            self._head_filter = DecimatingFirFilter(self.decim, self.taps, sample_dtype, algo=_lib.ALGO_DIRECT)
            head_compute = self._head_filter.filter
        else:
            head_compute = compute
        self.compute = compute
        self.head_compute = head_compute

    @property
    def chunk(self) -> torch.Tensor:
        """The rank's writable chunk (fill this with the step's samples)."""
        return self.xbuf[self.halo:]

    def _exchange(self, async_op: bool):
        """All-gather of the overlap region (every rank's last H samples). Returns a work handle or None."""
        if self.halo == 0:
            return None
        my_tail = self.xbuf[self.S:]                     # last `halo` samples of [halo|chunk], contiguous
        if self.world == 1:
            self.tails[0].copy_(my_tail)
            return None
        # complex tensors travel as their (re, im) float view
        real = (lambda t: torch.view_as_real(t) if t.is_complex() else t)
        return dist.all_gather_into_tensor(real(self.tails).reshape(-1), real(my_tail).reshape(-1),
                                           group=self.group, async_op=async_op)

    def step(self, out: torch.Tensor):
        """Filter this step's chunk (already written into ``self.chunk``).

        Returns (consumed, produced, status).  ``out`` must hold ``S // decim`` items.
        """
        H, D, N = self.halo, self.decim, self.ntaps
        first_of_stream = self.rank == 0 and not self.have_history
        if self.world == 1:
            # single rank: the history of step t+1 is the tail of step t's own chunk -- one device copy
            # per step, no gather buffers
            res = self.compute(self.xbuf[H:] if first_of_stream else self.xbuf, out)
            if H:
                self.xbuf[:H].copy_(self.xbuf[self.S:])
            self.have_history = True
            self.step_index += 1
            return res
        work = self._exchange(async_op=self.overlap)
        if first_of_stream:
            res = self.compute(self.xbuf[H:], out)                       # no history: S-(ntaps-1) outputs
        elif self.rank == 0:
            # history = rank W-1's tail of the PREVIOUS step (saved before this step's gather started)
            if H:
                self.xbuf[:H].copy_(self.prev_last_tail[:H])
            res = self.compute(self.xbuf, out)
        elif H == 0:
            res = self.compute(self.xbuf, out)
        elif self.overlap:
            A = self.split
            nh = A // D                                                  # outputs that (may) need the neighbour's tail
            c1, p1, st = self.compute(self.xbuf[A:], out[nh:])           # windows fully inside the own chunk
            if work is not None:
                work.wait()
                work = None
            self.xbuf[:H].copy_(self.tails[self.rank - 1][:H])
            c0, p0, _ = self.head_compute(self.xbuf[:A + N - 1], out[:nh])   # the first split/D outputs
            res = (c0 + c1, p0 + p1, st)
        else:
            if work is not None:
                work.wait()
                work = None
            self.xbuf[:H].copy_(self.tails[self.rank - 1][:H])
            res = self.compute(self.xbuf, out)
        if work is not None:
            work.wait()
        if self.rank == 0 and H:
            self.prev_last_tail.copy_(self.tails[self.world - 1])        # history for the next step
        self.have_history = True
        self.step_index += 1
        return res
