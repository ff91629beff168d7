"""cmb_datasummary mirror: running moments, Pebay merge, cross-GPU fold.

Reference: include/cmb_datasummary.h:42-51 (struct), src/cmb_datasummary.c:144-166
(add), :93-131 (merge).  The arithmetic lives in the C-ABI library
(cimba_b200/csrc/summary.cuh, shared by host and device); this class only holds
the struct.  The one cross-GPU step of the whole engine is here: every rank
reduces its shard on the device, the 8-double summaries are all-gathered (NCCL
on GPUs, gloo in the CPU tests) and folded in rank order.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional

import torch

from ._lib import DataSummaryStruct, lib, check


class DataSummary:
    """Same fields and semantics as the reference's ``struct cmb_datasummary``."""

    def __init__(self) -> None:
        self._s = DataSummaryStruct()
        lib.cimba_b200_datasummary_initialize(C.byref(self._s))

    # --- reference API names (cmb_datasummary_*) ---
    def add(self, y: float) -> int:
        return int(lib.cimba_b200_datasummary_add(C.byref(self._s), float(y)))

    @staticmethod
    def merge(a: "DataSummary", b: "DataSummary") -> "DataSummary":
        out = DataSummary()
        lib.cimba_b200_datasummary_merge(C.byref(out._s), C.byref(a._s), C.byref(b._s))
        return out

    def count(self) -> int:
        return int(self._s.count)

    def min(self) -> float:
        return self._s.min

    def max(self) -> float:
        return self._s.max

    def mean(self) -> float:
        return lib.cimba_b200_datasummary_mean(C.byref(self._s))

    def variance(self) -> float:
        return lib.cimba_b200_datasummary_variance(C.byref(self._s))

    def stddev(self) -> float:
        return lib.cimba_b200_datasummary_stddev(C.byref(self._s))

    def half_width_95(self) -> float:
        """1.96 * stddev / sqrt(n), as printed by benchmark/MM1_multi.c:151-157."""
        n = self.count()
        return 1.96 * self.stddev() / math.sqrt(n) if n > 1 else float("nan")

    # --- flat form {count,min,max,m1,m2,m3,m4,0} used on the wire ---
    def to_list(self) -> List[float]:
        s = self._s
        return [float(s.count), s.min, s.max, s.m1, s.m2, s.m3, s.m4, 0.0]

    @staticmethod
    def from_list(v: Iterable[float]) -> "DataSummary":
        v = list(v)
        out = DataSummary()
        s = out._s
        s.count = int(v[0])
        s.min, s.max, s.m1, s.m2, s.m3, s.m4 = v[1:7]
        return out

    @staticmethod
    def of(values: Iterable[float]) -> "DataSummary":
        out = DataSummary()
        for y in values:
            out.add(y)
        return out

    def __repr__(self) -> str:
        return (f"DataSummary(n={self.count()}, mean={self.mean():.9g}, "
                f"sd={self.stddev():.6g}, min={self.min():.6g}, max={self.max():.6g})")


def summarize_on_device(sum_wait: torch.Tensor, objects: torch.Tensor) -> torch.Tensor:
    """Device reduction of avg = sum_wait/objects over trials -> 8-double tensor."""
    if not sum_wait.is_cuda:
        raise ValueError("summarize_on_device needs CUDA tensors (no CPU path)")
    out = torch.empty(8, dtype=torch.float64, device=sum_wait.device)
    with torch.cuda.device(sum_wait.device):
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.cimba_b200_summarize(sum_wait.data_ptr(), objects.data_ptr(),
                                       sum_wait.numel(), out.data_ptr(), C.c_void_p(stream)))
    return out


def merge_across_ranks(local: torch.Tensor, group=None) -> DataSummary:
    """All-gather the per-rank 8-double summaries and fold them in rank order.

    Works on whatever device ``local`` lives on (CUDA+NCCL in production,
    CPU+gloo in the tests).  With no process group initialised it just wraps the
    local summary.
    """
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous(), group=group)
    else:
        parts = [local]
    acc: Optional[DataSummary] = None
    for p in parts:
        s = DataSummary.from_list(p.detach().cpu().tolist())
        acc = s if acc is None else DataSummary.merge(acc, s)
    return acc
