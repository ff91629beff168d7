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

from ._lib import DataSummaryStruct, WtdSummaryStruct, lib, check


def _printed(fn, struct, lead_ins: bool) -> str:
    """Run one of the C print functions on a libc FILE* and return what it wrote."""
    import tempfile
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    with tempfile.NamedTemporaryFile() as tmp:
        fp = libc.fopen(tmp.name.encode(), b"w")
        if not fp:
            raise OSError(f"fopen({tmp.name}) failed")
        fn(C.byref(struct), fp, 1 if lead_ins else 0)
        libc.fclose(fp)
        with open(tmp.name) as f:
            return f.read()


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

    def skewness(self) -> float:
        return lib.cimba_b200_datasummary_skewness(C.byref(self._s))

    def kurtosis(self) -> float:
        """Sample excess kurtosis (src/cmb_datasummary.c:233-249)."""
        return lib.cimba_b200_datasummary_kurtosis(C.byref(self._s))

    def line(self, lead_ins: bool = True) -> str:
        """The text cmb_datasummary_print writes (src/cmb_datasummary.c:168-212)."""
        return _printed(lib.cimba_b200_datasummary_print, self._s, lead_ins)

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


class WtdSummary:
    """Same fields and semantics as the reference's ``struct cmb_wtdsummary``
    (include/cmb_wtdsummary.h:41-44; add src/cmb_wtdsummary.c:82-137, merge :152-194)."""

    def __init__(self) -> None:
        self._s = WtdSummaryStruct()
        lib.cimba_b200_wtdsummary_initialize(C.byref(self._s))

    def add(self, x: float, w: float) -> int:
        return int(lib.cimba_b200_wtdsummary_add(C.byref(self._s), float(x), float(w)))

    @staticmethod
    def merge(a: "WtdSummary", b: "WtdSummary") -> "WtdSummary":
        out = WtdSummary()
        lib.cimba_b200_wtdsummary_merge(C.byref(out._s), C.byref(a._s), C.byref(b._s))
        return out

    def count(self) -> int:
        return int(self._s.base.count)

    def wsum(self) -> float:
        return self._s.wsum

    def mean(self) -> float:
        return lib.cimba_b200_wtdsummary_mean(C.byref(self._s))

    def variance(self) -> float:
        return lib.cimba_b200_wtdsummary_variance(C.byref(self._s))

    def stddev(self) -> float:
        return lib.cimba_b200_wtdsummary_stddev(C.byref(self._s))

    def skewness(self) -> float:
        return lib.cimba_b200_wtdsummary_skewness(C.byref(self._s))

    def kurtosis(self) -> float:
        return lib.cimba_b200_wtdsummary_kurtosis(C.byref(self._s))

    def line(self, lead_ins: bool = True) -> str:
        """The text cmb_wtdsummary_print writes (it prints the data summary part)."""
        return _printed(lib.cimba_b200_wtdsummary_print, self._s, lead_ins)

    # --- the 8-word row {count (u64), min, max, m1..m4, wsum (f64 bits)} the engine writes ---
    def to_row(self) -> List[int]:
        import struct
        b = self._s.base
        return [int(b.count)] + [struct.unpack("<Q", struct.pack("<d", v))[0]
                                 for v in (b.min, b.max, b.m1, b.m2, b.m3, b.m4, self._s.wsum)]

    @staticmethod
    def from_row(row: Iterable[int]) -> "WtdSummary":
        import struct
        row = [int(v) & 0xFFFFFFFFFFFFFFFF for v in row]
        f = [struct.unpack("<d", struct.pack("<Q", v))[0] for v in row[1:8]]
        out = WtdSummary()
        b = out._s.base
        b.count = row[0]
        b.min, b.max, b.m1, b.m2, b.m3, b.m4 = f[:6]
        out._s.wsum = f[6]
        return out

    def fields(self) -> List[float]:
        b = self._s.base
        return [float(b.count), b.min, b.max, b.m1, b.m2, b.m3, b.m4, self._s.wsum]

    def __repr__(self) -> str:
        return f"WtdSummary(n={self.count()}, mean={self.mean():.9g}, wsum={self.wsum():.9g})"


def summarize_weighted_on_device(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """cmb_wtdsummary_add over device-resident (x, w) -> one 8-word row (int64 tensor)."""
    if not x.is_cuda:
        raise ValueError("summarize_weighted_on_device needs CUDA tensors (no CPU path)")
    out = torch.empty(8, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.cimba_b200_summarize_weighted(x.data_ptr(), w.data_ptr(), x.numel(),
                                                out.data_ptr(), C.c_void_p(stream)))
    return out


def merge_weighted_rows_on_device(rows: torch.Tensor) -> torch.Tensor:
    """cmb_wtdsummary_merge over per-trial rows [n, 8] on the device -> one row."""
    if not rows.is_cuda:
        raise ValueError("merge_weighted_rows_on_device needs a CUDA tensor (no CPU path)")
    rows = rows.contiguous()
    out = torch.empty(8, dtype=torch.int64, device=rows.device)
    with torch.cuda.device(rows.device):
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.cimba_b200_merge_weighted_rows(rows.data_ptr(), rows.shape[0], out.data_ptr(),
                                                 C.c_void_p(stream)))
    return out


def merge_weighted_across_ranks(local_row: torch.Tensor, group=None) -> WtdSummary:
    """All-gather the per-rank rows and fold them in rank order with cmb_wtdsummary_merge
    (the weighted half of SURVEY.md section 8e's one exchange step)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        parts = [torch.empty_like(local_row) for _ in range(world)]
        dist.all_gather(parts, local_row.contiguous(), group=group)
    else:
        parts = [local_row]
    acc: Optional[WtdSummary] = None
    for p in parts:
        s = WtdSummary.from_row(p.detach().cpu().tolist())
        acc = s if acc is None else WtdSummary.merge(acc, s)
    return acc


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
