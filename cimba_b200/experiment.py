"""Experiment executive: the GPU counterpart of ``cimba_run_experiment``.

Reference contract (include/cimba.h:144-147, src/cimba.c:151-188): the caller owns
an array of trial structs holding each trial's parameters and result fields; the
executive runs every trial (there: one pthread per core pulling trial indices off
an atomic counter) and returns when all results have been written in place.
Here the trial function is one of the device-resident models and trial ``i`` is
seeded with ``cmb_random_fmix64(master_seed, first_trial + i)``
(src/cmb_random.c:70-80, the scheme of test/test_cimba.c:396).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

# struct trial of benchmark/MM1_multi.c:39-45, plus optional engine outputs
TRIAL_DTYPE = np.dtype([
    ("arr_mean", "<f8"), ("srv_mean", "<f8"), ("obj_cnt", "<u8"),
    ("sum_wait", "<f8"), ("avg_wait", "<f8"),
    ("events", "<u8"), ("t_end", "<f8"), ("status", "<u4"), ("_pad", "<u4"),
])


def fmix64(seed: int, nonce: int) -> int:
    """cmb_random_fmix64 (src/cmb_random.c:70-80)."""
    return int(lib.cimba_b200_fmix64(seed & (2**64 - 1), nonce & (2**64 - 1)))


def cimba_run_experiment(experiment_array: np.ndarray, *, model: int = _lib.MODEL_MM1,
                         num_objects: int, master_seed: int, first_trial: int = 0,
                         servers: int = 1, mapping: int = 0, device: int = -1, variant: int = 0,
                         all_gpus: bool = False, max_gpus: int = 0, queue_spill_cap: int = 0, params=()) -> None:
    """Run every trial of a host-resident experiment array on the GPU, in place.

    ``experiment_array`` is a 1-D numpy structured array (any dtype that has
    ``arr_mean`` and ``srv_mean`` double fields; result fields named ``obj_cnt``,
    ``sum_wait``, ``avg_wait``, ``events``, ``t_end``, ``status``, ``max_queue`` and
    ``counters`` (8 x uint64) are filled when present).  Host->device and device->host
    copies happen inside the call.
    Raises CimbaError(ETRIAL) if any trial overflowed a device structure.
    With ``all_gpus`` the array is sharded over every visible GPU, one host thread
    each (the counterpart of the reference's one pthread per core).
    """
    arr = experiment_array
    if not isinstance(arr, np.ndarray) or arr.ndim != 1 or arr.dtype.fields is None:
        raise TypeError("experiment_array must be a 1-D numpy structured array")
    if len(arr) == 0:
        raise ValueError("num_trials must be > 0 (reference asserts this, src/cimba.c:157)")
    if not arr.flags["C_CONTIGUOUS"] or not arr.flags["WRITEABLE"]:
        raise ValueError("experiment_array must be C-contiguous and writeable")
    f = arr.dtype.fields

    def off(name, kind):
        if name not in f:
            return _lib.NO_FIELD
        if f[name][0] != np.dtype(kind):
            raise TypeError(f"field {name} must have dtype {kind}")
        return f[name][1]

    off_counters = _lib.NO_FIELD
    if "counters" in f:
        if f["counters"][0] != np.dtype(("<u8", (8,))):
            raise TypeError("field counters must be 8 x uint64")
        off_counters = f["counters"][1]

    if "arr_mean" not in f or "srv_mean" not in f:
        raise TypeError("trial struct needs arr_mean and srv_mean double fields")
    desc = _lib.Experiment(
        model=model, servers=servers, mapping=mapping, device=device, variant=variant, queue_spill_cap=queue_spill_cap,
        master_seed=master_seed & (2**64 - 1), first_trial=first_trial, num_objects=num_objects,
        off_arr_mean=off("arr_mean", "<f8"), off_srv_mean=off("srv_mean", "<f8"),
        off_obj_cnt=off("obj_cnt", "<u8"), off_sum_wait=off("sum_wait", "<f8"),
        off_avg_wait=off("avg_wait", "<f8"), off_events=off("events", "<u8"),
        off_t_end=off("t_end", "<f8"), off_status=off("status", "<u4"),
        off_max_queue=off("max_queue", "<u4"), off_counters=off_counters)
    par = (C.c_double * max(1, len(params)))(*[float(v) for v in params])
    if len(params):
        desc.params = C.cast(par, C.POINTER(C.c_double))
        desc.num_params = len(params)
    if all_gpus:
        check(lib.cimba_b200_run_experiment_all_gpus(arr.ctypes.data_as(C.c_void_p), len(arr),
                                                     arr.dtype.itemsize, C.byref(desc), max_gpus))
    else:
        check(lib.cimba_b200_run_experiment(arr.ctypes.data_as(C.c_void_p), len(arr),
                                            arr.dtype.itemsize, C.byref(desc)))


@dataclass
class TrialResults:
    """Device-resident per-trial results (SoA), one entry per trial."""
    events: torch.Tensor      # int64 view of uint64 pop counts
    objects: torch.Tensor
    t_end: torch.Tensor
    sum_wait: torch.Tensor
    status: torch.Tensor
    max_queue: torch.Tensor
    counters: torch.Tensor    # [n, 8] model counters (MODEL_GUARDED)
    trace_key: Optional[torch.Tensor] = None
    trace_time: Optional[torch.Tensor] = None

    def total_events(self) -> int:
        return int(self.events.sum().item())


def load_model(path) -> int:
    """cimba_b200_model_load: register a model library built from a .cu file written against
    cimba_b200/csrc/cmb_device.cuh (scripts/build_model.py); returns the model id to pass as ``model=``."""
    rc = lib.cimba_b200_model_load(str(path).encode())
    if rc < 0:
        check(rc)
    return int(rc)


class TrialBuffers:
    """Reusable device buffers for repeated launches of the same shape."""

    def __init__(self, num_trials: int, device: torch.device, trace_cap: int = 0,
                 model: int = _lib.MODEL_MM1, servers: int = 1, variant: int = 0, queue_spill_cap: int = 0,
                 num_objects: int = 0):
        n = num_trials
        self.n = n
        self.device = device
        self.events = torch.zeros(n, dtype=torch.int64, device=device)
        self.objects = torch.zeros(n, dtype=torch.int64, device=device)
        self.t_end = torch.zeros(n, dtype=torch.float64, device=device)
        self.sum_wait = torch.zeros(n, dtype=torch.float64, device=device)
        self.status = torch.zeros(n, dtype=torch.int32, device=device)
        self.max_queue = torch.zeros(n, dtype=torch.int32, device=device)
        self.counters = torch.zeros((n, 8), dtype=torch.int64, device=device)
        self.trace_cap = trace_cap
        self.trace_key = self.trace_time = None
        if trace_cap:
            self.trace_key = torch.zeros((n, trace_cap), dtype=torch.int64, device=device)
            self.trace_time = torch.zeros((n, trace_cap), dtype=torch.float64, device=device)
        job = _lib.DeviceJob(model=model, num_trials=n, servers=servers, variant=variant, queue_spill_cap=queue_spill_cap,
                             num_objects=num_objects)
        ws = int(lib.cimba_b200_workspace_bytes(C.byref(job)))
        self.workspace = torch.empty(max(ws, 8), dtype=torch.uint8, device=device)
        self.workspace_bytes = ws

    def results(self) -> TrialResults:
        return TrialResults(self.events, self.objects, self.t_end, self.sum_wait,
                            self.status, self.max_queue, self.counters, self.trace_key, self.trace_time)


def launch_trials(arr_mean: torch.Tensor, srv_mean: torch.Tensor, *, num_objects: int,
                  master_seed: int, first_trial: int = 0, model: int = _lib.MODEL_MM1,
                  servers: int = 1, mapping: int = 0, buffers: Optional[TrialBuffers] = None,
                  trace_cap: int = 0, variant: int = 0, queue_spill_cap: int = 0, params=(),
                  diag: Optional[torch.Tensor] = None) -> TrialResults:
    """Asynchronously run one trial per element of ``arr_mean`` on the current stream.

    Inputs are float64 CUDA tensors already resident in HBM (the device-resident
    hot path: no host copies).  Returns device tensors; synchronise before reading.
    """
    if not (arr_mean.is_cuda and srv_mean.is_cuda):
        raise ValueError("arr_mean and srv_mean must be CUDA tensors (there is no CPU path)")
    if arr_mean.dtype != torch.float64 or srv_mean.dtype != torch.float64:
        raise TypeError("arr_mean and srv_mean must be float64")
    if arr_mean.shape != srv_mean.shape or arr_mean.dim() != 1:
        raise ValueError("arr_mean and srv_mean must be 1-D and of equal length")
    arr_mean = arr_mean.contiguous()
    srv_mean = srv_mean.contiguous()
    n = arr_mean.numel()
    if n == 0:
        raise ValueError("num_trials must be > 0 (reference asserts this, src/cimba.c:157)")
    b = buffers if buffers is not None else TrialBuffers(n, arr_mean.device, trace_cap, model, servers, variant,
                                                        queue_spill_cap, num_objects)
    if b.n != n or b.trace_cap != trace_cap:
        raise ValueError("buffers do not match this launch")
    par = (C.c_double * max(1, len(params)))(*[float(v) for v in params])
    job = _lib.DeviceJob(
        queue_spill_cap=queue_spill_cap, params=par if len(params) else None, num_params=len(params),
        diag=diag.data_ptr() if diag is not None else None,
        model=model, servers=servers, mapping=mapping, variant=variant,
        master_seed=master_seed & (2**64 - 1), first_trial=first_trial,
        num_trials=n, num_objects=num_objects,
        arr_mean=arr_mean.data_ptr(), srv_mean=srv_mean.data_ptr(),
        events=b.events.data_ptr(), objects=b.objects.data_ptr(),
        t_end=b.t_end.data_ptr(), sum_wait=b.sum_wait.data_ptr(),
        status=b.status.data_ptr(), max_queue=b.max_queue.data_ptr(), counters=b.counters.data_ptr(),
        workspace=b.workspace.data_ptr(), workspace_bytes=b.workspace_bytes,
        trace_cap=trace_cap,
        trace_key=b.trace_key.data_ptr() if trace_cap else None,
        trace_time=b.trace_time.data_ptr() if trace_cap else None)
    with torch.cuda.device(arr_mean.device):
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.cimba_b200_launch(C.byref(job), C.c_void_p(stream)))
    return b.results()


def run_trials(num_trials: int, *, arr_mean: float, srv_mean: float, num_objects: int,
               master_seed: int, first_trial: int = 0, model: int = _lib.MODEL_MM1,
               servers: int = 1, mapping: int = 0, trace_cap: int = 0, variant: int = 0,
               device: Optional[torch.device] = None, queue_spill_cap: int = 0, params=()) -> TrialResults:
    """Convenience: identical parameters for every trial, results after a sync."""
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    a = torch.full((num_trials,), arr_mean, dtype=torch.float64, device=dev)
    s = torch.full((num_trials,), srv_mean, dtype=torch.float64, device=dev)
    res = launch_trials(a, s, num_objects=num_objects, master_seed=master_seed,
                        first_trial=first_trial, model=model, servers=servers,
                        mapping=mapping, trace_cap=trace_cap, variant=variant, queue_spill_cap=queue_spill_cap,
                        params=params)
    torch.cuda.synchronize(dev)
    return res


def rng_draws(seed: int, kind: int, n: int, p0: float = 0.0, p1: float = 0.0,
              device: Optional[torch.device] = None) -> torch.Tensor:
    """n variates from the device-side stream seeded with ``seed`` (KAT helper)."""
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    out = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.cimba_b200_rng_draws(seed & (2**64 - 1), kind, p0, p1, n,
                                       out.data_ptr(), C.c_void_p(stream)))
    torch.cuda.synchronize(dev)
    return out


def rng_draws_ex(seed: int, kind: int, n: int, params=(), device: Optional[torch.device] = None) -> torch.Tensor:
    """n variates of one of the remaining cmb_random distributions (kinds 9..33, see
    include/cimba_b200.h) from the device-side stream seeded with ``seed``."""
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    out = torch.empty(n, dtype=torch.float64, device=dev)
    par = (C.c_double * max(1, len(params)))(*[float(v) for v in params])
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.cimba_b200_rng_draws_ex(seed & (2**64 - 1), kind, par, len(params), n,
                                          out.data_ptr(), C.c_void_p(stream)))
    torch.cuda.synchronize(dev)
    return out


def alias_create(probabilities):
    """cmb_random_alias_create: (uprob, alias) Vose tables for the given probabilities."""
    n = len(probabilities)
    pa = (C.c_double * n)(*[float(v) for v in probabilities])
    uprob = (C.c_uint64 * n)()
    alias = (C.c_uint32 * n)()
    check(lib.cimba_b200_alias_create(n, pa, uprob, alias))
    return list(uprob), list(alias)


# ---------------------------------------------------------------- AWACS (tutorial/tut_5_1.c, BASELINE config 5)
AWACS_TARGETS, AWACS_STRIDE = 1000, 1024
AWACS_STATE_BYTES = AWACS_STRIDE * (7 * 4 + 4 + 4 + 8)
_awacs_maps = {}        # device index -> the registered map tensor (kept alive while registered)


def awacs_set_terrain(elevations: torch.Tensor, cols: int, rows: int, geom) -> None:
    """Register the terrain every MODEL_AWACS trial on this device reads (struct terrain, tut_5_1.c:96-108):
    ``elevations`` = float32 CUDA tensor of rows*cols metres, ``geom`` = (x_scale, y_scale, x_min, x_max,
    y_min, y_max) as terrain_init (:197-294) computes them."""
    if not elevations.is_cuda or elevations.dtype != torch.float32 or elevations.numel() != cols * rows:
        raise ValueError("elevations must be a float32 CUDA tensor of rows * cols values (there is no CPU path)")
    elevations = elevations.contiguous()
    g = [float(v) for v in geom]
    desc = _lib.AwacsTerrain(map=elevations.data_ptr(), cols=cols, rows=rows, x_scale=g[0], y_scale=g[1],
                             x_min=g[2], x_max=g[3], y_min=g[4], y_max=g[5])
    with torch.cuda.device(elevations.device):
        check(lib.cimba_b200_awacs_set_terrain(C.byref(desc)))
    _awacs_maps[elevations.device.index] = elevations


def awacs_upload_terrain(elevations: np.ndarray, cols: int, rows: int, geom, device: Optional[torch.device] = None) -> None:
    """The same from a HOST array (float32, rows * cols): the library makes and owns the device copy
    (cimba_b200_awacs_upload_terrain) - what a C caller without CUDA code of its own uses."""
    m = np.ascontiguousarray(elevations, dtype=np.float32)
    if m.size != cols * rows:
        raise ValueError("elevations must hold rows * cols values")
    g = [float(v) for v in geom]
    desc = _lib.AwacsTerrain(map=m.ctypes.data, cols=cols, rows=rows, x_scale=g[0], y_scale=g[1],
                             x_min=g[2], x_max=g[3], y_min=g[4], y_max=g[5])
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        check(lib.cimba_b200_awacs_upload_terrain(C.byref(desc)))
    _awacs_maps.pop(dev.index, None)


def awacs_run(num_trials: int, *, duration_s: int, master_seed: int, first_trial: int = 0, trace_cap: int = 0,
              device: Optional[torch.device] = None):
    """Run MODEL_AWACS trials (seeds cmb_random_fmix64(master_seed, first_trial + i)) for ``duration_s`` simulated
    seconds each.  Returns (TrialResults, per_target) after a sync; ``per_target`` holds [num_trials, 1000]
    tensors x, y, alt, mode, tds, found read from the state blocks.  results.objects = targets found."""
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    b = TrialBuffers(num_trials, dev, trace_cap, _lib.MODEL_AWACS)
    job = _lib.DeviceJob(
        model=_lib.MODEL_AWACS, master_seed=master_seed & (2**64 - 1), first_trial=first_trial,
        num_trials=num_trials, num_objects=int(duration_s),
        events=b.events.data_ptr(), objects=b.objects.data_ptr(), t_end=b.t_end.data_ptr(),
        sum_wait=b.sum_wait.data_ptr(), status=b.status.data_ptr(), max_queue=b.max_queue.data_ptr(),
        counters=b.counters.data_ptr(), workspace=b.workspace.data_ptr(), workspace_bytes=b.workspace_bytes,
        trace_cap=trace_cap, trace_key=b.trace_key.data_ptr() if trace_cap else None,
        trace_time=b.trace_time.data_ptr() if trace_cap else None)
    with torch.cuda.device(dev):
        check(lib.cimba_b200_launch(C.byref(job), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize(dev)
    blocks = b.workspace[:num_trials * AWACS_STATE_BYTES].view(num_trials, AWACS_STATE_BYTES)
    col = AWACS_STRIDE * 4
    f32 = lambda k: blocks[:, k * col:(k + 1) * col].contiguous().view(torch.float32)[:, :AWACS_TARGETS]   # noqa: E731
    flags = blocks[:, 7 * col:8 * col].contiguous().view(torch.int32)[:, :AWACS_TARGETS]
    per_target = dict(x=f32(0), y=f32(1), alt=f32(2), mode=flags & 3, tds=(flags >> 4) & 7, found=(flags >> 8) & 1)
    return b.results(), per_target
