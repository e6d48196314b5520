"""One-shot all-reduce over NVLink peer memory for the tensor-parallel decode path (csrc/comm.cu, C ABI
b200awq_comm_*).  One process per GPU; torch.distributed is the plumbing that carries the 64-byte CUDA
IPC handles between the ranks once, the data path afterwards is ONE kernel launch per collective:

    ar = OneShotAllReduce(group=None, max_elems=65536)     # after dist.init_process_group("nccl", ...)
    y = linear_forward(...row-parallel shard...)           # fp16 partial [M, hidden]
    ar(y)                                                  # in place; CUDA-graph capturable

Messages larger than the symmetric buffer (prefill-sized) go through NCCL (`dist.all_reduce`).
The reference has no collective to mirror (multi-GPU = accelerate layer placement, awq/models/base.py:527-535).
"""
from __future__ import annotations

import ctypes

import torch

from . import ext
from ._cabi import B200AwqError, check, lib


class OneShotAllReduce:
    def __init__(self, group=None, max_elems: int = 65536, device=None):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise B200AwqError("b200awq: OneShotAllReduce needs an initialised torch.distributed process group")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_elems = int(max_elems)
        self._h = ctypes.c_void_p()
        with ext._DeviceGuard(self.dev):
            check(lib.b200awq_comm_create(self.rank, self.world, self.max_elems, ctypes.byref(self._h)),
                  "b200awq_comm_create")
            mine = ctypes.create_string_buffer(64)
            check(lib.b200awq_comm_ipc_handle(self._h, mine), "b200awq_comm_ipc_handle")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine.raw), group=group)
            blob = ctypes.create_string_buffer(b"".join(handles), 64 * self.world)
            check(lib.b200awq_comm_open(self._h, blob), "b200awq_comm_open")
        dist.barrier(group=group)          # every rank has mapped every buffer before the first push

    def __call__(self, y: torch.Tensor) -> torch.Tensor:
        """Sum of `y` over the ranks, in place.  fp16, contiguous, on this comm's device."""
        if self.world == 1:
            return y
        n = y.numel()
        if y.dtype != torch.float16 or not y.is_contiguous() or n % 8 or n > self.max_elems or y.data_ptr() % 16:
            import torch.distributed as dist

            dist.all_reduce(y, group=self.group)     # prefill-sized / odd-shaped messages: NCCL
            return y
        with ext._DeviceGuard(self.dev):
            check(lib.b200awq_comm_all_reduce(self._h, y.data_ptr(), n, ext._stream(self.dev)), "b200awq_comm_all_reduce")
        return y

    def check(self) -> None:
        """Synchronises; raises if a wait inside a collective ever timed out (a peer died)."""
        check(lib.b200awq_comm_error(self._h), "b200awq_comm_error (a peer did not arrive within 2 s)")

    def close(self) -> None:
        if self._h:
            lib.b200awq_comm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
