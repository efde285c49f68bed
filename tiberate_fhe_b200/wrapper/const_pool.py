"""tiberate/libs/wrapper/const_pool.py mirror (schemas: csrc/ops/constant_mem.cpp:3-10; implementation
csrc/ops/cuda/constant_pool_cuda.cu:8-111).

The reference keeps its per-prime constants in a 4 KiB __constant__ pool per device, written by
`upload_tensor_list` at byte offsets counted from the left or from the right end, and read back by
`read_constant_chunk` (its only live unit test, tests/test_constant_mem.py, round-trips this).  libtb200
reads its constants from per-context device tables instead (no 64-prime cap, not process-global), so the
pool here is plain device memory with the same addressing: uploads and read-backs behave as in the
reference, and nothing on the compute path depends on it.
"""

from __future__ import annotations

import torch

MAX_CONST_BYTES = 4 * 1024  # constant_mem.cuh:7
_pools: dict[int, torch.Tensor] = {}


def _pool(device_id: int) -> torch.Tensor:
    p = _pools.get(device_id)
    if p is None:
        p = _pools[device_id] = torch.zeros(MAX_CONST_BYTES, dtype=torch.uint8, device=f"cuda:{device_id}")
    return p


def upload_tensor_list(tensor_list, offset_list, layout: int, device_id: int) -> None:
    if len(tensor_list) != len(offset_list):
        raise RuntimeError("Mismatch: tensor list and offset list must have same length")
    pool = _pool(int(device_id))
    for t, off in zip(tensor_list, offset_list):
        if not t.is_contiguous():
            raise RuntimeError("Tensors must be contiguous")
        n = t.numel() * t.element_size()
        start = MAX_CONST_BYTES - int(off) - n if layout == 1 else int(off)
        if start < 0 or start + n > MAX_CONST_BYTES:
            raise RuntimeError("Upload exceeds constant memory")
        pool[start:start + n] = t.reshape(-1).view(torch.uint8).to(pool.device)


def read_constant_chunk(dummy: torch.Tensor, offset_bytes: int, count: int, dtype, layout: int) -> torch.Tensor:
    dev = dummy.device.index or 0
    n = int(count) * torch.empty(0, dtype=dtype).element_size()
    start = MAX_CONST_BYTES - int(offset_bytes) - n if layout == 1 else int(offset_bytes)
    if start < 0 or start + n > MAX_CONST_BYTES:
        raise RuntimeError("Read exceeds constant memory")
    return _pool(dev)[start:start + n].clone().view(dtype)
