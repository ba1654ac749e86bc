"""Multi-GPU plumbing: independent utterances are batch-sharded across ranks (one process per GPU); the only
collective on the path is one all-gather of the output latents (the reference has no cross-sample reduction, so no
data-path exchange exists: SURVEY.md section 8e). Works with the NCCL backend on GPUs and with gloo on CPU (tests)."""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

_COMM: Optional[C.c_void_p] = None       # dv_comm of this process (one GPU per process), set by init_native_comm


def _nccl_path() -> Optional[bytes]:
    """The libnccl the host process already uses: torch's bundled copy (so both sides share one library instance)."""
    env = os.environ.get("DV_NCCL_LIB")
    if env:
        return env.encode()
    site = os.path.dirname(os.path.dirname(torch.__file__))
    hits = glob.glob(os.path.join(site, "nvidia", "nccl", "lib", "libnccl.so*"))
    return hits[0].encode() if hits else None


def init_native_comm(device: torch.device) -> bool:
    """Create the library-owned NCCL communicator (C ABI dv_comm_*) so that the latent all-gather is issued in-stream behind
    the sampling loop instead of through torch.distributed. Needs an initialised process group (the 128-byte unique id is
    broadcast through it); returns False (and leaves the torch.distributed path in place) when NCCL cannot be loaded."""
    global _COMM
    if _COMM is not None or not dist.is_initialized() or dist.get_world_size() == 1 or device.type != "cuda":
        return _COMM is not None
    from . import _lib
    lib = _lib.load()
    path = _nccl_path()
    ident = torch.zeros(128, dtype=torch.uint8)
    ok = torch.ones(1, dtype=torch.int32)
    if dist.get_rank() == 0 and lib.dv_comm_unique_id(path, C.c_void_p(ident.data_ptr())) != 0:
        ok.zero_()
    on_dev = dist.get_backend() == "nccl"
    ident_t, ok_t = (ident.to(device), ok.to(device)) if on_dev else (ident, ok)
    dist.broadcast(ok_t, 0)
    dist.broadcast(ident_t, 0)
    if int(ok_t.cpu()) == 0:
        return False
    ident = ident_t.cpu().contiguous()
    comm = C.c_void_p()
    with torch.cuda.device(device):
        rc = lib.dv_comm_init(path, C.c_void_p(ident.data_ptr()), dist.get_rank(), dist.get_world_size(), C.byref(comm))
    good = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=device if on_dev else "cpu")
    dist.all_reduce(good, op=dist.ReduceOp.MIN)
    if int(good.cpu()) == 0:
        if rc == 0:
            lib.dv_comm_destroy(comm)
        return False
    _COMM = comm
    return True


def native_comm_active() -> bool:
    return _COMM is not None


def close_native_comm() -> None:
    global _COMM
    if _COMM is not None:
        from . import _lib
        _lib.load().dv_comm_destroy(_COMM)
        _COMM = None


def shard_slice(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous shard of the batch for `rank`; the remainder is spread over the first ranks."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def shard_inputs(tensors: Tuple[torch.Tensor, ...], rank: int, world: int) -> Tuple[torch.Tensor, ...]:
    """Slice globally generated inputs (noise drawn once for the whole batch, then sharded, so the result is
    bit-identical to a single-GPU run)."""
    sl = shard_slice(tensors[0].shape[0], rank, world)
    return tuple(t[sl].contiguous() for t in tensors)


def gather_latents(latent: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather the per-rank output latents [B_r, C, T] into [global_batch, C, T] (ragged shards are padded)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return latent
    world = dist.get_world_size()
    per = (global_batch + world - 1) // world
    buf = latent
    if latent.shape[0] != per:
        buf = latent.new_zeros((per,) + tuple(latent.shape[1:]))
        buf[: latent.shape[0]] = latent
    out = latent.new_empty((world * per,) + tuple(latent.shape[1:]))
    buf = buf.contiguous()
    if _COMM is not None and latent.is_cuda and latent.dtype == torch.float32:
        # in-stream ncclAllGather through the C ABI, on the stream the sampling loop was enqueued on
        from . import _lib
        with torch.cuda.device(latent.device):
            _lib.check(_lib.load().dv_comm_all_gather(_COMM, C.c_void_p(buf.data_ptr()), C.c_void_p(out.data_ptr()),
                                                      buf.numel(), C.c_void_p(torch.cuda.current_stream(latent.device).cuda_stream)))
    else:
        dist.all_gather_into_tensor(out, buf)
    if world * per == global_batch:
        return out
    parts = []
    for r in range(world):
        sl = shard_slice(global_batch, r, world)
        parts.append(out[r * per: r * per + (sl.stop - sl.start)])
    return torch.cat(parts, dim=0)
