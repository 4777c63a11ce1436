"""Frame-parallel plumbing: one process per GPU, no data-path collective.

The reference shards video frames round-robin over devices inside one process
(nunif/utils/video.py:1695 `devices[rr % n]`, nunif/models/data_parallel.py:53-62) and replicates the
weights with torch.nn.parallel.replicate (data_parallel.py:16,58).  Here every rank owns one GPU and a
full weight replica: rank 0's packed blob is broadcast once (NCCL over NVLink on GPUs, gloo in the CPU
tests), frames are assigned round-robin, and the only other collective is the timing reduction.
"""
import torch
import torch.distributed as dist


def shard_frames(n_frames, rank, world):
    """Frame indices owned by `rank`: rank, rank+world, ... (video.py:1695 round-robin)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} out of range for world size {world}")
    return list(range(rank, n_frames, world))


def owner_of(frame_index, world):
    return frame_index % world


def broadcast_blob(blob, src=0):
    """Broadcast a flat uint8 tensor (the packed weight blob) from `src`; no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return blob


def model_blob_tensor(model):
    """A torch uint8 view (no copy) of a B200I2IModel's packed device weights."""
    ptr, nbytes = model.weight_blob()

    class _Blob:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Blob(), device=model.device)


def broadcast_model_weights(model, src=0):
    return broadcast_blob(model_blob_tensor(model), src=src)


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over all ranks (timings are reported as the slowest rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def gather_frame_order(local_indices, world):
    """All ranks' frame indices, to check that every frame is produced exactly once."""
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return [list(local_indices)]
    out = [None] * world
    dist.all_gather_object(out, list(local_indices))
    return out
