"""Multi-GPU plumbing: one process per GPU, frames/streams partitioned by rank, no data-path
collective.  The only cross-rank traffic is the sum of the 8-entry BER/EVM counter vector
(include/ofdm_b200.h: ofdm_counters), reduced with torch.distributed (NCCL on GPUs, gloo in the
CPU tests)."""


def stream_owner(stream_id, world_size):
    """Streams are dealt round-robin; all frames of a stream stay on one GPU."""
    return stream_id % world_size


def my_streams(n_streams, rank, world_size):
    return list(range(rank, n_streams, world_size))


def frame_block(n_frames, rank, world_size):
    """Contiguous [lo, hi) block of a flat frame list for this rank (sizes differ by at most one)."""
    base, rem = divmod(n_frames, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_counters(cnt, group=None):
    """In-place sum of the counter vector over all ranks; no-op when not running distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return cnt


def summarize(cnt):
    """dict(ber, evm_rms, locked_fraction) from a (reduced) counter vector."""
    c = [float(v) for v in cnt]
    return {"bit_errors": c[0], "bits": c[1], "ber": c[0] / c[1] if c[1] else 0.0,
            "evm_rms": (c[2] / c[3]) ** 0.5 if c[3] else 0.0,
            "frames_locked": c[4], "frames": c[5], "locked_fraction": c[4] / c[5] if c[5] else 0.0}
