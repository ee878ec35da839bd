"""Sharding of windows / channels across the GPUs of one box.

The path has no exchange step: every window (and every candidate in it) is independent
(FDR_impl::transform and demodulate read only their own PDU).  Rank r of R takes the
contiguous slice [r*N/R, (r+1)*N/R) of the flattened (channel, window) index and, for a
sliding-window stream, the contiguous span of samples that slice covers.  Results stay per
rank; an optional gather of the (small) candidate lists is the only collective.
"""


def shard_range(n, rank, world):
    """contiguous slice of n units for `rank` of `world`: sizes differ by at most one"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def stream_span(lo, hi, stride, fl):
    """sample span [start, stop) of a sliding-window stream that windows lo..hi-1 read"""
    if hi <= lo:
        return lo * stride, lo * stride
    return lo * stride, (hi - 1) * stride + fl


def gather_counts(local_count, dist=None):
    """all ranks' unit counts (torch.distributed all_gather of one integer); [local] without a group"""
    if dist is None or not dist.is_initialized():
        return [int(local_count)]
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(v.item()) for v in out]
