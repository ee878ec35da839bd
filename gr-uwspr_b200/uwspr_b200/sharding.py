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


def gather_floats(value, dist=None):
    """all ranks' values of one float (all_gather); [value] without a group"""
    if dist is None or not dist.is_initialized():
        return [float(value)]
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(v.item()) for v in out]


def balanced_counts(total, rates, lo=1, hi=None):
    """Split `total` independent units over ranks in proportion to their measured rates (units/s), each
    share within [lo, hi]: the host links of one box are not equally fast when every GPU copies at once
    (shared PCIe uplinks, a socket hop), and a fixed equal split makes every rank wait for the slowest
    link.  Deterministic (largest-remainder rounding), so every rank computes the same answer from the
    same gathered rates."""
    n = len(rates)
    hi = total if hi is None else hi
    if n == 0 or total < n * lo or total > n * hi:
        raise ValueError("no split of %d units over %d ranks within [%d, %d]" % (total, n, lo, hi))
    rates = [max(float(r), 1e-12) for r in rates]
    counts = [0] * n
    free = list(range(n))
    rest = total
    # proportional shares; ranks whose share leaves [lo, hi] are pinned to the bound and the others
    # share what is left (repeat until no share leaves the range)
    while free:
        s = sum(rates[i] for i in free)
        want = {i: rest * rates[i] / s for i in free}
        bad = [i for i in free if want[i] > hi or want[i] < lo]
        if bad:
            # pin the worst offender first, then recompute
            i = max(bad, key=lambda j: max(want[j] - hi, lo - want[j]))
            counts[i] = hi if want[i] > hi else lo
            rest -= counts[i]
            free.remove(i)
            continue
        fl = {i: int(want[i]) for i in free}
        left = rest - sum(fl.values())
        order = sorted(free, key=lambda i: (-(want[i] - fl[i]), i))
        for i in free:
            counts[i] = fl[i]
        for i in order[:left]:
            counts[i] += 1
        break
    assert sum(counts) == total and all(lo <= c <= hi for c in counts)
    return counts
