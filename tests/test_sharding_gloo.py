"""world_size-2 gloo run of the sharding logic bench.py uses under torchrun (CPU only)."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "gr-uwspr_b200"))
import torch, torch.distributed as dist
from uwspr_b200.sharding import shard_range, gather_counts, gather_floats, balanced_counts
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(1001, r, w)
counts = gather_counts(hi - lo, dist)
assert sum(counts) == 1001 and len(counts) == w, counts
t = torch.tensor([float(hi - lo)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
# link-rate balanced split, as bench.py does for its host-fed arm: same answer on every rank
rates = gather_floats(100.0 if r == 0 else 300.0, dist)
bal = balanced_counts(2000, rates, lo=100, hi=1600)
assert bal == [500, 1500] and gather_counts(bal[r], dist) == bal, bal
dist.barrier()
if r == 0:
    print("OK", counts, int(t.item()))
dist.destroy_process_group()
'''


def test_two_rank_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK [500, 501] 501" in out.stdout or "OK [501, 500] 501" in out.stdout, out.stdout


def test_balanced_counts_properties():
    import sys as _s
    _s.path.insert(0, os.path.join(ROOT, "gr-uwspr_b200"))
    from uwspr_b200.sharding import balanced_counts
    assert balanced_counts(80000, [1.0] * 8, lo=1000, hi=15000) == [10000] * 8
    c = balanced_counts(80000, [34.6] * 4 + [28.5] + [21.0] * 3, lo=1000, hi=15000)
    assert sum(c) == 80000 and c[0] == c[1] == c[2] == c[3] > c[4] > c[5] and max(c) <= 15000
    assert abs(c[0] / c[5] - 34.6 / 21.0) < 0.01
    c = balanced_counts(80000, [100.0] + [1.0] * 7, lo=1000, hi=15000)   # one link far faster: capped
    assert c[0] == 15000 and sum(c) == 80000 and min(c) >= 1000
    assert balanced_counts(7, [1, 1, 1], lo=1) in ([3, 2, 2],)
    import pytest
    with pytest.raises(ValueError):
        balanced_counts(10, [1, 1], lo=1, hi=4)
