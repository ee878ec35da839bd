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
from uwspr_b200.sharding import shard_range, gather_counts
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(1001, r, w)
counts = gather_counts(hi - lo, dist)
assert sum(counts) == 1001 and len(counts) == w, counts
t = torch.tensor([float(hi - lo)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
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
