"""Data-parallel protocol on ONE GPU (world size 1 over NCCL): every exchange tools/dp_check.py knows -- the
peer-memory path (IPC arena, scatter epilogue, flags, sharded update), the NCCL reduce-scatter / all-gather path and
the chunked all-reduce with its row-block entry points (dbn_rbm_first_hidden / dbn_rbm_delta_rows /
dbn_rbm_apply_delta with h0 > 0 and global row offsets) -- must train the same weights as the single-GPU epoch
runner.  The N > 1 runs of the same script are kept under profiles/ (r02_dp_check_n*.log); `world` virtual ranks on
one GPU are tests/test_gpu_dp_peer.py; the host-side sharding logic is covered on CPU by test_dist_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_dp_runner_matches_single_gpu_runner_at_world_size_one():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("dp_check")]
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    checked = [l for l in lines if "replicas" in l]
    assert len(checked) == 10 and all(l.endswith("OK") for l in checked), "\n".join(lines)
    assert any(" peer " in l and "H=4096" in l for l in checked), "the peer-memory path did not run at config-5 width"
