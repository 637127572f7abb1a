"""N>1 path: world_size-2 runs of the tenant partition + shared-quota fold.
CPU: gloo, engine = oracle behind the same protocol (host logic). GPU: nccl, engine = CUDA library (needs 2 GPUs)."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(backend, tmp_path, world=2):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi_rank_worker.py"), str(r), str(world), backend,
                               str(tmp_path), str(port)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{outs[r][-3000:]}"
        assert os.path.exists(os.path.join(tmp_path, f"rank{r}.ok"))


def test_partition_is_a_function_of_the_namespace():
    from arks_b200.sharding import fnv1a64, shard_of
    assert fnv1a64(b"") == 0xCBF29CE484222325
    assert fnv1a64(b"a") == 0xAF63DC4C8601EC8C
    for world in (1, 2, 4, 8):
        owners = [shard_of("tenant-%05d" % t, world) for t in range(2000)]
        assert set(owners) == set(range(world))
        assert owners == [shard_of("tenant-%05d" % t, world) for t in range(2000)]


def test_two_ranks_gloo_cpu(tmp_path):
    launch("gloo", tmp_path)


@pytest.mark.gpu
def test_two_ranks_nccl_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import __graft_entry__ as ge
    ge.build()
    launch("nccl", tmp_path)
