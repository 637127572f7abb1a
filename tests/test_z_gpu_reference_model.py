"""The CUDA path, through the C ABI, held directly to the Go-shaped model of tests/test_reference_model.py (a string-keyed Redis
dict, json.loads parsing): random clusters, micro-batches of 1-60 requests (the warp-per-body latency path with the exact engine
behind it), their responses, minute and day roll-overs; every decision, 429 payload value, pick and counter equal. The other GPU
tests compare the library with the C oracle; this one does not involve the oracle at all."""
import pytest

from test_reference_model import run_scenario


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [100, 101, 102, 103, 104, 105])
def test_cuda_path_against_the_go_shaped_model(seed, gwmod):
    def make(tables):
        g = gwmod.Gateway(0, 4096, 8 << 20)
        g.load_tables(tables)
        return g
    run_scenario(make, seed)
