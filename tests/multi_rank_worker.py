"""Worker for the world_size-2 tests of the N>1 path (tenant partitioning + shared-quota delta fold).

backend "gloo": CPU, the engine is the oracle wrapped with the same take/apply protocol (host logic under test).
backend "nccl": one GPU per rank, the engine is the CUDA library with quota sharing enabled.
Every rank checks itself against a serial oracle that sees ALL traffic, and writes "ok" to its result file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

NOW = 1_700_000_000


class OracleEngine:
    """The oracle behind the delta-exchange protocol: delta = usage now - usage right after the last fold."""

    def __init__(self, tables):
        import orklib
        self.o = orklib.Oracle(tables)
        self.tables = tables
        self.base = self.o.snapshot_quota()

    def handle_request_body(self, b):
        return self.o.request_batch(b)

    def handle_response_body(self, b):
        return self.o.response_batch(b)

    def snapshot_quota(self):
        return self.o.snapshot_quota()

    def snapshot_rate(self, now):
        return self.o.snapshot_rate(now)

    def take_quota_delta(self):
        cur = self.o.snapshot_quota()
        d = cur - self.base
        self.base = cur
        return d

    def apply_quota_delta(self, remote):
        for q in np.nonzero(np.any(remote != 0, axis=1))[0]:
            self.o.incr_quota_usage(int(q), remote[q])
        self.base = self.o.snapshot_quota()


def objects(n_tenants):
    from arks_b200.tables import simple_endpoint, simple_quota, simple_token
    toks, quotas, eps = [], [], []
    for t in range(n_tenants):
        ns = "tenant-%03d" % t
        toks.append(simple_token("user", ns, "tk-%03d" % t, "m", [("rpm", 7 + t % 5), ("tpm", 10**7)], "q"))
        quotas.append(simple_quota("q", ns, [("prompt", 10**6), ("total", 3 * 10**6)]))
        eps.append(simple_endpoint("m", ns))
    # one hot tenant replicated on every GPU: its quota is the shared one
    toks.append(simple_token("hot-user", "hot", "tk-hot", "m", [("tpm", 10**9)], "hotq"))
    quotas.append(simple_quota("hotq", "hot", [("prompt", 5000), ("response", 10**7), ("total", 10**7)]))
    eps.append(simple_endpoint("m", "hot"))
    return toks, quotas, eps


def traffic_for(tok_strings, seed, n):
    rng = np.random.default_rng(seed)
    who = [tok_strings[int(i)] for i in rng.integers(0, len(tok_strings), n)]
    return [b'{"model":"m"}'] * n, who


def run(rank, world, backend, outdir, port):
    import torch
    import torch.distributed as dist
    from arks_b200 import abi
    from arks_b200.abi import RequestBatch, ResponseBatch
    from arks_b200.sharding import QuotaDeltaExchange, partition_objects, shard_of
    from arks_b200.tables import Tables
    import orklib

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    toks, quotas, eps = objects(24)
    mine = partition_objects(toks, quotas, eps, rank, world, replicate_namespaces=("hot",))
    t_local, t_global = Tables(*mine), Tables(toks, quotas, eps)
    # partition sanity: every non-replicated namespace has exactly one owner
    owned = {t["metadata"]["namespace"] for t in mine[0]} - {"hot"}
    assert all(shard_of(ns, world) == rank for ns in owned)
    if backend == "nccl":
        from arks_b200.gateway import Gateway
        torch.cuda.set_device(rank)
        eng = Gateway(rank, 4096, 8 << 20, share_quota=True)
        eng.load_tables(t_local)
        dev = torch.device("cuda", rank)
    else:
        eng, dev = OracleEngine(t_local), None
    ex = QuotaDeltaExchange(eng, t_local.n_quotas, [t_local.n_quotas - 1], dev)  # the hot tenant's quota is the shared one
    glob = orklib.Oracle(t_global)  # serial oracle over everybody's traffic (same on every rank)
    hot_local, hot_global = t_local.n_quotas - 1, t_global.n_quotas - 1
    local_tok = [t["spec"]["token"].encode() for t in mine[0]]
    for epoch in range(3):
        now = NOW + 7 * epoch
        outs = []
        for r in range(world):  # every rank replays every rank's traffic into the global oracle, in rank order
            r_objs = partition_objects(toks, quotas, eps, r, world, replicate_namespaces=("hot",))
            r_tok = [t["spec"]["token"].encode() for t in r_objs[0]]
            bodies, who = traffic_for(r_tok, 100 * epoch + r, 300)
            g_req = RequestBatch.from_lists(bodies, who, now)
            g_res = glob.request_batch(g_req)
            ok = np.nonzero(g_res.reason == 0)[0]
            rng = np.random.default_rng(1000 * epoch + r)
            usage = rng.integers(1, 40, (len(ok), 2))
            rb = [b'{"model":"m","usage":{"prompt_tokens":%d,"completion_tokens":%d,"total_tokens":%d}}' % (p, c, p + c)
                  for p, c in usage]
            glob.response_batch(ResponseBatch.from_lists(rb, g_res.qos[ok], [abi.RESP_END_OF_STREAM] * len(ok), now + 1))
            if r == rank:
                outs = (bodies, who, g_res, ok, rb)
        bodies, who, g_res, ok, rb = outs
        # my own traffic through my engine
        l_res = eng.handle_request_body(RequestBatch.from_lists(bodies, who, now))
        # single-owner tenants: decisions identical to the global serial oracle (their keys live only here)
        single = np.array([w != b"tk-hot" for w in who])
        assert np.array_equal(l_res.reason[single], g_res.reason[single]), "single-owner decisions differ"
        l_ok = np.nonzero(l_res.reason == 0)[0]
        # responses for the requests the GLOBAL order admitted and that I admitted too (hot tenant may lag: staleness)
        both = np.intersect1d(ok, l_ok)
        sel = [int(np.nonzero(ok == i)[0][0]) for i in both]
        eng.handle_response_body(ResponseBatch.from_lists([rb[k] for k in sel], l_res.qos[both],
                                                          [abi.RESP_END_OF_STREAM] * len(both), now + 1))
        ex.fold()
        dist.barrier()
    # after the last fold every replica of the shared quota holds the sum of all ranks' increments
    mine_hot = eng.snapshot_quota()[hot_local]
    gathered = [torch.zeros(3, dtype=torch.int64, device=dev or "cpu") for _ in range(world)]
    dist.all_gather(gathered, torch.tensor(mine_hot, dtype=torch.int64, device=dev or "cpu"))
    assert all(torch.equal(g.cpu(), gathered[0].cpu()) for g in gathered), f"replicas disagree: {gathered}"
    assert mine_hot.sum() > 0
    # single-owner quotas equal the global oracle's exactly
    gq = glob.snapshot_quota()
    lq = eng.snapshot_quota()
    names = lambda objs: [(q["metadata"]["namespace"], q["metadata"]["name"]) for q in objs]
    gidx = {k: i for i, k in enumerate(names(quotas))}
    for i, k in enumerate(names(mine[1])):
        if k[0] != "hot":
            assert np.array_equal(lq[i], gq[gidx[k]]), (k, lq[i], gq[gidx[k]])
    open(os.path.join(outdir, f"rank{rank}.ok"), "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]))
