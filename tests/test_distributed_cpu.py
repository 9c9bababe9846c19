"""World-size-2 gloo test of the N>1 path: contiguous sharding + the single all-gather of scores."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from t2v_metrics_b200.parallel import gather_scores, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n, dtype=torch.float32) * 0.5 + 1.0          # the scores a single process would produce
    s, e, _ = shard_bounds(n, world, rank)
    got = gather_scores(full[s:e].clone(), n)
    q.put((rank, torch.equal(got, full)))
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    ctx = mp.get_context("spawn")
    for n in (10, 7, 1):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
        [p.start() for p in procs]
        [p.join(60) for p in procs]
        res = sorted(q.get(timeout=5) for _ in range(2))
        assert res == [(0, True), (1, True)], (n, res)


class _FakeModel:
    """score(image i, text j) = i + j / 100: lets the test see exactly which pairs a rank was asked for."""
    def __init__(self):
        self.seen = []

    def forward(self, images, texts, **kw):
        self.seen += images
        return torch.tensor([float(i[3:]) + float(t[1:]) / 100.0 for i, t in zip(images, texts)])


def _score_worker(rank, world, port, m, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from t2v_metrics_b200.score import Score

    class S(Score):
        def prepare_scoremodel(self, model, device, cache_dir, **kw):
            return _FakeModel()

        def list_all_models(self):
            return ["fake"]

    sc = S("fake", device="cpu")
    sc.shard_over_images = True
    sc.max_pairs = 4
    images, texts = [f"img{i}" for i in range(m)], [f"t{j}" for j in range(n)]
    got = sc(images=images, texts=texts)
    want = torch.tensor([[i + j / 100.0 for j in range(n)] for i in range(m)])
    mine = sorted(set(sc.model.seen))
    s, e, _ = shard_bounds(m, world, rank)
    q.put((rank, torch.allclose(got, want), mine == images[s:e]))
    dist.destroy_process_group()


def test_score_forward_shards_images_over_ranks():
    """Score.forward under torchrun: each rank scores only its contiguous share of the images (all texts), every rank returns the full [m, n]."""
    ctx = mp.get_context("spawn")
    for m, n in ((5, 3), (1, 4), (2, 1)):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_score_worker, args=(r, 2, port, m, n, q)) for r in range(2)]
        [p.start() for p in procs]
        [p.join(120) for p in procs]
        res = sorted(q.get(timeout=5) for _ in range(2))
        assert res == [(0, True, True), (1, True, True)], ((m, n), res)
