"""world_size-2 `gloo` tests (CPU) of the multi-GPU plumbing: template sharding by tmpl_idx and the
top-K all-gather/merge.  The per-rank top-K lists come from the oracle here (no GPU in this container);
on a GPU box the same functions run over NCCL (bench.py)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from openfdcm_b200 import distributed as fd
from openfdcm_b200._lib import MATCH_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import fdcm_oracle as orc
        from tests.util import plant_instances, synth_scene, synth_templates
        tmpls = synth_templates(10, 12, 320, seed=7)
        scene = plant_instances(synth_scene(320, 240, 60, seed=6), tmpls, 320, 240, seed=8)
        fm = orc.Dt3Cpu(scene, 8, 5.0, 1.5, orc.L2, nthreads=2)
        mine, base = fd.shard_templates(tmpls, rank, world)
        raw = fm.search(mine, scene, 3, 3, batch=10, nthreads=2)
        raw["tmpl_idx"] += base                                      # global template indices
        pen = orc.penalize(1, 1.5, raw, orc.template_lengths(tmpls))
        local = pen[np.argsort(pen["score"], kind="stable")[:5]]
        merged = fd.allgather_topk(local, 5)
        full = fm.search(tmpls, scene, 3, 3, batch=10, nthreads=2)
        full = orc.penalize(1, 1.5, full, orc.template_lengths(tmpls))
        want = full[np.argsort(full["score"], kind="stable")[:5]]
        ok = bool(np.array_equal(merged["score"], want["score"]) and np.array_equal(merged["tmpl_idx"], want["tmpl_idx"])
                  and np.array_equal(merged["transform"], want["transform"]))
        # a rank with fewer than k matches must not poison the merge
        short = fd.allgather_topk(local[: (1 if rank == 0 else 5)], 5)
        ok = ok and len(short) == 5 and bool(np.all(np.isfinite(short["score"])))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_topk_allgather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_helpers():
    for n, w in ((10, 3), (5000, 8), (7, 8), (0, 2)):
        spans = [fd.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    assert fd.shard_scenes(10, 1, 4) == [1, 5, 9]
    a = np.zeros(3, MATCH_DTYPE); a["score"] = [1, 3, 5]; a["tmpl_idx"] = [0, 1, 2]
    b = np.zeros(2, MATCH_DTYPE); b["score"] = [2, 3]; b["tmpl_idx"] = [7, 8]
    m = fd.merge_topk_records([a, b], 4)
    assert m["score"].tolist() == [1, 2, 3, 3] and m["tmpl_idx"].tolist() == [0, 7, 1, 8]
