"""Multi-GPU parity + broadcast-vs-rebuild measurement (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/multigpu_parity.py out.json
  * the bench's 5000 templates sharded over the ranks through the C ABI communicator (fdcm_comm_search_topk: NCCL
    all-gather of the device top-K + merge kernel): merged top-10 == single-GPU top-10 == oracle top-10 (rank 0 checks);
  * a shard that is empty on some ranks still completes the collective;
  * fdcm_comm_rebuild_broadcast: planes identical to a local rebuild; time of "every rank rebuilds" vs "root builds +
    ncclBroadcast" at 1080p (2880^2 x 30 = 995 MB) and 4K (5760^2 x 30 = 3.98 GB) maps.
Rank 0 writes one JSON document."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
import openfdcm_b200 as fdcm
from openfdcm_b200 import distributed as fd
from tests.util import synth_scene


def timed(fn, n, dev):
    torch.cuda.synchronize(dev)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    t = torch.tensor([(time.perf_counter() - t0) / n * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/multigpu_parity.json"
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = fd.Communicator.from_torch(local)
    res = {"world": world}

    scene, tmpls = bench.make_workload(0)
    searcher, optimizer, penalty = fdcm.DefaultSearch(bench.MAX_T, bench.MAX_S), fdcm.BatchOptimize(bench.BATCH), fdcm.ExponentialPenalty(bench.TAU)
    params = fdcm.Dt3CudaParameters(bench.DEPTH, bench.COEFF, bench.PADDING, fdcm.distance.L2, device=local)
    fm = fdcm.build_cuda_featuremap(scene, params)
    b, e = comm.shard(len(tmpls))
    shard = fdcm.TemplateSet(tmpls[b:e], device=local)
    merged = comm.search_topk(fm, shard, None, searcher, optimizer, penalty, bench.TOP_K, b)
    ok = True
    if rank == 0:
        single = fdcm.search_topk(fm, tmpls, None, searcher, optimizer, penalty, bench.TOP_K)
        from oracle import fdcm_oracle as orc
        c = orc.Dt3Cpu(scene, bench.DEPTH, bench.COEFF, bench.PADDING, orc.L2)
        pen = orc.penalize(1, bench.TAU, c.search(tmpls, scene, bench.MAX_T, bench.MAX_S, batch=bench.BATCH), orc.template_lengths(tmpls))
        want = pen[np.lexsort((np.arange(len(pen)), pen["score"]))[:bench.TOP_K]]
        res["sharded_top10_equals_single_gpu"] = bool(np.array_equal(merged, single))
        res["sharded_top10_equals_oracle"] = bool(np.array_equal(merged, want))
        res["top10"] = [[int(r["tmpl_idx"]), float(r["score"])] for r in merged]
        ok = res["sharded_top10_equals_single_gpu"] and res["sharded_top10_equals_oracle"]
    # every rank got the same list
    mine = torch.from_numpy(merged.view(np.uint8).copy()).to(dev)
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([int(torch.equal(mine, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    res["all_ranks_same_list"] = bool(same.item())

    # fewer templates than ranks: some shards are empty, the collective still completes
    few = tmpls[:1]
    b1, e1 = comm.shard(len(few))
    tiny = comm.search_topk(fm, fdcm.TemplateSet(few[b1:e1], device=local), None, searcher, optimizer, penalty, bench.TOP_K, b1)
    if rank == 0:
        alone = fdcm.search_topk(fm, few, None, searcher, optimizer, penalty, bench.TOP_K)
        res["empty_shard_ok"] = bool(np.array_equal(tiny, alone))
        ok = ok and res["empty_shard_ok"]

    # broadcast vs rebuild
    res["broadcast_vs_rebuild"] = {}
    for name, (w, h, n) in (("1080p", (1920, 1080, 2000)), ("4k", (3840, 2160, 5000))):
        sc = synth_scene(w, h, n, seed=2000)
        m2 = fdcm.build_cuda_featuremap(sc, params)
        want_planes = [m2.plane(d) for d in (0, 15, 29)]
        t_rebuild = timed(lambda: m2.rebuild(sc), 5, dev)
        comm.rebuild_broadcast(m2, sc, root=0)
        planes_ok = all(np.array_equal(m2.plane(d), p) for d, p in zip((0, 15, 29), want_planes))
        t_bcast = timed(lambda: comm.rebuild_broadcast(m2, sc, root=0), 5, dev)
        flag = torch.tensor([int(planes_ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        nbytes = m2.device_ptr()[1]
        res["broadcast_vs_rebuild"][name] = {"map_bytes": nbytes, "rebuild_everywhere_ms": t_rebuild, "root_build_plus_ncclBroadcast_ms": t_bcast,
                                             "planes_identical_on_all_ranks": bool(flag.item()),
                                             "faster": "rebuild" if t_rebuild <= t_bcast else "broadcast"}
        ok = ok and bool(flag.item())
        del m2
    res["ok"] = bool(ok and res["all_ranks_same_list"])
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res))
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
