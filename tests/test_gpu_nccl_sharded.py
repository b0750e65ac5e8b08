"""Two B200s, NCCL over NVLink: the post-sharded evaluation (fancyrec_b200/sharded.py, the path bench.py --gpus N runs)
returns on every rank exactly the single-GPU result for the concatenated posts -- top-k lists, label statistics,
first-positive ranks, exact AUC numerators and the final 8-tuple.  Skipped on a box with one GPU (bench.py's in-run
`sharded_check` is the N > 1 evidence there); run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl_sharded.py`.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, shapes, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from fancyrec_b200 import ranking, sharded
    for case, (nb, n, d, k, want_auc) in enumerate(shapes):
        g = torch.Generator(device=dev).manual_seed(900 + case)                 # same stream on every rank: same problem
        brand = torch.randn((nb, d), generator=g, device=dev)
        labels = (torch.randperm(n, generator=g, device=dev) % (nb + 1)).to(torch.int32)
        posts = torch.randn((n, d), generator=g, device=dev) + 0.2 * brand[labels.long() % nb]
        posts[::53] = posts[1]                                                   # exact ties across shard boundaries
        brand_op, post_op = ranking.to_operand(brand), ranking.to_operand(posts)
        lo, hi = sharded.shard_bounds(n, world, rank)
        st = sharded.sharded_rank_statistics(brand_op, post_op[lo:hi].contiguous(), labels[lo:hi].contiguous(), d, k, n,
                                             want_auc=want_auc)
        got = ranking.aggregate(ranking.host_statistics(st, n, want_auc), n, want_auc)
        ref = ranking.device_rank_statistics(brand_op, post_op, labels, d, k=k, want_auc=want_auc)
        want = ranking.aggregate(ranking.host_statistics(ref, n, want_auc), n, want_auc)
        for key in ("topk_index", "topk_scores", "n_pos", "best_index", "hit_mask", "first_in_list", "before_first") + \
                (("auc_num",) if want_auc else ()):
            assert torch.equal(st[key], ref[key]), (case, key)
        a, b = tuple(map(float, got)), tuple(map(float, want))
        assert all(x == y or (x != x and y != y) for x, y in zip(a, b)), (case, a, b)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL over NVLink)")
def test_nccl_sharded_evaluation_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    shapes = [(37, 5001, 96, 64, True),            # ragged shards, k >= posts per brand, AUC
              (300, 400000, 256, 100, False),      # sample pass + histogram thresholds active
              (1000, 600000, 128, 100, True)]
    mp.spawn(_worker, args=(2, port, shapes, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok0")) and os.path.exists(os.path.join(str(tmp_path), "ok1"))
