"""Post-sharded evaluation across the GPUs of one box.

Posts are independent columns of the score matrix: rank r owns the contiguous range
[NP*r/G, NP*(r+1)/G) (global index = offset + local, so the (score desc, index asc) comparator is
shard-invariant) and the small brand operand is replicated.  Each rank runs the fused score + top-k
kernel on its shard; the only data-path exchange is ONE all-gather of the per-brand candidate lists
([NB, k] x (fp32 score, int32 global index), with the NB-length label statistics and the labels (4 B / post)
riding in the same buffer) followed by a merge with the same comparator, plus one NB-length all-reduce of the
count-pass result.  One process per GPU, torch.distributed (NCCL over
NVLink on the box, gloo in the CPU tests).

`kernels` is the provider of the device steps (default: fancyrec_b200.ops, i.e. libfrx_b200.so); the
CPU tests inject an oracle-backed provider to exercise the exchange logic without a GPU.
"""
import torch
import torch.distributed as dist

from . import ops as _ops


def shard_bounds(n_posts, world, rank):
    return n_posts * rank // world, n_posts * (rank + 1) // world


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_gather_stack(t, group=None):
    """[n, ...] per rank -> [G, n, ...] (rank-major), one all-gather."""
    world, _ = _world(group)
    flat = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(flat, t.contiguous(), group=group)
    return flat.view((world,) + tuple(t.shape))


def _mark(events, name, t):
    """Measurement hook: a CUDA event on the current stream (bench.py reads exchange / merge times from them)."""
    if events is not None and t.is_cuda:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(t.device))
        events[name] = ev


def all_sum(t, group=None):
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def sharded_rank_statistics(brand_op, post_op_local, labels_local, d, k, n_posts_total, group=None,
                            kernels=_ops, workspace=None, want_auc=False, events=None):
    """Device statistics of the WHOLE job from this rank's shard.  Returns the same dict layout as
    ranking.device_rank_statistics, identical on every rank.  `events`: dict that receives CUDA events around the
    exchange and the merge (measurement only).  want_auc: the positives' scores ride in the exchange
    block too (4 B / post); every rank then sweeps ITS posts against the job-wide sorted positives of each brand and the
    exact AUC numerators (and the counts before the first positive) are summed over the ranks."""
    world, rank = _world(group)
    lo, hi = shard_bounds(n_posts_total, world, rank)
    if not (post_op_local.shape[0] == hi - lo == labels_local.numel()):
        raise ValueError("rank %d of %d must hold posts [%d, %d) of %d (%d rows), got %d operand rows and %d labels"
                         % (rank, world, lo, hi, n_posts_total, hi - lo, post_op_local.shape[0], labels_local.numel()))
    nb = brand_op.shape[0]
    k = max(int(k), 64)                            # NDCG@50 reads 50 relevance bits per brand
    from . import ranking as _ranking
    fused_dense = want_auc and _ranking.dense_fits(nb, hi - lo)      # AUC fast path: this shard's scores, written on the way
    res = kernels.score_topk(brand_op, post_op_local, k, d=d, labels=labels_local, index_base=lo,
                             workspace=workspace, dense=fused_dense)
    n_pos_l, best_s_l, best_i_l = kernels.label_stats(labels_local, res["pos_score"], nb, lo)
    if world > 1:
        # ONE all-gather: every rank contributes [scores | index | n_pos | best score | best index | labels] as 32-bit
        # words; the lists are merged and the label statistics combined in place out of the gathered buffer.
        sizes = [shard_bounds(n_posts_total, world, r)[1] - shard_bounds(n_posts_total, world, r)[0] for r in range(world)]
        width = max(sizes)                           # ragged shards are padded to a common width for the collective
        pad = labels_local.new_zeros(width - labels_local.numel())
        parts = [res["scores"].reshape(-1).view(torch.int32), res["index"].reshape(-1), n_pos_l,
                 best_s_l.view(torch.int32), best_i_l, labels_local, pad]
        if want_auc:
            parts += [res["pos_score"].view(torch.int32), pad]
        block = torch.cat(parts)
        _mark(events, "exchange_begin", block)
        gathered = _all_gather_stack(block, group)
        _mark(events, "exchange_end", block)
        head = 2 * nb * k + 3 * nb
        top_s, top_i, n_pos, best_s, best_i = kernels.merge_gathered(gathered[:, :head], nb, k, k)
        _mark(events, "merge_end", block)
        if events is not None:
            events["exchange_bytes_sent"] = block.numel() * 4
            events["exchange_bytes_received"] = block.numel() * 4 * (world - 1)
        labels_all = torch.cat([gathered[r, head:head + sizes[r]] for r in range(world)])
        if want_auc:
            pos_all = torch.cat([gathered[r, head + width:head + width + sizes[r]] for r in range(world)]).view(torch.float32)
    else:
        top_s, top_i, n_pos, best_s, best_i = res["scores"], res["index"], n_pos_l, best_s_l, best_i_l
        labels_all = labels_local
        pos_all = res["pos_score"]
    hit_mask, first_in_list = kernels.rank_from_topk(top_i, labels_all, 0)
    before = torch.zeros(nb, dtype=torch.int64, device=post_op_local.device)
    out = dict(topk_scores=top_s, topk_index=top_i, n_pos=n_pos, best_score=best_s, best_index=best_i,
               hit_mask=hit_mask, first_in_list=first_in_list, before_first=before, workspace=res.get("workspace"))
    if want_auc:
        seg_ptr, pos_sorted = kernels.group_positives(labels_all, pos_all, n_pos)        # job-wide positives per brand
        auc_num = _ranking.auc_sweep(kernels, brand_op, post_op_local, d, labels_local, seg_ptr, pos_sorted, best_s,
                                     best_i, lo, before, dense=res.get("dense"))          # this shard's posts
        out["auc_num"] = all_sum(auc_num, group)
    else:
        # Count pass (rank of a first positive that fell outside the list), enqueued unconditionally: the kernel skips
        # every 128-brand tile without a missing row, so it returns at once in the common case -- no host round trip.
        thr_index = kernels.missing_thresholds(n_pos, first_in_list, best_i)
        kernels.score_count(brand_op, post_op_local, best_s, thr_index, d=d, index_base=lo, out=before)
    all_sum(before, group)
    return out
