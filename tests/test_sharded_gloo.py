"""world_size-2 gloo test of the post-sharded exchange logic (fancyrec_b200/sharded.py) on CPU.

The device kernels are replaced by an oracle-backed provider (tests only) so that what is exercised is
the host-side N>1 path: shard bounds, global indices, the candidate-list all-gather + merge, the
global-best reduction, the label gather and the count-pass reduction.  The merged statistics must equal
the single-process oracle on the unsharded problem, bit for bit.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ranking as oref
from oracle import synth


class OracleKernels:
    """CPU stand-in for fancyrec_b200.ops with the same call signatures (TEST ONLY)."""

    def __init__(self, scores_full):
        self.scores_full = scores_full       # [NB, NP] numpy, the whole problem

    def _local(self, index_base, n):
        return self.scores_full[:, index_base:index_base + n]

    def score_topk(self, brand_op, post_op, k, d=None, labels=None, index_base=0, workspace=None, dense=False):
        s = self._local(index_base, post_op.shape[0])
        idx = oref.topk_indices(s, k)
        sc = np.take_along_axis(s, idx, 1)
        pad = k - idx.shape[1]
        if pad > 0:
            idx = np.concatenate([idx, np.full((idx.shape[0], pad), -1 - index_base)], 1)
            sc = np.concatenate([sc, np.full((sc.shape[0], pad), -np.inf, np.float32)], 1)
        lab = labels.numpy()
        pos = s[lab, np.arange(len(lab))]
        gidx = np.where(idx >= 0, idx + index_base, -1).astype(np.int32)
        return dict(scores=torch.from_numpy(sc.astype(np.float32)), index=torch.from_numpy(gidx),
                    pos_score=torch.from_numpy(pos.astype(np.float32)), workspace=None)

    def label_stats(self, labels, pos_score, nb, index_base=0):
        lab, ps = labels.numpy(), pos_score.numpy()
        n_pos = np.bincount(lab, minlength=nb).astype(np.int32)
        bs = np.full(nb, -np.inf, np.float32)
        bi = np.full(nb, -1, np.int32)
        for b in range(nb):
            js = np.where(lab == b)[0]
            if len(js):
                j = js[np.lexsort((js, -ps[js].astype(np.float64)))[0]]
                bs[b], bi[b] = ps[j], j + index_base
        return torch.from_numpy(n_pos), torch.from_numpy(bs), torch.from_numpy(bi)

    def topk_merge(self, scores, index, k_out):
        s, i = oref.merge_topk(list(scores.numpy()), list(index.numpy()), k_out)
        return torch.from_numpy(s), torch.from_numpy(i.astype(np.int32))

    def rank_from_topk(self, topk_index, labels, index_base=0):
        idx, lab = topk_index.numpy(), labels.numpy()
        nb = idx.shape[0]
        mask = np.zeros(nb, dtype=np.uint64)
        first = np.full(nb, -1, np.int32)
        for b in range(nb):
            hits = np.array([i >= 0 and lab[i - index_base] == b for i in idx[b]])
            for r in np.where(hits[:64])[0]:
                mask[b] |= np.uint64(1) << np.uint64(r)
            if hits.any():
                first[b] = int(np.argmax(hits))
        return torch.from_numpy(mask.view(np.int64)), torch.from_numpy(first)

    def score_count(self, brand_op, post_op, thr_score, thr_index, d=None, index_base=0, out=None):
        s = self._local(index_base, post_op.shape[0])
        ts, ti = thr_score.numpy(), thr_index.numpy()
        j = np.arange(s.shape[1]) + index_base
        for b in range(s.shape[0]):
            if ti[b] >= 0:
                out[b] += int(((s[b] > ts[b]) | ((s[b] == ts[b]) & (j < ti[b]))).sum())
        return out


    def merge_gathered(self, gathered, nb, k_in, k_out):
        g = gathered.shape[0]
        rows = [gathered[r].contiguous() for r in range(g)]
        o = nb * k_in
        sc = torch.stack([r[:o].view(torch.float32).reshape(nb, k_in) for r in rows])
        ix = torch.stack([r[o:2 * o].reshape(nb, k_in) for r in rows])
        top_s, top_i = self.topk_merge(sc, ix, k_out)
        n_pos = torch.stack([r[2 * o:2 * o + nb] for r in rows]).sum(0).to(torch.int32)
        gs = torch.stack([r[2 * o + nb:2 * o + 2 * nb].view(torch.float32) for r in rows])
        gi = torch.stack([r[2 * o + 2 * nb:2 * o + 3 * nb] for r in rows])
        valid = gi >= 0                      # best positive over the shards under (score desc, index asc)
        s = torch.where(valid, gs, torch.full_like(gs, float("-inf")))
        top = s.max(dim=0).values
        cand = torch.where(valid & (s == top.unsqueeze(0)), gi, torch.full_like(gi, torch.iinfo(gi.dtype).max))
        idx = torch.where(~valid.any(dim=0), torch.full_like(cand[0], -1), cand.min(dim=0).values)
        return top_s, top_i, n_pos, top, idx

    def group_positives(self, labels, pos_score, n_pos):
        lab, ps, npos = labels.numpy(), pos_score.numpy(), n_pos.numpy()
        seg = np.concatenate([[0], np.cumsum(npos)]).astype(np.int64)
        srt = np.concatenate([np.sort(ps[lab == b]) for b in range(len(npos))] + [np.zeros(0, np.float32)])
        return torch.from_numpy(seg), torch.from_numpy(srt.astype(np.float32))

    def score_dense(self, brand_op, post_op, d=None, out=None):
        return out                                   # auc_rows reads the oracle's scores directly

    def auc_rows(self, scores, row0, labels, seg_ptr, pos_sorted, best_score, best_index, auc_num, before_first,
                 index_base=0):
        lab, seg, srt = labels.numpy(), seg_ptr.numpy(), pos_sorted.numpy()
        bs, bi = best_score.numpy(), best_index.numpy()
        s_loc = self._local(index_base, len(lab))
        j = np.arange(len(lab)) + index_base
        for r in range(scores.shape[0]):
            b = row0 + r
            pos = srt[seg[b]:seg[b + 1]]
            if len(pos) == 0:
                continue
            neg = s_loc[b][lab != b]
            auc_num[b] += int((len(pos) - np.searchsorted(pos, neg, side="right")).sum())   # positives > negative
            before_first[b] += int(((s_loc[b] > bs[b]) | ((s_loc[b] == bs[b]) & (j < bi[b]))).sum())

    def missing_thresholds(self, n_pos, first_in_list, best_index):
        missing = (first_in_list < 0) & (n_pos > 0)
        return torch.where(missing, best_index, torch.full_like(best_index, -1))

    def pack_rank_stats(self, n_pos, first_in_list, before_first, hit_mask, auc_num=None, all_valid=False):
        valid = torch.ones_like(before_first) if all_valid else ((first_in_list < 0) & (n_pos > 0)).to(torch.int64)
        rows = [n_pos.to(torch.int64), first_in_list.to(torch.int64), before_first, valid, hit_mask]
        if auc_num is not None:
            rows.append(auc_num)
        return torch.stack(rows)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, out_dir, want_auc=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fancyrec_b200 import ranking, sharded
    scores, lab = _problem()
    nb, n_posts = scores.shape
    lo, hi = sharded.shard_bounds(n_posts, world, rank)
    st = sharded.sharded_rank_statistics(torch.zeros(nb, 8), torch.zeros(hi - lo, 8),
                                         torch.from_numpy(lab[lo:hi].astype(np.int32)), 8, k, n_posts,
                                         kernels=OracleKernels(scores), want_auc=want_auc)
    stats = ranking.host_statistics(st, n_posts, want_auc=want_auc, kernels=OracleKernels(scores))
    res = ranking.aggregate(stats, n_posts, want_auc=want_auc)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), first_rank=stats["first_rank"], n_pos=stats["n_pos"],
             hits=stats["hits"], topk=st["topk_index"].numpy(), res=np.array([float(x) for x in res]),
             auc_num=stats.get("auc_num", np.zeros(0, np.int64)))
    dist.destroy_process_group()


def _problem():
    rs = np.random.RandomState(123)
    nb, n_posts = 9, 1201                       # ragged shards, heavy ties, one empty brand
    scores = (rs.randint(-40, 41, size=(nb, n_posts)) / 32.0).astype(np.float32)
    lab = synth.labels(5, n_posts, nb, empty_brands=(6,))
    scores[3, lab == 3] = -2.0                  # brand 3: every positive ranks last -> needs the count pass
    return scores, lab


def test_shard_bounds_cover_everything():
    from fancyrec_b200 import sharded
    for n in (0, 1, 7, 1000, 1201):
        for world in (1, 2, 3, 8):
            b = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))


def test_two_rank_exchange_matches_unsharded_oracle(tmp_path):
    world, k = 2, 64
    port = _free_port()
    mp.spawn(_worker, args=(world, port, k, str(tmp_path)), nprocs=world, join=True)
    scores, lab = _problem()
    ost = oref.rank_stats(scores, lab)
    want = oref.aggregate(ost, scores.shape[1])
    for rank in range(world):
        g = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert np.array_equal(g["n_pos"], ost["n_pos"])
        assert np.array_equal(g["first_rank"], ost["first_rank"])
        assert np.array_equal(g["hits"], ost["hits"])
        assert np.array_equal(g["topk"], oref.topk_indices(scores, k))
        got = tuple(g["res"])
        assert got[:2] == tuple(map(float, want[:2])) and got[3:] == tuple(map(float, want[3:]))
    # make sure the count-pass branch ran for at least one brand
    assert (ost["first_rank"] >= k).any()


def test_two_rank_exchange_with_auc_matches_unsharded_oracle(tmp_path):
    """want_auc: positives' scores ride in the exchange, every rank sweeps its own posts against the job-wide sorted
    positives, numerators and first-positive counts are summed -> the full reference 8-tuple, bit for bit."""
    world, k = 2, 64
    port = _free_port()
    mp.spawn(_worker, args=(world, port, k, str(tmp_path), True), nprocs=world, join=True)
    scores, lab = _problem()
    ost = oref.rank_stats(scores, lab)
    want = tuple(map(float, oref.rank_metrics_vec(scores, lab)))
    for rank in range(world):
        g = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert np.array_equal(g["auc_num"], ost["auc_num"])
        assert np.array_equal(g["first_rank"], ost["first_rank"])
        assert tuple(g["res"]) == want
