"""Pipelined evaluation of a stream of batches on one GPU (or one shard of a post-sharded job).

`evaluator.test_post_ranking` is one synchronous pass: finalise the posts (HBM-bound), contract them against the brands
with the fused top-k epilogue (tensor-bound), reduce the rank statistics, copy 40 KB to the host, aggregate in float64
-- and the GPU idles while the host copies and aggregates (0.6 ms of an 8 ms step at BASELINE.json's config 2).  When
evaluations come as a stream -- validation every epoch (trainer.py:282-288), a sweep over checkpoints or test splits, the
chunks of a post set larger than HBM -- this module keeps the device busy:

  * main stream   post finalisation, brand side, fused score + top-k, rank statistics, pack, ASYNCHRONOUS copy of the
                  packed block into pinned host memory of batch t;
  * host          float64 aggregation (evaluator.py:129-143) of batch t-1 while the device works on batch t.

`overlap=True` additionally moves the finalisation of batch t+1 to a side stream so that it runs UNDER the contraction of
batch t: the contraction CTA (256 threads, 30 720 registers) leaves room on every SM for two finalise blocks, the
side-stream launch is bounded to that many blocks and asks for the same shared-memory carve-out (a block only joins an SM
whose shared-memory / L1 split matches its own).  Measured on B200 (tools/gpu_overlap_probe.py): the two
kernels do overlap, and each runs 55-60 % slower while they do (they contend for L2 / HBM): 6.8 ms instead of 7.14 ms for the
pair, 1 % on the pipelined step, and an isolated evaluation gets slower.  It is therefore off by default.

Results are the same 8-tuples the synchronous call returns, bit for bit (tests/test_gpu_auc_pipeline.py).
"""
import os

import numpy as np
import torch

from . import ops, ranking, sharded


class EvalPipeline:
    def __init__(self, device, nb, n_posts_local, dv, dt=0, k=ranking.MIN_TOPK, n_posts_total=None, group=None,
                 want_auc=False, overlap=False, depth=2, visual_norm=True, text_norm=True):
        self.dev = torch.device(device)
        self.nb, self.n_local, self.dv, self.dt, self.k = nb, n_posts_local, dv, dt, k
        self.d = dv + dt
        self.n_total = n_posts_local if n_posts_total is None else n_posts_total
        self.group, self.want_auc, self.depth = group, want_auc, max(2, int(depth))
        self.visual_norm, self.text_norm = visual_norm, text_norm
        self.fin_stream = torch.cuda.Stream(self.dev) if overlap else None
        self.main_stream = torch.cuda.Stream(self.dev)
        self.side_blocks_per_sm = int(os.environ.get("FRX_SIDE_BLOCKS", "2"))   # finalise blocks per SM next to a contraction CTA
        ld = ops.round_up(self.d, 64)
        rows = 6 if want_auc else 5
        self.post_op = [torch.empty((n_posts_local, ld), dtype=torch.bfloat16, device=self.dev) for _ in range(self.depth)]
        self.host = [torch.empty((rows, nb), dtype=torch.int64, pin_memory=True) for _ in range(self.depth)]
        self.fin_done = [torch.cuda.Event() for _ in range(self.depth)]
        self.op_free = [None] * self.depth          # recorded once the contraction that read post_op[slot] is enqueued
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self.pending = [None] * self.depth          # ticket whose packed block sits (or will sit) in host[slot]
        self.results = {}
        self.workspace = None
        self.ticket = 0
        self.last_stats = None
        self.last_brand_op = None

    # -----------------------------------------------------------------------------------------
    def submit(self, w, e, visual, text, labels_i32):
        """Enqueue one evaluation: aspect tables w [NB+1, A] / e [A, D], post rows visual [NP, Dv] (+ text [NP, Dt]) and
        labels [NP] int32, all on the device.  Returns a ticket for result(); never blocks on the device unless the
        slot's previous result has not been collected yet (then it is collected first)."""
        slot = self.ticket % self.depth
        if self.pending[slot] is not None:
            self._collect(slot)
        caller = torch.cuda.current_stream(self.dev)
        post_op = self.post_op[slot]
        fin, main = self.fin_stream, self.main_stream
        # The evaluation runs on the pipeline's OWN streams; they wait for what the caller has enqueued so far (the producer
        # of the inputs), not for the evaluations submitted before this one.
        inputs_ready = torch.cuda.Event()
        inputs_ready.record(caller)
        main.wait_event(inputs_ready)
        for t in (w, e, visual, text, labels_i32):
            if t is not None:
                t.record_stream(main)
        if fin is not None:
            fin.wait_event(inputs_ready)
            if self.op_free[slot] is not None:
                fin.wait_event(self.op_free[slot])         # the contraction that last read this operand buffer is over
            with torch.cuda.stream(fin):
                # bounded launch: as many finalise blocks per SM as fit next to a resident contraction CTA
                ops.finalize_posts(visual, text, visual_norm=self.visual_norm, text_norm=self.text_norm and text is not None,
                                   final_norm=True, out_bf16=post_op, blocks_per_sm=self.side_blocks_per_sm)
                self.fin_done[slot].record(fin)
            for t in (visual, text):
                if t is not None:
                    t.record_stream(fin)
        with torch.cuda.stream(main):
            if fin is None:
                if self.op_free[slot] is not None:
                    main.wait_event(self.op_free[slot])
                ops.finalize_posts(visual, text, visual_norm=self.visual_norm, text_norm=self.text_norm and text is not None,
                                   final_norm=True, out_bf16=post_op)
            brand = ops.brand_embed(w, e, nb=self.nb)
            brand_op = ops.finalize_posts(brand, final_norm=True)[1]
            if fin is not None:
                main.wait_event(self.fin_done[slot])
            st = sharded.sharded_rank_statistics(brand_op, post_op, labels_i32, self.d, self.k, self.n_total, group=self.group,
                                                 workspace=self.workspace, want_auc=self.want_auc)
            self.workspace = st["workspace"]
            packed = ranking.pack_statistics(st, self.want_auc)
            self.host[slot].copy_(packed, non_blocking=True)
            self.done[slot].record(main)
            free = torch.cuda.Event()
            free.record(main)
        self.op_free[slot] = free
        self.last_stats, self.last_brand_op = st, brand_op
        ticket = self.ticket
        self.pending[slot] = ticket
        self.ticket += 1
        return ticket

    def _collect(self, slot):
        ticket = self.pending[slot]
        self.done[slot].synchronize()
        stats = ranking.unpack_statistics(self.host[slot].numpy(), self.n_total, self.want_auc)
        self.results[ticket] = ranking.aggregate(stats, self.n_total, self.want_auc)
        self.pending[slot] = None

    def result(self, ticket):
        """The 8-tuple (MedR, MeanR, AUC, NDCG@10, NDCG@50, r1, r5, r10) of a submitted batch (blocks until it is there)."""
        if ticket not in self.results:
            slot = ticket % self.depth
            if self.pending[slot] != ticket:
                raise KeyError("ticket %r is neither pending nor collected" % (ticket,))
            self._collect(slot)
        return self.results.pop(ticket)

    def drain(self):
        for slot in range(self.depth):
            if self.pending[slot] is not None:
                self._collect(slot)
