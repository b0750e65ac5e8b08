"""torch.Tensor-level wrappers over the C ABI (include/frx.h).

torch is plumbing here: it owns device memory and the current stream; every wrapper unwraps
``data_ptr()`` and calls the hand-written sm_100a kernels in libfrx_b200.so.  Inputs that are not
CUDA tensors are rejected -- there is no CPU path.
"""
import torch

from . import _lib

VISUAL_NORM, TEXT_NORM, FINAL_NORM = 1, 2, 4


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _req(t, dtype, name, dims=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.FrxError("%s must be a CUDA tensor (fancyrec_b200 has no CPU fallback)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if dims is not None and t.dim() != dims:
        raise ValueError("%s must be %d-D" % (name, dims))
    return t


def round_up(v, m):
    return (v + m - 1) // m * m


# ---------------------------------------------------------------------------------------------
def finalize_posts(visual, text=None, row_ptr=None, row_idx=None, visual_norm=False, text_norm=False,
                   final_norm=True, want_f32=False, want_bf16=True, out_f32=None, out_bf16=None, blocks_per_sm=0):
    """A1-A3 in one pass.  Returns (out_f32 | None, out_bf16 | None); out_bf16 is [NP, round_up(D, 64)]
    with zero padding, the operand layout of score_*.  `out_f32` / `out_bf16`: write into these caller-owned
    buffers (same shapes) instead of allocating.  `blocks_per_sm` bounds the resident blocks per SM (0 = the shape's
    measured optimum; 1 = co-residency form for a launch that runs under a contraction, see pipeline.py)."""
    lib = _lib.load()
    _req(visual, torch.float32, "visual", 2)
    dv = visual.shape[1]
    if row_ptr is not None:
        _req(row_ptr, torch.int64, "row_ptr", 1)
        n_posts = row_ptr.numel() - 1
    else:
        n_posts = visual.shape[0]
    if row_idx is not None:
        _req(row_idx, torch.int32, "row_idx", 1)
    dt = 0
    if text is not None:
        _req(text, torch.float32, "text", 2)
        dt = text.shape[1]
        if text.shape[0] != n_posts:
            raise ValueError("text has %d rows, expected %d" % (text.shape[0], n_posts))
    d = dv + dt
    flags = (VISUAL_NORM if visual_norm else 0) | (TEXT_NORM if text_norm else 0) | (FINAL_NORM if final_norm else 0)
    ld = round_up(d, 64)
    if out_f32 is not None:
        _req(out_f32, torch.float32, "out_f32", 2)
        if tuple(out_f32.shape) != (n_posts, d):
            raise ValueError("out_f32 must be [%d, %d]" % (n_posts, d))
    elif want_f32:
        out_f32 = torch.empty((n_posts, d), dtype=torch.float32, device=visual.device)
    if out_bf16 is not None:
        _req(out_bf16, torch.bfloat16, "out_bf16", 2)
        if tuple(out_bf16.shape) != (n_posts, ld):
            raise ValueError("out_bf16 must be [%d, %d]" % (n_posts, ld))
    elif want_bf16:
        out_bf16 = torch.empty((n_posts, ld), dtype=torch.bfloat16, device=visual.device)
    with torch.cuda.device(visual.device):
        rc = lib.frx_finalize_posts_bounded(_ptr(visual), _ptr(row_ptr), _ptr(row_idx), _ptr(text), n_posts, dv, dt, flags,
                                            _ptr(out_f32), _ptr(out_bf16), ld, int(blocks_per_sm), _stream(visual))
    _lib.check(rc, "frx_finalize_posts")
    return out_f32, out_bf16


def masked_mean_pool(x, lengths, l2norm=False):
    """Mean over the first lengths[b] steps of every sequence: x [B, T, D] fp32, lengths [B] -> [B, D] fp32.
    Replaces the per-sample Python loops `torch.mean(out[i][:len_i], 0)` of the reference encoders
    (model.py:105-114, 163-168, 271-274, 344-346) with ONE pass of the pooling kernel: the padded steps are
    skipped through the gather list, so only valid rows are read."""
    _req(x, torch.float32, "x", 3)
    b, t, d = x.shape
    lengths = torch.as_tensor(lengths, device=x.device).to(torch.int64)
    if lengths.numel() != b:
        raise ValueError("lengths has %d entries, expected %d" % (lengths.numel(), b))
    row_ptr = torch.zeros(b + 1, dtype=torch.int64, device=x.device)
    row_ptr[1:] = torch.cumsum(lengths, 0)
    steps = torch.arange(t, device=x.device)
    valid = steps.unsqueeze(0) < lengths.unsqueeze(1)                                   # [B, T]
    row_idx = (torch.arange(b, device=x.device).unsqueeze(1) * t + steps.unsqueeze(0))[valid].to(torch.int32)
    return finalize_posts(x.reshape(b * t, d), row_ptr=row_ptr, row_idx=row_idx.contiguous(), final_norm=l2norm,
                          want_f32=True, want_bf16=False)[0]


def masked_softmax_pool(x, logits, lengths):
    """MultiHeadSelfAttention's weighting step (model.py:105-114) without the per-sample Python loop:
    x [B, T, D] fp32, logits [B, T] or [B, T, 1] fp32 (head-averaged attention scores), lengths [B] ->
    out[b] = mean over the padded T of softmax(logits[b, :len_b]) * x[b]  ([B, D] fp32)."""
    lib = _lib.load()
    _req(x, torch.float32, "x", 3)
    b, t, d = x.shape
    logits = logits.reshape(b, t)
    _req(logits, torch.float32, "logits", 2)
    lengths = torch.as_tensor(lengths, device=x.device).to(torch.int64).contiguous()
    if lengths.numel() != b:
        raise ValueError("lengths has %d entries, expected %d" % (lengths.numel(), b))
    out = torch.empty((b, d), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.frx_softmax_pool(_ptr(x), _ptr(logits), _ptr(lengths), b, t, d, _ptr(out), _stream(x))
    _lib.check(rc, "frx_softmax_pool")
    return out


def split_tf32x3(x, side):
    """fp32 [N, D] -> the K-concatenated 3xTF32 operand [N, 3D] (side 0 = brand: [hi|lo|hi]; side 1 = post:
    [hi|hi|lo]).  score_*(a3, b3, d=3D) on these gives fp32-grade scores on the tf32 tensor-core path."""
    lib = _lib.load()
    _req(x, torch.float32, "x", 2)
    out = torch.empty((x.shape[0], 3 * x.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.frx_split_tf32x3(_ptr(x), x.shape[0], x.shape[1], x.stride(0), int(side), _ptr(out), _stream(x))
    _lib.check(rc, "frx_split_tf32x3")
    return out


def brand_embed(w, e, brand_ids=None, nb=None, tensor_cores=True):
    """A4: out[i] = mean_a W[ids[i], a] * E[a, :]  -> [nb, D] fp32.  tensor_cores: 3xTF32 tcgen05 GEMM (fp32-grade);
    False: fp32 FMA GEMM on the CUDA cores."""
    lib = _lib.load()
    _req(w, torch.float32, "w", 2)
    _req(e, torch.float32, "e", 2)
    if brand_ids is not None:
        _req(brand_ids, torch.int64, "brand_ids", 1)
        nb = brand_ids.numel()
    elif nb is None:
        nb = w.shape[0]
    a, d = e.shape
    if w.shape[1] != a:
        raise ValueError("w is [*, %d] but e is [%d, *]" % (w.shape[1], a))
    out = torch.empty((nb, d), dtype=torch.float32, device=w.device)
    need = lib.frx_brand_embed_workspace_bytes(nb, a, d) if tensor_cores else 0
    ws = torch.empty(need, dtype=torch.uint8, device=w.device) if need else None
    with torch.cuda.device(w.device):
        rc = lib.frx_brand_embed(_ptr(w), w.shape[0], _ptr(e), _ptr(brand_ids), nb, a, d, _ptr(out), _ptr(ws), need,
                                 _stream(w))
    _lib.check(rc, "frx_brand_embed")
    return out


def brand_train_fwd(w_rows, e, seed):
    """Training-time brand embedding with dropout(0.5) on the products, no [B, A, D] tensor: -> [B, D] fp32."""
    lib = _lib.load()
    _req(w_rows, torch.float32, "w_rows", 2)
    _req(e, torch.float32, "e", 2)
    b, a = w_rows.shape
    if e.shape[0] != a:
        raise ValueError("w_rows is [*, %d] but e is [%d, *]" % (a, e.shape[0]))
    out = torch.empty((b, e.shape[1]), dtype=torch.float32, device=w_rows.device)
    with torch.cuda.device(w_rows.device):
        rc = lib.frx_brand_train_fwd(_ptr(w_rows), w_rows.stride(0), _ptr(e), b, a, e.shape[1], int(seed), _ptr(out),
                                     _stream(w_rows))
    _lib.check(rc, "frx_brand_train_fwd")
    return out


def brand_train_bwd(grad_out, w_rows, e, seed):
    """-> (d_w_rows [B, A] incl. the L1Penalty term 1e-4 * sign(w), d_e [A, D]) for the mask of `seed`."""
    lib = _lib.load()
    _req(grad_out, torch.float32, "grad_out", 2)
    _req(w_rows, torch.float32, "w_rows", 2)
    _req(e, torch.float32, "e", 2)
    b, a = w_rows.shape
    d_w = torch.empty((b, a), dtype=torch.float32, device=w_rows.device)
    d_e = torch.empty_like(e)
    with torch.cuda.device(w_rows.device):
        rc = lib.frx_brand_train_bwd(_ptr(grad_out), _ptr(w_rows), w_rows.stride(0), _ptr(e), b, a, e.shape[1], int(seed),
                                     _ptr(d_w), _ptr(d_e), _stream(w_rows))
    _lib.check(rc, "frx_brand_train_bwd")
    return d_w, d_e


def brand_dropout_mask(b, a, d, seed, device):
    """The keep mask m(b, a, d) of `seed` as uint8 [B, A, D] (tests: parity with the reference formula under a fixed mask)."""
    lib = _lib.load()
    mask = torch.empty((b, a, d), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        rc = lib.frx_brand_dropout_mask(b, a, d, int(seed), _ptr(mask), torch.cuda.current_stream(device).cuda_stream)
    _lib.check(rc, "frx_brand_dropout_mask")
    return mask


# ---------------------------------------------------------------------------------------------
def _operands(brand_op, post_op, d):
    """Operands are both bf16 (default precision) or both fp32 (tf32 tensor-core path)."""
    dt = post_op.dtype if isinstance(post_op, torch.Tensor) else None
    if dt not in (torch.bfloat16, torch.float32):
        raise TypeError("score operands must be bfloat16 or float32 CUDA tensors")
    _req(brand_op, dt, "brand operand", 2)
    _req(post_op, dt, "post operand", 2)
    if d is None:
        d = min(brand_op.shape[1], post_op.shape[1])
    return d


def _variant(lib, name, post_op):
    return getattr(lib, name + ("_tf32" if post_op.dtype == torch.float32 else ""))


def score_topk(brand_op, post_op, k, d=None, labels=None, index_base=0, workspace=None, dense=False):
    """A5+A6 fused.  Returns dict(scores [NB,k] f32, index [NB,k] i32, pos_score [NP] f32 | None,
    dense [NB,NP] f32 | None)."""
    lib = _lib.load()
    d = _operands(brand_op, post_op, d)
    nb, n_posts = brand_op.shape[0], post_op.shape[0]
    dev = post_op.device
    need = lib.frx_score_topk_workspace_bytes(nb, n_posts, d, k)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    scores = torch.empty((nb, k), dtype=torch.float32, device=dev)
    index = torch.empty((nb, k), dtype=torch.int32, device=dev)
    pos_score = None
    if labels is not None:
        _req(labels, torch.int32, "labels", 1)
        pos_score = torch.empty(n_posts, dtype=torch.float32, device=dev)
    dense_out = torch.empty((nb, n_posts), dtype=torch.float32, device=dev) if dense else None
    with torch.cuda.device(dev):
        rc = _variant(lib, "frx_score_topk", post_op)(_ptr(brand_op), brand_op.stride(0), _ptr(post_op), post_op.stride(0), nb,
                                n_posts, d, k, _ptr(labels), index_base, _ptr(scores), _ptr(index), _ptr(pos_score),
                                _ptr(dense_out), n_posts, _ptr(workspace), workspace.numel(), _stream(post_op))
    _lib.check(rc, "frx_score_topk")
    return dict(scores=scores, index=index, pos_score=pos_score, dense=dense_out, workspace=workspace)


def score_dense(brand_op, post_op, d=None, out=None):
    lib = _lib.load()
    d = _operands(brand_op, post_op, d)
    nb, n_posts = brand_op.shape[0], post_op.shape[0]
    if out is None:
        out = torch.empty((nb, n_posts), dtype=torch.float32, device=post_op.device)
    with torch.cuda.device(post_op.device):
        rc = _variant(lib, "frx_score_dense", post_op)(_ptr(brand_op), brand_op.stride(0), _ptr(post_op), post_op.stride(0), nb,
                                 n_posts, d, _ptr(out), out.stride(0), _stream(post_op))
    _lib.check(rc, "frx_score_dense")
    return out


def score_count(brand_op, post_op, thr_score, thr_index, d=None, index_base=0, out=None):
    """Per brand: number of posts preceding (thr_score, thr_index).  Accumulates into ``out`` (int64)."""
    lib = _lib.load()
    d = _operands(brand_op, post_op, d)
    nb, n_posts = brand_op.shape[0], post_op.shape[0]
    _req(thr_score, torch.float32, "thr_score", 1)
    _req(thr_index, torch.int32, "thr_index", 1)
    if out is None:
        out = torch.zeros(nb, dtype=torch.int64, device=post_op.device)
    with torch.cuda.device(post_op.device):
        rc = _variant(lib, "frx_score_count", post_op)(_ptr(brand_op), brand_op.stride(0), _ptr(post_op), post_op.stride(0), nb,
                                 n_posts, d, index_base, _ptr(thr_score), _ptr(thr_index), _ptr(out),
                                 _stream(post_op))
    _lib.check(rc, "frx_score_count")
    return out


def topk_merge(scores, index, k_out):
    """[G, NB, k_in] candidate lists -> merged [NB, k_out] (score desc, index asc)."""
    lib = _lib.load()
    _req(scores, torch.float32, "scores", 3)
    _req(index, torch.int32, "index", 3)
    g, nb, k_in = scores.shape
    out_s = torch.empty((nb, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nb, k_out), dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = lib.frx_topk_merge(_ptr(scores), _ptr(index), g, nb, k_in, _ptr(out_s), _ptr(out_i), k_out,
                                _stream(scores))
    _lib.check(rc, "frx_topk_merge")
    return out_s, out_i


def merge_gathered(gathered, nb, k_in, k_out):
    """Multi-GPU exchange, device side.  `gathered` int32 [G, W]: rank r's row is the packed block
    [scores nb*k_in (fp32 bits) | index nb*k_in | n_pos nb | best_score nb (fp32 bits) | best_index nb] it contributed to
    the all-gather.  Returns (top scores [nb, k_out], top index, n_pos [nb], best_score [nb], best_index [nb]) of the
    whole job: the lists are merged and the label statistics combined in place out of that buffer (two kernels)."""
    lib = _lib.load()
    if not (isinstance(gathered, torch.Tensor) and gathered.is_cuda and gathered.dtype == torch.int32 and
            gathered.dim() == 2 and gathered.stride(1) == 1):
        raise _lib.FrxError("gathered must be a 2-D int32 CUDA tensor with unit column stride")
    g, head = gathered.shape
    w = gathered.stride(0) if g > 1 else head          # rows may be longer than the head (labels ride behind it)
    if head != 2 * nb * k_in + 3 * nb:
        raise ValueError("packed row has %d words, expected %d" % (head, 2 * nb * k_in + 3 * nb))
    dev = gathered.device
    out_s = torch.empty((nb, k_out), dtype=torch.float32, device=dev)
    out_i = torch.empty((nb, k_out), dtype=torch.int32, device=dev)
    n_pos = torch.empty(nb, dtype=torch.int32, device=dev)
    best_s = torch.empty(nb, dtype=torch.float32, device=dev)
    best_i = torch.empty(nb, dtype=torch.int32, device=dev)
    base = gathered.data_ptr()
    o_idx, o_np = 4 * nb * k_in, 8 * nb * k_in
    with torch.cuda.device(dev):
        rc = lib.frx_topk_merge_strided(base, base + o_idx, g, nb, k_in, w, _ptr(out_s), _ptr(out_i), k_out,
                                        _stream(gathered))
        _lib.check(rc, "frx_topk_merge_strided")
        rc = lib.frx_reduce_shard_stats(base + o_np, base + o_np + 4 * nb, base + o_np + 8 * nb, g, nb, w, _ptr(n_pos),
                                        _ptr(best_s), _ptr(best_i), _stream(gathered))
        _lib.check(rc, "frx_reduce_shard_stats")
    return out_s, out_i, n_pos, best_s, best_i


# ---------------------------------------------------------------------------------------------
def label_stats(labels, pos_score, nb, index_base=0):
    lib = _lib.load()
    _req(labels, torch.int32, "labels", 1)
    _req(pos_score, torch.float32, "pos_score", 1)
    dev = labels.device
    n_pos = torch.empty(nb, dtype=torch.int32, device=dev)
    best_score = torch.empty(nb, dtype=torch.float32, device=dev)
    best_index = torch.empty(nb, dtype=torch.int32, device=dev)
    ws = torch.empty(nb, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.frx_label_stats(_ptr(labels), _ptr(pos_score), labels.numel(), nb, index_base, _ptr(n_pos),
                                 _ptr(best_score), _ptr(best_index), _ptr(ws), _stream(labels))
    _lib.check(rc, "frx_label_stats")
    return n_pos, best_score, best_index


def rank_from_topk(topk_index, labels, index_base=0):
    lib = _lib.load()
    _req(topk_index, torch.int32, "topk_index", 2)
    _req(labels, torch.int32, "labels", 1)
    nb, k = topk_index.shape
    dev = labels.device
    hit_mask = torch.empty(nb, dtype=torch.int64, device=dev)
    first_rank = torch.empty(nb, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.frx_rank_from_topk(_ptr(topk_index), nb, k, _ptr(labels), labels.numel(), index_base,
                                    _ptr(hit_mask), _ptr(first_rank), _stream(labels))
    _lib.check(rc, "frx_rank_from_topk")
    return hit_mask, first_rank


def missing_thresholds(n_pos, first_in_list, best_index):
    """thr_index for score_count: best_index where the first positive fell outside the top-k list, else -1."""
    lib = _lib.load()
    _req(n_pos, torch.int32, "n_pos", 1)
    _req(first_in_list, torch.int32, "first_in_list", 1)
    _req(best_index, torch.int32, "best_index", 1)
    out = torch.empty_like(best_index)
    with torch.cuda.device(n_pos.device):
        rc = lib.frx_missing_thresholds(_ptr(n_pos), _ptr(first_in_list), _ptr(best_index), n_pos.numel(), _ptr(out),
                                        _stream(n_pos))
    _lib.check(rc, "frx_missing_thresholds")
    return out


def pack_rank_stats(n_pos, first_in_list, before_first, hit_mask, auc_num=None, all_valid=False):
    """-> int64 [5 | 6, NB]: n_pos, first rank in list, count before first positive, its validity, hit mask (, AUC num)."""
    lib = _lib.load()
    _req(n_pos, torch.int32, "n_pos", 1)
    _req(first_in_list, torch.int32, "first_in_list", 1)
    _req(before_first, torch.int64, "before_first", 1)
    _req(hit_mask, torch.int64, "hit_mask", 1)
    if auc_num is not None:
        _req(auc_num, torch.int64, "auc_num", 1)
    nb = n_pos.numel()
    out = torch.empty((6 if auc_num is not None else 5, nb), dtype=torch.int64, device=n_pos.device)
    with torch.cuda.device(n_pos.device):
        rc = lib.frx_pack_rank_stats(_ptr(n_pos), _ptr(first_in_list), _ptr(before_first), _ptr(hit_mask), _ptr(auc_num),
                                     nb, int(bool(all_valid)), _ptr(out), _stream(n_pos))
    _lib.check(rc, "frx_pack_rank_stats")
    return out


def group_positives(labels, pos_score, n_pos):
    lib = _lib.load()
    _req(labels, torch.int32, "labels", 1)
    _req(pos_score, torch.float32, "pos_score", 1)
    _req(n_pos, torch.int32, "n_pos", 1)
    nb = n_pos.numel()
    dev = labels.device
    seg_ptr = torch.empty(nb + 1, dtype=torch.int64, device=dev)
    pos_sorted = torch.empty(max(labels.numel(), 1), dtype=torch.float32, device=dev)
    ws = torch.empty(nb, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.frx_group_positives(_ptr(labels), _ptr(pos_score), labels.numel(), nb, _ptr(n_pos), _ptr(seg_ptr),
                                     _ptr(pos_sorted), _ptr(ws), _stream(labels))
    _lib.check(rc, "frx_group_positives")
    return seg_ptr, pos_sorted


def auc_rows(scores, row0, labels, seg_ptr, pos_sorted, best_score, best_index, auc_num, before_first, index_base=0):
    lib = _lib.load()
    _req(scores, torch.float32, "scores", 2)
    _req(auc_num, torch.int64, "auc_num", 1)
    _req(before_first, torch.int64, "before_first", 1)
    n_rows, n_posts = scores.shape
    with torch.cuda.device(scores.device):
        rc = lib.frx_auc_rows(_ptr(scores), scores.stride(0), row0, n_rows, n_posts, _ptr(labels), _ptr(seg_ptr),
                              _ptr(pos_sorted), _ptr(best_score), _ptr(best_index), index_base, _ptr(auc_num),
                              _ptr(before_first), _stream(scores))
    _lib.check(rc, "frx_auc_rows")


def linear(x, weight, bias=None, col_scale=None, relu=False, out=None):
    """Encoder-side Linear layer on the tensor cores: out = act((x @ weight.T) * col_scale + bias), fp32-grade (3xTF32
    split inside the kernel).  x [M, K], weight [N, K] (nn.Linear layout), bias / col_scale [N] or None -- an eval-mode
    BatchNorm1d folds into (col_scale, bias), see model.fold_batchnorm."""
    lib = _lib.load()
    for name, t in (("x", x), ("weight", weight)) + ((("out", out),) if out is not None else ()):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise _lib.FrxError("%s must be a CUDA tensor (fancyrec_b200 has no CPU fallback)" % name)
        if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError("%s must be a 2-D float32 tensor with unit column stride (row pitch is free)" % name)
    m, k = x.shape
    n = weight.shape[0]
    if weight.shape[1] != k:
        raise ValueError("x is [*, %d] but weight is [*, %d]" % (k, weight.shape[1]))
    for name, v in (("bias", bias), ("col_scale", col_scale)):
        if v is not None:
            _req(v, torch.float32, name, 1)
            if v.numel() != n:
                raise ValueError("%s has %d entries, expected %d" % (name, v.numel(), n))
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    if m == 0:
        return out
    need = lib.frx_linear_workspace_bytes(m, n, k)
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.frx_linear(_ptr(x), x.stride(0), _ptr(weight), weight.stride(0), _ptr(col_scale), _ptr(bias), int(bool(relu)),
                            m, n, k, _ptr(out), out.stride(0), _ptr(ws), need, _stream(x))
    _lib.check(rc, "frx_linear")
    return out


def matmul3x(a, b, a_transposed=False, b_transposed=False, alpha=1.0, out=None):
    """out[M, N] = alpha * sum_k A(m, k) B(n, k) on the tensor cores, fp32-grade (the kernel behind `linear` and the
    losses' gradient products).  A is a [M, K], or a [K, M] when a_transposed; B is b [N, K], or b [K, N] when
    b_transposed -- transposed operands are read in place (no copy)."""
    lib = _lib.load()
    for name, t in (("a", a), ("b", b)) + ((("out", out),) if out is not None else ()):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise _lib.FrxError("%s must be a CUDA tensor (fancyrec_b200 has no CPU fallback)" % name)
        if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError("%s must be a 2-D float32 tensor with unit column stride (row pitch is free)" % name)
    k, m = (a.shape if a_transposed else a.shape[::-1])
    kb, n = (b.shape if b_transposed else b.shape[::-1])
    if k != kb:
        raise ValueError("inner dimensions differ: %d vs %d" % (k, kb))
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    if m == 0 or n == 0:
        return out
    need = lib.frx_linear_workspace_bytes(m, n, k)
    ws = torch.empty(need, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        rc = lib.frx_matmul3x(_ptr(a), a.stride(0), int(bool(a_transposed)), _ptr(b), b.stride(0), int(bool(b_transposed)),
                              m, n, k, float(alpha), _ptr(out), out.stride(0), _ptr(ws), need, _stream(a))
    _lib.check(rc, "frx_matmul3x")
    return out


SCORER_KINDS = {"P": 0, "AP": 1, "RR": 2, "NDCG": 3, "DCG": 4}
_LOG2_TABLES = {}


def metric_scores(labels, kind, k=0, lengths=None):
    """A10 on the device: util/metric.py's scorer `kind` ("P", "AP", "RR", "NDCG", "DCG") at cut-off k over a batch of
    sorted label lists -- labels int32 [N, L] on the device, lengths [N] int32 (None: all L long) -> float64 [N],
    bit-identical to getScorer("KIND@k").score(list) per list (NaN where the reference raises)."""
    import math
    lib = _lib.load()
    _req(labels, torch.int32, "labels", 2)
    n, length = labels.shape
    if lengths is not None:
        _req(lengths, torch.int32, "lengths", 1)
    key = (labels.device, length)
    table = _LOG2_TABLES.get(key)
    if table is None:      # log2_table[i] = math.log(i, 2): the reference's own expression, evaluated on the host
        table = torch.tensor([0.0] + [math.log(i, 2) for i in range(1, length + 2)], dtype=torch.float64).to(labels.device)
        _LOG2_TABLES[key] = table
    out = torch.empty(n, dtype=torch.float64, device=labels.device)
    if n == 0:
        return out
    with torch.cuda.device(labels.device):
        rc = lib.frx_metric_scores(_ptr(labels), labels.stride(0), _ptr(lengths), n, length, SCORER_KINDS[kind], int(k),
                                   _ptr(table), _ptr(out), _stream(labels))
    _lib.check(rc, "frx_metric_scores")
    return out


# ---------------------------------------------------------------------------------------------
def triplet_fwd_bwd(brand_ids, brand, post, margin, mean_style, want_grad=True):
    lib = _lib.load()
    _req(brand_ids, torch.int64, "brand_ids", 1)
    _req(brand, torch.float32, "brand", 2)
    _req(post, torch.float32, "post", 2)
    b, d = brand.shape
    dev = brand.device
    ws = torch.empty(lib.frx_triplet_workspace_bytes(b, d), dtype=torch.uint8, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d_brand = torch.empty_like(brand) if want_grad else None
    d_post = torch.empty_like(post) if want_grad else None
    with torch.cuda.device(dev):
        rc = lib.frx_triplet_fwd_bwd(_ptr(brand_ids), _ptr(brand), _ptr(post), b, d, float(margin), int(mean_style),
                                     _ptr(loss), _ptr(d_brand), _ptr(d_post), _ptr(ws), ws.numel(), _stream(brand))
    _lib.check(rc, "frx_triplet_fwd_bwd")
    return loss, d_brand, d_post


def vsepp_fwd_bwd(brand_ids, brand, post, margin, mean_style, want_grad=True):
    """Opt-in hardest-negative (VSE++) hinge on the in-batch tile: see frx_vsepp_fwd_bwd in include/frx.h."""
    lib = _lib.load()
    _req(brand_ids, torch.int64, "brand_ids", 1)
    _req(brand, torch.float32, "brand", 2)
    _req(post, torch.float32, "post", 2)
    b, d = brand.shape
    dev = brand.device
    ws = torch.empty(lib.frx_triplet_workspace_bytes(b, d), dtype=torch.uint8, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d_brand = torch.empty_like(brand) if want_grad else None
    d_post = torch.empty_like(post) if want_grad else None
    with torch.cuda.device(dev):
        rc = lib.frx_vsepp_fwd_bwd(_ptr(brand_ids), _ptr(brand), _ptr(post), b, d, float(margin), int(mean_style),
                                   _ptr(loss), _ptr(d_brand), _ptr(d_post), _ptr(ws), ws.numel(), _stream(brand))
    _lib.check(rc, "frx_vsepp_fwd_bwd")
    return loss, d_brand, d_post


def normalize_rows(x):
    lib = _lib.load()
    _req(x, torch.float32, "x", 2)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib.frx_normalize_rows(_ptr(x), x.shape[0], x.shape[1], _ptr(out), _stream(x))
    _lib.check(rc, "frx_normalize_rows")
    return out


def contrastive_fwd_bwd(brand, post, keys, mask_col0, no_intra, temperature, negative_weight, mean_style,
                        want_grad=True):
    lib = _lib.load()
    _req(brand, torch.float32, "brand", 2)
    _req(post, torch.float32, "post", 2)
    b, d = brand.shape
    n_keys = 0
    if keys is not None:
        _req(keys, torch.float32, "keys", 2)
        n_keys = keys.shape[0]
    dev = brand.device
    ws = torch.empty(lib.frx_contrastive_workspace_bytes(b, d, n_keys), dtype=torch.uint8, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d_brand = torch.empty_like(brand) if want_grad else None
    d_post = torch.empty_like(post) if want_grad else None
    with torch.cuda.device(dev):
        rc = lib.frx_contrastive_fwd_bwd(_ptr(brand), _ptr(post), b, d, _ptr(keys), n_keys, int(mask_col0),
                                         int(bool(no_intra)), float(temperature), float(negative_weight),
                                         int(mean_style), _ptr(loss), _ptr(d_brand), _ptr(d_post), _ptr(ws),
                                         ws.numel(), _stream(brand))
    _lib.check(rc, "frx_contrastive_fwd_bwd")
    return loss, d_brand, d_post


def crossclr_fwd_bwd(brand, post, temperature, negative_weight, mean_style, want_grad=True):
    lib = _lib.load()
    _req(brand, torch.float32, "brand", 2)
    _req(post, torch.float32, "post", 2)
    b, d = brand.shape
    dev = brand.device
    ws = torch.empty(lib.frx_crossclr_workspace_bytes(b, d), dtype=torch.uint8, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d_brand = torch.empty_like(brand) if want_grad else None
    d_post = torch.empty_like(post) if want_grad else None
    with torch.cuda.device(dev):
        rc = lib.frx_crossclr_fwd_bwd(_ptr(brand), _ptr(post), b, d, float(temperature), float(negative_weight),
                                      int(mean_style), _ptr(loss), _ptr(d_brand), _ptr(d_post), _ptr(ws), ws.numel(),
                                      _stream(brand))
    _lib.check(rc, "frx_crossclr_fwd_bwd")
    return loss, d_brand, d_post


def lab_fwd_bwd(brand, want_grad=True):
    lib = _lib.load()
    _req(brand, torch.float32, "brand", 2)
    b, d = brand.shape
    dev = brand.device
    ws = torch.empty(lib.frx_lab_workspace_bytes(b, d), dtype=torch.uint8, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d_brand = torch.empty_like(brand) if want_grad else None
    with torch.cuda.device(dev):
        rc = lib.frx_lab_fwd_bwd(_ptr(brand), b, d, _ptr(loss), _ptr(d_brand), _ptr(ws), ws.numel(), _stream(brand))
    _lib.check(rc, "frx_lab_fwd_bwd")
    return loss, d_brand
