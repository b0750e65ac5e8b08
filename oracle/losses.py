"""Oracle restatement of the in-batch similarity-tile losses (TEST INFRASTRUCTURE ONLY).

Forward AND analytic backward, float64 NumPy.  Reference lines:
  * TripletLoss.forward            loss.py:87-143
  * ContrastiveLoss.forward        loss_ctrs.py:179-214 (+ :138-177 helpers)
  * CrossCLR_onlyIntraModality     loss_ctrs.py:52-117
  * LabLoss                        loss.py:55-63
Pinned against the reference's autograd in tests/test_oracle_losses.py via
tests/golden/losses_*.npz.

Rank tie rule (reference leaves it to torch.sort, unspecified): position of the
diagonal = #strictly greater + #equal with a smaller index.
Hinge derivative at exactly 0 follows torch.clamp backward (grad passes where
x >= min).
"""
import numpy as np


def _rank_weight_rows(s):
    """loss.py:96-100: rank_p[i] = 1/(B - rank_1[i] + 1) + 1, rank_1[i] = 1 + position of
    S[i, i] in row i sorted descending."""
    b = s.shape[0]
    d = np.diag(s)[:, None]
    idx = np.arange(b)
    pos = (s > d).sum(1) + ((s == d) & (idx[None, :] < idx[:, None])).sum(1)
    rank_1 = (pos + 1).astype(np.float32)
    return (1.0 / (np.float32(b) - rank_1 + 1.0) + 1.0).astype(np.float32), pos


def _rank_weight_cols(s):
    """loss.py:102-105: same along columns (position of S[j, j] in column j)."""
    w, pos = _rank_weight_rows(s.T)
    return w, pos


def sim_tile(brand, post):
    """loss.py:91-93: S[i, j] = post_i . brand_j (raw dot product, no normalisation)."""
    return np.asarray(post, np.float64) @ np.asarray(brand, np.float64).T


def triplet_loss(brand_ids, brand, post, margin=0.0, cost_style='sum', s_override=None):
    """Returns (loss, d_brand, d_post, aux).  loss.py:87-143 with direction='all'.
    max_violation / measure / loss_fun are accepted by the reference constructor and
    never read in forward (loss.py:79-85 vs :87-143) -- there is nothing to restate."""
    brand = np.asarray(brand, np.float64)
    post = np.asarray(post, np.float64)
    ids = np.asarray(brand_ids)
    b = brand.shape[0]
    # s_override: evaluate everything downstream of the tile on a GIVEN tile (the tests pass the device's own fp32
    # tile, so that the integer ranks and the hinge active set are decided on identical numbers)
    s = sim_tile(brand, post) if s_override is None else np.asarray(s_override, np.float64)
    s32 = s.astype(np.float32)
    rank_p, pos_r = _rank_weight_rows(s32)
    rank_b, pos_c = _rank_weight_cols(s32)
    diag = np.diag(s)
    mask = ids[:, None] == ids[None, :]                      # loss.py:116-119
    xp = margin + s - diag[:, None]                          # d1: S[i,i] along row i
    xb = margin + s - diag[None, :]                          # d2: S[j,j] along column j
    cost_p = np.where(mask, 0.0, np.maximum(xp, 0.0))
    cost_b = np.where(mask, 0.0, np.maximum(xb, 0.0))
    # loss.py:131-132: the [B] weight vectors broadcast along the LAST axis (column index)
    wp = rank_p.astype(np.float64)[None, :]
    wb = rank_b.astype(np.float64)[None, :]
    scale = 1.0 if cost_style == 'sum' else 1.0 / (b * b)
    loss = scale * ((cost_b * wb).sum() + (cost_p * wp).sum())
    gp = np.where(mask | (xp < 0), 0.0, 1.0) * wp            # d loss / d cost_p-argument
    gb = np.where(mask | (xb < 0), 0.0, 1.0) * wb
    ds = gp + gb
    ds[np.arange(b), np.arange(b)] -= gp.sum(1)              # -S[i,i] in every entry of row i
    ds[np.arange(b), np.arange(b)] -= gb.sum(0)              # -S[j,j] in every entry of column j
    ds *= scale
    d_post = ds @ brand
    d_brand = ds.T @ post
    return loss, d_brand, d_post, dict(s=s, ds=ds, rank_p=rank_p, rank_b=rank_b,
                                       pos_r=pos_r, pos_c=pos_c, active_p=~(mask | (xp < 0)), active_b=~(mask | (xb < 0)))


def vsepp_loss(brand_ids, brand, post, margin=0.0, cost_style='sum', s_override=None):
    """The hardest-negative hinge (VSE++, Faghri et al. 2018: `cost_s.max(1)[0] + cost_im.max(0)[0]`) on the reference's
    tile S[i,j] = post_i . brand_j with its same-brand mask (loss.py:116-119) -- the mode loss.py's `max_violation` flag
    names and never implements (checked against this restatement, never against loss.py).  No rank
    weights; 'mean' divides by B.  Returns (loss, d_brand, d_post, aux); ties take the first position."""
    brand = np.asarray(brand, np.float64)
    post = np.asarray(post, np.float64)
    ids = np.asarray(brand_ids)
    b = brand.shape[0]
    s = sim_tile(brand, post) if s_override is None else np.asarray(s_override, np.float64)
    diag = np.diag(s)
    mask = ids[:, None] == ids[None, :]
    cost_p = np.where(mask, 0.0, np.maximum(margin + s - diag[:, None], 0.0))     # row i against S[i,i]
    cost_b = np.where(mask, 0.0, np.maximum(margin + s - diag[None, :], 0.0))     # column j against S[j,j]
    jr = cost_p.argmax(1)
    ic = cost_b.argmax(0)
    vr = cost_p[np.arange(b), jr]
    vc = cost_b[ic, np.arange(b)]
    scale = 1.0 if cost_style == 'sum' else 1.0 / b
    loss = scale * (vr.sum() + vc.sum())
    ds = np.zeros((b, b))
    for i in range(b):
        if vr[i] > 0:
            ds[i, jr[i]] += scale
            ds[i, i] -= scale
        if vc[i] > 0:
            ds[ic[i], i] += scale
            ds[i, i] -= scale
    return loss, ds.T @ post, ds @ brand, dict(s=s, ds=ds, row_arg=np.where(vr > 0, jr, -1), col_arg=np.where(vc > 0, ic, -1))


def _normalize(x, eps=1e-12):
    n = np.maximum(np.sqrt((x * x).sum(1, keepdims=True)), eps)   # F.normalize
    return x / n, n


def _normalize_bwd(dy, y, n):
    return (dy - y * (y * dy).sum(1, keepdims=True)) / n


def contrastive_loss(brand, post, queue=None, queue_ptr=0, temperature=0.03, negative_weight=0.8,
                     cost_style='sum', no_queue=False, no_intra=False):
    """loss_ctrs.py:179-214.  Returns (loss, d_brand, d_post, new_queue, new_ptr).

    queue [Q, D] is updated as loss_ctrs.py:138-147 does (rows ptr:ptr+B <- normalised
    post, detached); the positive mask uses the pointer AFTER the move
    (loss_ctrs.py:149-159).  An out-of-range mask column raises IndexError as the
    reference does."""
    brand = np.asarray(brand, np.float64)
    post = np.asarray(post, np.float64)
    b = brand.shape[0]
    s32 = sim_tile(brand, post).astype(np.float32)
    weight, _ = _rank_weight_rows(s32)                       # loss_ctrs.py:182-192
    weight = weight.astype(np.float64)
    bn, nb_ = _normalize(brand)
    pn, np_ = _normalize(post)
    self_keys = no_queue or no_intra
    if self_keys:
        keys = pn
        ptr = int(queue_ptr)
        new_queue, new_ptr = queue, ptr
    else:
        q = queue.shape[0]
        ptr = int(queue_ptr)
        new_queue = np.array(queue, dtype=np.float64, copy=True)
        if ptr + b > q:
            raise RuntimeError("queue slice shorter than batch (loss_ctrs.py:146)")
        new_queue[ptr:ptr + b] = pn
        new_ptr = (ptr + b) % q
        keys = new_queue
        ptr = new_ptr
    ori = pn @ keys.T
    mask = np.ones_like(ori)
    for i in range(b):
        if ptr + i >= ori.shape[1]:
            # the reference has already mutated queue/queue_ptr at this point (loss_ctrs.py:200-201)
            err = IndexError("positive-mask column out of range (loss_ctrs.py:155-158)")
            err.state = (new_queue, new_ptr)
            raise err
        mask[i, ptr + i] = 0.0
    inter = bn @ pn.T / temperature
    intra = ori * mask / temperature
    if no_intra:
        intra = np.zeros_like(intra)
    e_inter = np.exp(inter)
    e_intra = np.exp(intra)
    z = e_inter.sum(1) + negative_weight * e_intra.sum(1)
    p = np.diag(e_inter) / z
    scale = 1.0 if cost_style == 'sum' else 1.0 / b
    loss = scale * (-np.log(p) * weight).sum()
    # backward
    d_inter = scale * weight[:, None] * (e_inter / z[:, None])
    d_inter[np.arange(b), np.arange(b)] -= scale * weight
    d_bn = d_inter @ pn / temperature
    d_pn = d_inter.T @ bn / temperature
    if not no_intra:
        g = scale * weight[:, None] * negative_weight * e_intra / z[:, None] * mask / temperature
        if self_keys:
            d_pn += g @ pn + g.T @ pn
        else:
            d_pn += g @ keys                                   # queue rows are detached copies
    d_brand = _normalize_bwd(d_bn, bn, nb_)
    d_post = _normalize_bwd(d_pn, pn, np_)
    return loss, d_brand, d_post, new_queue, new_ptr


def crossclr_loss(brand, post, temperature=0.03, negative_weight=0.8, cost_style='sum'):
    """loss_ctrs.py:52-117 forward value only."""
    brand = np.asarray(brand, np.float64)
    post = np.asarray(post, np.float64)
    b = brand.shape[0]
    s32 = sim_tile(brand, post).astype(np.float32)
    rank_p, _ = _rank_weight_rows(s32)
    rank_b, _ = _rank_weight_cols(s32)
    bn, _ = _normalize(brand)
    pn, _ = _normalize(post)
    off = 1.0 - np.eye(b)
    lb = np.concatenate([bn @ pn.T / temperature, negative_weight * (bn @ bn.T / temperature) * off], 1)
    lp = np.concatenate([pn @ bn.T / temperature, negative_weight * (pn @ pn.T / temperature) * off], 1)

    def nll(logits):
        m = logits.max(1, keepdims=True)
        e = np.exp(logits - m)
        return -np.log(np.diag(e[:, :b]) / e.sum(1))
    loss_b = rank_b.astype(np.float64) * nll(lb)
    loss_p = rank_p.astype(np.float64) * nll(lp)
    if cost_style == 'sum':
        return (loss_b.sum() + loss_p.sum()) / 2
    return (loss_b.mean() + loss_p.mean()) / 2


def lab_loss(brand):
    """loss.py:55-63: (sum(exp(cos(brand, brand) with zeroed diagonal)) - B) / B."""
    x = np.asarray(brand, np.float64)
    n = x / np.sqrt((x * x).sum(1, keepdims=True))
    s = n @ n.T
    np.fill_diagonal(s, 0.0)
    return (np.exp(s).sum() - s.shape[0]) / s.shape[0]
