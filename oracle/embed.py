"""Oracle restatement of post finalisation and brand embedding (TEST INFRASTRUCTURE ONLY).

Reference lines:
  * frame mean-pool   util/data_provider.py:40,91,132  (videos_origin[i] = torch.mean(frames, 0), ALL frames)
  * l2norm            model.py:39-44 (no epsilon)
  * concat            model.py:482-485 (torch.cat((visual, text), 1))
  * per-branch norms  model.py:207-208 (visual_norm), :301-302 / :382-383 (text_norm)
  * brand embedding   model.py:419-428 (BrandAspects.forward, eval mode: dropout off,
                      L1Penalty forward = identity) + model.py:594 / evaluator.py:94
                      (.permute(1,0,2).mean(0) == mean over the aspect axis)
Accumulation is done in float64 and rounded once to float32; the CUDA path and
torch accumulate in float32, so comparisons use rtol 2e-6 (stated in the tests).
"""
import numpy as np


def mean_pool_csr(frames, row_ptr):
    """frames [Nrows, Dv] fp32 (feature.bin rows), row_ptr [NP+1] -> [NP, Dv] fp32."""
    frames = np.asarray(frames, dtype=np.float32)
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    out = np.empty((len(row_ptr) - 1, frames.shape[1]), dtype=np.float32)
    for p in range(len(row_ptr) - 1):
        seg = frames[row_ptr[p]:row_ptr[p + 1]].astype(np.float64)
        out[p] = (seg.sum(axis=0) / seg.shape[0]).astype(np.float32) if seg.shape[0] else np.nan
    return out


def mean_pool_gather(frames, row_idx, row_ptr):
    """Same, but post p owns rows frames[row_idx[row_ptr[p]:row_ptr[p+1]]] -- the rows
    of one video are not guaranteed contiguous in feature.bin (preprocess/get_frameInfo.py:55)."""
    frames = np.asarray(frames, dtype=np.float32)
    out = np.empty((len(row_ptr) - 1, frames.shape[1]), dtype=np.float32)
    for p in range(len(row_ptr) - 1):
        seg = frames[np.asarray(row_idx[row_ptr[p]:row_ptr[p + 1]], dtype=np.int64)].astype(np.float64)
        out[p] = (seg.sum(axis=0) / seg.shape[0]).astype(np.float32) if seg.shape[0] else np.nan
    return out


def l2norm64(x):
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        return x / np.sqrt((x * x).sum(axis=1, keepdims=True))


def finalize_posts(visual, text=None, visual_norm=True, text_norm=True, final_norm=True):
    """visual [NP, Dv] (already pooled), text [NP, Dt] or None ->
    [NP, Dv+Dt] fp32: optional per-branch l2norm, concat, optional whole-row l2norm
    (the l2norm cal_sim applies to the post operand, evaluator.py:28)."""
    v = np.asarray(visual, dtype=np.float64)
    if visual_norm:
        v = l2norm64(v)
    parts = [v]
    if text is not None:
        t = np.asarray(text, dtype=np.float64)
        if text_norm:
            t = l2norm64(t)
        parts.append(t)
    x = np.concatenate(parts, axis=1)
    if final_norm:
        x = l2norm64(x)
    return x.astype(np.float32)


def brand_embed(w, e, brand_ids=None, normalize=False):
    """brand[b, :] = (1/A) * sum_a W[id_b, a] * E[a, :]   (model.py:419-428 + :594).

    w [NB+1, A] (nn.Embedding table), e [A, D]."""
    w = np.asarray(w, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    if brand_ids is not None:
        w = w[np.asarray(brand_ids, dtype=np.int64)]
    out = (w @ e) / w.shape[1]
    if normalize:
        out = l2norm64(out)
    return out.astype(np.float32)
