"""Host -> device ingest of post features feeding the finalisation kernel.

The reference reads one frame per `open` + `seek` + `fromfile` + `tolist` (util/imgbigfile.py:19-57), averages the
frames of a post in the collate function (util/data_provider.py:40) and normalises on the device later
(evaluator.py:14-19): 94 posts/s.  Here the frame rows stay where they are (an np.memmap over feature.bin, a NumPy
array or a pinned torch tensor); posts are cut into chunks, each chunk's rows are staged through pinned memory,
copied on a second CUDA stream and finalised by ONE kernel pass (mean-pool -> per-branch norm -> concat -> row norm
-> bf16 / fp32) while the next chunk is in flight.  The 1.31 TB of config 5 never has to be resident: only the
finalised [NP, D] operand is.

Results are bit-identical to `ops.finalize_posts` on device-resident inputs (same kernel, rows are independent).
"""
import numpy as np
import torch

from . import _lib, ops


def _is_pinned_tensor(x):
    return isinstance(x, torch.Tensor) and not x.is_cuda and x.is_pinned()


def chunk_bounds(n_posts, row_ptr, chunk_posts):
    """Cut posts 0..n_posts into consecutive chunks of at most `chunk_posts` posts; with a CSR `row_ptr` (pooled posts) a
    chunk is also bounded in frame ROWS (chunk_posts rows, or one post if a single post has more) so that the staging
    buffers stay small.  Returns (bounds, max_rows): chunk c is posts bounds[c] .. bounds[c+1]-1."""
    bounds = [0]
    if row_ptr is None:
        while bounds[-1] < n_posts:
            bounds.append(min(n_posts, bounds[-1] + chunk_posts))
        return bounds, min(chunk_posts, n_posts)
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    row_budget = max(chunk_posts, int((row_ptr[1:] - row_ptr[:-1]).max())) if n_posts else chunk_posts
    while bounds[-1] < n_posts:
        p0 = bounds[-1]
        p1 = int(np.searchsorted(row_ptr, row_ptr[p0] + row_budget, side="right")) - 1
        bounds.append(min(n_posts, max(p0 + 1, min(p1, p0 + chunk_posts))))
    max_rows = max([int(row_ptr[b1] - row_ptr[b0]) for b0, b1 in zip(bounds[:-1], bounds[1:])] + [0])
    return bounds, max_rows


class _Stager:
    """Two pinned host buffers + two device buffers of `rows` x `cols` fp32 and the events that guard them."""

    def __init__(self, rows, cols, device, need_pinned):
        self.dev = [torch.empty((rows, cols), dtype=torch.float32, device=device) for _ in range(2)]
        self.host = [torch.empty((rows, cols), dtype=torch.float32, pin_memory=True) for _ in range(2)] \
            if need_pinned else None
        self.copied = [None, None]       # H2D of slot s finished (the pinned buffer may be refilled)
        self.consumed = [None, None]     # the finalise pass that read device slot s finished


def finalize_from_host(visual, text=None, row_ptr=None, row_idx=None, visual_norm=False, text_norm=False,
                       final_norm=True, want_f32=False, want_bf16=True, device=None, chunk_posts=131072,
                       out_f32=None, out_bf16=None):
    """A1-A3 from HOST features.  Same contract as ops.finalize_posts, but `visual` [rows, Dv] / `text` [NP, Dt] are
    host arrays (np.ndarray, np.memmap, torch CPU tensor; pinned tensors are copied from directly) and `row_ptr`
    [NP+1] / `row_idx` [total] are host integer arrays (CSR of frame rows per post; None = one row per post).
    Returns (out_f32 | None, out_bf16 | None) on `device`."""
    lib = _lib.load()
    device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    if isinstance(visual, torch.Tensor) and visual.is_cuda:
        raise ValueError("finalize_from_host takes host features; use ops.finalize_posts for device tensors")
    dv = int(visual.shape[1])
    dt = int(text.shape[1]) if text is not None else 0
    if row_ptr is not None:
        row_ptr = np.ascontiguousarray(np.asarray(row_ptr, dtype=np.int64))
        n_posts = len(row_ptr) - 1
        if row_idx is not None:
            row_idx = np.asarray(row_idx, dtype=np.int64)
    else:
        if row_idx is not None:
            raise ValueError("row_idx needs row_ptr")
        n_posts = int(visual.shape[0])
    if text is not None and int(text.shape[0]) != n_posts:
        raise ValueError("text has %d rows, expected %d" % (text.shape[0], n_posts))
    d = dv + dt
    ld = ops.round_up(d, 64)
    flags = (ops.VISUAL_NORM if visual_norm else 0) | (ops.TEXT_NORM if text_norm else 0) | \
            (ops.FINAL_NORM if final_norm else 0)
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty((n_posts, d), dtype=torch.float32, device=device)
    if want_bf16 and out_bf16 is None:
        out_bf16 = torch.empty((n_posts, ld), dtype=torch.bfloat16, device=device)
    if out_f32 is None and out_bf16 is None:
        raise ValueError("no output requested")
    if n_posts == 0:
        return out_f32, out_bf16

    chunk_posts = max(1, int(chunk_posts))
    bounds, max_rows = chunk_bounds(n_posts, row_ptr, chunk_posts)
    direct_v = _is_pinned_tensor(visual) and row_idx is None
    direct_t = text is None or _is_pinned_tensor(text)
    sv = _Stager(max(max_rows, 1), dv, device, not direct_v)
    st = _Stager(min(chunk_posts, n_posts), dt, device, not direct_t) if text is not None else None
    ptr_dev = [torch.empty(min(chunk_posts, n_posts) + 1, dtype=torch.int64, device=device) for _ in range(2)] \
        if row_ptr is not None else None
    ptr_host = [torch.empty(min(chunk_posts, n_posts) + 1, dtype=torch.int64, pin_memory=True) for _ in range(2)] \
        if row_ptr is not None else None
    vis_np = visual.numpy() if isinstance(visual, torch.Tensor) and not direct_v else visual
    txt_np = text.numpy() if isinstance(text, torch.Tensor) and not direct_t else text

    main = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device=device)
    copy.wait_stream(main)
    with torch.cuda.device(device):
        for ci, (p0, p1) in enumerate(zip(bounds[:-1], bounds[1:])):
            s = ci & 1
            npost = p1 - p0
            r0, r1 = (int(row_ptr[p0]), int(row_ptr[p1])) if row_ptr is not None else (p0, p1)
            nrows = r1 - r0
            # ---- stage on the host (pinned) -----------------------------------------------------
            if sv.copied[s] is not None:
                sv.copied[s].synchronize()               # the previous H2D out of this pinned slot is done
            if not direct_v and nrows:
                dst = sv.host[s].numpy()[:nrows]
                if row_idx is not None:
                    np.take(vis_np, row_idx[r0:r1], axis=0, out=dst)     # gather the chunk's frame rows in order
                else:
                    dst[...] = vis_np[r0:r1]
            if text is not None and not direct_t:
                st.host[s].numpy()[:npost] = txt_np[p0:p1]
            if row_ptr is not None:
                ptr_host[s].numpy()[:npost + 1] = row_ptr[p0:p1 + 1] - r0
            # ---- H2D on the copy stream ------------------------------------------------------------
            with torch.cuda.stream(copy):
                if sv.consumed[s] is not None:
                    copy.wait_event(sv.consumed[s])      # device slot still being finalised
                if nrows:
                    src = visual[r0:r1] if direct_v else sv.host[s][:nrows]
                    sv.dev[s][:nrows].copy_(src, non_blocking=True)
                if text is not None:
                    src = text[p0:p1] if direct_t else st.host[s][:npost]
                    st.dev[s][:npost].copy_(src, non_blocking=True)
                if row_ptr is not None:
                    ptr_dev[s][:npost + 1].copy_(ptr_host[s][:npost + 1], non_blocking=True)
                sv.copied[s] = torch.cuda.Event()
                sv.copied[s].record(copy)
            # ---- finalise on the main stream --------------------------------------------------------
            main.wait_event(sv.copied[s])
            rc = lib.frx_finalize_posts(
                sv.dev[s].data_ptr(), ptr_dev[s].data_ptr() if row_ptr is not None else 0, 0,
                st.dev[s].data_ptr() if text is not None else 0, npost, dv, dt, flags,
                out_f32[p0:p1].data_ptr() if out_f32 is not None else 0,
                out_bf16[p0:p1].data_ptr() if out_bf16 is not None else 0, ld, main.cuda_stream)
            _lib.check(rc, "frx_finalize_posts")
            sv.consumed[s] = torch.cuda.Event()
            sv.consumed[s].record(main)
    main.wait_stream(copy)
    # the staging tensors were used on both streams; keep them alive until the main stream has passed this point
    for t in sv.dev + (st.dev if st is not None else []) + (ptr_dev or []):
        t.record_stream(main)
    return out_f32, out_bf16
