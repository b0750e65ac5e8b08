"""Host-side chunking of the ingest path (no GPU needed): every post lands in exactly one chunk, chunks respect the post
and row budgets, a post with more frames than the budget gets a chunk of its own."""
import numpy as np

from fancyrec_b200 import ingest


def _check(n_posts, row_ptr, chunk):
    bounds, max_rows = ingest.chunk_bounds(n_posts, row_ptr, chunk)
    assert bounds[0] == 0 and bounds[-1] == n_posts and all(b1 > b0 for b0, b1 in zip(bounds[:-1], bounds[1:]))
    assert all(b1 - b0 <= chunk for b0, b1 in zip(bounds[:-1], bounds[1:]))
    if row_ptr is not None:
        rows = [int(row_ptr[b1] - row_ptr[b0]) for b0, b1 in zip(bounds[:-1], bounds[1:])]
        budget = max(chunk, int(np.diff(row_ptr).max())) if n_posts else chunk
        assert max(rows + [0]) == max_rows <= budget
    return bounds


def test_unpooled_chunks():
    assert _check(10, None, 4) == [0, 4, 8, 10]
    assert _check(0, None, 4) == [0]
    assert _check(3, None, 100) == [0, 3]


def test_pooled_chunks_respect_row_budget():
    rs = np.random.RandomState(0)
    for trial in range(50):
        n = int(rs.randint(1, 400))
        counts = rs.randint(0, 40, n)                      # zero-frame posts allowed
        if trial % 5 == 0:
            counts[rs.randint(0, n)] = 500                 # one post larger than the budget
        row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        _check(n, row_ptr, int(rs.choice([1, 7, 64, 1000])))


def test_config5_shape():
    n, f = 100000, 32
    row_ptr = np.arange(n + 1, dtype=np.int64) * f
    bounds, max_rows = ingest.chunk_bounds(n, row_ptr, 131072)
    assert max_rows == 131072 and bounds[1] == 131072 // f and bounds[-1] == n
