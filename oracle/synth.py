"""Seeded synthetic inputs shared by the golden generator and the tests
(TEST INFRASTRUCTURE ONLY).  Uses the legacy ``np.random.RandomState`` stream,
which NumPy guarantees to be bit-stable across versions, so the golden files
only need to store seeds and reference OUTPUTS, not inputs.
"""
import numpy as np


def labels(seed, n_posts, n_brands, empty_brands=()):
    """brands[j] in [0, n_brands); brands in ``empty_brands`` get no post."""
    rs = np.random.RandomState(seed)
    lab = rs.randint(0, n_brands, size=n_posts).astype(np.int64)
    if len(empty_brands):
        allowed = np.array([b for b in range(n_brands) if b not in set(empty_brands)], dtype=np.int64)
        lab = allowed[rs.randint(0, len(allowed), size=n_posts)]
    return lab


def gaussian(seed, rows, dim, scale=1.0):
    rs = np.random.RandomState(seed)
    return (rs.standard_normal((rows, dim)) * scale).astype(np.float32)


def planted_posts(seed, brand_emb, lab, noise=1.0, signal=0.25):
    """post_j = noise * N(0, 1) + signal * sqrt(D) * brand[label_j] / ||brand[label_j]||:
    a trained-model-like workload (the positive brand tends to rank high)."""
    rs = np.random.RandomState(seed)
    d = brand_emb.shape[1]
    x = rs.standard_normal((len(lab), d)).astype(np.float32) * np.float32(noise)
    bn = brand_emb / np.linalg.norm(brand_emb, axis=1, keepdims=True)
    x += np.float32(signal * np.sqrt(d)) * bn[lab].astype(np.float32)
    return x


def lattice(seed, rows, dim, nnz=1024):
    """Exact-lattice rows: entries in {-1, 0, +1} with exactly ``nnz``
    non-zeros, nnz a power of 4 => ||x|| = sqrt(nnz) exactly, x/||x|| = +-2^-m exact in
    bf16/tf32/fp32, every dot product an integer multiple of 1/nnz below 2^24: the score
    matrix is identical in any precision and any accumulation order, with many exact ties."""
    assert nnz <= dim and int(round(np.sqrt(nnz))) ** 2 == nnz
    rs = np.random.RandomState(seed)
    out = np.zeros((rows, dim), dtype=np.float32)
    for r in range(rows):
        cols = rs.permutation(dim)[:nnz]
        out[r, cols] = rs.randint(0, 2, size=nnz).astype(np.float32) * 2 - 1
    return out


def frames_csr(seed, n_posts, dim, fmin=1, fmax=8):
    """ResNet-avgpool-like non-negative frame rows, ragged: F_p ~ U{fmin..fmax}."""
    rs = np.random.RandomState(seed)
    counts = rs.randint(fmin, fmax + 1, size=n_posts)
    row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    frames = np.maximum(rs.standard_normal((int(row_ptr[-1]), dim)) * 0.5 + 0.3, 0).astype(np.float32)
    return frames, row_ptr


# ----------------------------------------------------------------------------
# named fixture inputs (shared by tests/golden/make_golden.py and the tests)
# ----------------------------------------------------------------------------
RANKING_CASES = {
    # name: (kind, NB, NP, D, A, seed, empty brands)
    "planted": ("planted", 6, 300, 64, 40, 101, (4,)),
    "gauss": ("gauss", 5, 257, 48, 16, 202, ()),
    "lattice": ("lattice", 8, 500, 256, 0, 303, ()),
    "lattice_small_k": ("lattice", 3, 37, 64, 0, 404, (1,)),
}


def ranking_inputs(name):
    kind, nb, npost, d, a, seed, empty = RANKING_CASES[name]
    lab = labels(seed, npost, nb, empty)
    if kind == "lattice":
        # brand table: A == D aspects with identity-like aspect matrix so that the brand
        # embedding is itself a lattice row (W = lattice, E = A * I  =>  mean_a W[b,a]E[a,:] = W[b,:])
        a = d
        w = np.zeros((nb + 1, a), dtype=np.float32)
        w[:nb] = lattice(seed + 1, nb, d, nnz=16 if d == 64 else 64)
        e = (np.eye(a, dtype=np.float32) * np.float32(a))
        posts = lattice(seed + 2, npost, d, nnz=16 if d == 64 else 64)
    else:
        w = gaussian(seed + 1, nb + 1, a)
        e = gaussian(seed + 2, a, d)
        if kind == "planted":
            brand = (w[:nb].astype(np.float64) @ e.astype(np.float64) / a).astype(np.float32)
            posts = planted_posts(seed + 3, brand, lab)
        else:
            posts = gaussian(seed + 3, npost, d)
    return nb, lab, w, e, posts


def loss_inputs(seed=505, b=24, d=32, n_ids=7):
    rs = np.random.RandomState(seed)
    ids = rs.randint(0, n_ids, size=b).astype(np.int64)
    brand = gaussian(seed + 1, b, d, 0.5)
    post = gaussian(seed + 2, b, d, 0.5)
    return ids, brand, post


