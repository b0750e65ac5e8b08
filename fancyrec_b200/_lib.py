"""ctypes binding of libfrx_b200.so (C ABI in include/frx.h).

There is NO CPU fallback: if the shared library has not been built, importing any
compute entry point raises, and every compute call fails loudly on a box without an
sm_100 GPU (frx_device_check).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfrx_b200.so")
STAMP_PATH = os.path.join(_HERE, "libfrx_b200.stamp")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "frx.h")
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-cudart", "static"]


def sources():
    """Every translation unit of libfrx_b200.so (csrc/*.cu) and the headers they include."""
    names = sorted(os.listdir(CSRC))
    return [n for n in names if n.endswith(".cu")], [n for n in names if n.endswith(".cuh")]


def source_hash():
    """Hash of everything libfrx_b200.so is built from; __graft_entry__.build() writes it next to the library and
    load() refuses a library whose stamp differs (a stale binary would silently run old kernels)."""
    import hashlib
    h = hashlib.sha256()
    cu, cuh = sources()
    for name in cu + cuh:
        h.update(name.encode())
        h.update(open(os.path.join(CSRC, name), "rb").read())
    h.update(open(HEADER, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()

c_i32, c_i64, c_f32, c_sz, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/frx.h one to one
SIGNATURES = {
    "frx_abi_version": (c_i32, []),
    "frx_last_error": (ctypes.c_char_p, []),
    "frx_device_check": (c_i32, [c_i32]),
    "frx_finalize_posts": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "frx_finalize_posts_bounded": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "frx_brand_embed_workspace_bytes": (c_sz, [c_i32, c_i32, c_i32]),
    "frx_brand_embed": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_sz, c_vp]),
    "frx_score_topk_workspace_bytes": (c_sz, [c_i32, c_i64, c_i32, c_i32]),
    "frx_score_topk": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp,
                               c_vp, c_i64, c_vp, c_sz, c_vp]),
    "frx_score_dense": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "frx_score_count": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "frx_score_topk_tf32": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp,
                                    c_vp, c_i64, c_vp, c_sz, c_vp]),
    "frx_score_dense_tf32": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "frx_score_count_tf32": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "frx_softmax_pool": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "frx_split_tf32x3": (c_i32, [c_vp, c_i64, c_i32, c_i64, c_i32, c_vp, c_vp]),
    "frx_topk_merge_strided": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_vp, c_vp, c_i32, c_vp]),
    "frx_topk_merge": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp]),
    "frx_label_stats": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "frx_rank_from_topk": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "frx_reduce_shard_stats": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "frx_missing_thresholds": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "frx_pack_rank_stats": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "frx_group_positives": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "frx_auc_rows": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "frx_brand_train_fwd": (c_i32, [c_vp, c_i64, c_vp, c_i32, c_i32, c_i32, ctypes.c_uint64, c_vp, c_vp]),
    "frx_brand_train_bwd": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_i32, c_i32, c_i32, ctypes.c_uint64, c_vp, c_vp, c_vp]),
    "frx_brand_dropout_mask": (c_i32, [c_i32, c_i32, c_i32, ctypes.c_uint64, c_vp, c_vp]),
    "frx_linear_workspace_bytes": (c_sz, [c_i32, c_i32, c_i32]),
    "frx_linear": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_sz, c_vp]),
    "frx_matmul3x": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp, c_i64, c_vp, c_sz, c_vp]),
    "frx_metric_scores": (c_i32, [c_vp, c_i64, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "frx_triplet_workspace_bytes": (c_sz, [c_i32, c_i32]),
    "frx_triplet_fwd_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "frx_vsepp_fwd_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "frx_contrastive_workspace_bytes": (c_sz, [c_i32, c_i32, c_i32]),
    "frx_contrastive_fwd_bwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i32, c_f32, c_f32, c_i32,
                                        c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "frx_crossclr_workspace_bytes": (c_sz, [c_i32, c_i32]),
    "frx_crossclr_fwd_bwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_f32, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "frx_lab_workspace_bytes": (c_sz, [c_i32, c_i32]),
    "frx_lab_fwd_bwd": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "frx_normalize_rows": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp]),
    "frx_set_cta_pairs": (c_i32, [c_i32]),
    "frx_set_cluster": (c_i32, [c_i32]),
    "frx_probe_enable": (c_i32, [c_i32]),
    "frx_probe_read": (c_i32, [c_vp, c_i32]),
}

_lib = None


class FrxError(RuntimeError):
    pass


def load():
    """Load libfrx_b200.so (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FrxError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(fancyrec_b200 has no CPU fallback)" % LIB_PATH)
    if os.environ.get("FRX_SKIP_STAMP_CHECK") != "1":
        have = open(STAMP_PATH).read().strip() if os.path.exists(STAMP_PATH) else "(no stamp)"
        want = source_hash()
        if have != want:
            raise FrxError("%s is stale: built from sources %s, tree is %s -- rebuild with "
                           "`python -c 'import __graft_entry__ as g; g.build()'`" % (LIB_PATH, have[:12], want[:12]))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.frx_abi_version() != 1:
        raise FrxError("libfrx_b200.so ABI version %d, expected 1" % lib.frx_abi_version())
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().frx_last_error().decode("utf8", "replace")
        raise FrxError("%s failed (code %d): %s" % (what, rc, msg))
