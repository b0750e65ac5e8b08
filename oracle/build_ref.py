"""Recipe for oracle/_ref (TEST / BENCH INFRASTRUCTURE ONLY -- never imported by fancyrec_b200/).

The reference is pure Python, so "building" it means byte-compiling the modules of the hot path from the sources
where they lie under /root/reference into oracle/_ref/*.ref (CPython code objects in .pyc format; the extension is
ours because snapshot tools tend to drop *.pyc).  No reference source is copied into
this repository: oracle/_ref/ is git-ignored, holds only compiled code objects, and travels to the GPU box like our own
built libfrx_b200.so.  There it lets bench.py time the UNMODIFIED reference (`evaluator.test_post_ranking`,
evaluator.py:85-143, with the reference's own BrandAspects, model.py:406-428) on the box's host cores, and lets the
tests cross-check the oracle port against it.

    python oracle/build_ref.py            # -> oracle/_ref/{evaluator,model,loss,loss_ctrs}.ref, oracle/_ref/util/*.ref
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("FRX_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ["evaluator.py", "model.py", "loss.py", "loss_ctrs.py",
           "util/__init__.py", "util/constant.py", "util/util.py", "util/ndcg.py", "util/metric.py",
           "util/imgbigfile.py", "util/wordbigfile.py"]


def build(verbose=False):
    """Byte-compile the reference modules into oracle/_ref.  Returns False (and leaves any existing build alone)
    when the reference tree is not present -- the GPU box only uses the prebuilt files."""
    if not os.path.isdir(REF_SRC):
        return False
    for rel in MODULES:
        src = os.path.join(REF_SRC, rel)
        dst = os.path.join(OUT, rel[:-3] + ".ref")     # <module>.ref where <module>.py would be, loaded by _RefFinder
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile="reference/" + rel, doraise=True, optimize=0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print("compiled", src, "->", dst)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d" % sys.version_info[:2])
    return True


def available():
    ver = os.path.join(OUT, "PYTHON_VERSION")
    return os.path.exists(ver) and open(ver).read().strip() == "%d.%d" % sys.version_info[:2]


def load():
    """Import the byte-compiled reference modules.  Returns a namespace with evaluator, model, loss, loss_ctrs, ndcg,
    metric.  Shims applied before import, both semantic no-ops for the reference's pinned versions: np.asfarray
    (removed in NumPy 2; util/ndcg.py:37 calls it) restored with NumPy 1.21's definition."""
    if not available():
        raise RuntimeError("oracle/_ref is not built for this interpreter: run `python oracle/build_ref.py` where "
                           "/root/reference exists")
    import importlib
    import types

    import numpy as np
    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "util" or k.startswith("util.")}
    for k in saved_mods:
        del sys.modules[k]
    finder = _RefFinder()
    sys.meta_path.insert(0, finder)
    try:
        ns = types.SimpleNamespace()
        ns.evaluator = importlib.import_module("evaluator")
        ns.model = importlib.import_module("model")
        ns.loss = importlib.import_module("loss")
        ns.loss_ctrs = importlib.import_module("loss_ctrs")
        ns.ndcg = importlib.import_module("util.ndcg")
        ns.metric = importlib.import_module("util.metric")
    finally:
        sys.meta_path.remove(finder)
    return ns


class _RefFinder:
    """Meta-path finder for the byte-compiled reference modules (top-level names evaluator / model / loss / loss_ctrs and
    the package util.*), active only while load() imports them."""

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        import importlib.util
        rel = name.replace(".", os.sep)
        pkg = os.path.join(OUT, rel, "__init__.ref")
        mod = os.path.join(OUT, rel + ".ref")
        if os.path.exists(pkg):
            loader = importlib.machinery.SourcelessFileLoader(name, pkg)
            return importlib.util.spec_from_file_location(name, pkg, loader=loader,
                                                          submodule_search_locations=[os.path.join(OUT, rel)])
        if os.path.exists(mod):
            loader = importlib.machinery.SourcelessFileLoader(name, mod)
            return importlib.util.spec_from_file_location(name, mod, loader=loader)
        return None


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref built" if ok else "reference tree %s not present; nothing built" % REF_SRC)
