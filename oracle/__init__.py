"""CPU oracle for the brand x post scoring + ranking hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain NumPy restatement of the
reference algorithm (pinskyrobin/FancyRec: evaluator.py, util/ndcg.py,
util/metric.py, loss.py, loss_ctrs.py, model.py BrandAspects/l2norm and the
frame mean-pool of util/data_provider.py).  It exists to CHECK the CUDA path.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``fancyrec_b200/`` imports it, and the product path raises when the CUDA
library is missing instead of falling back to this code.

Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py)
against golden vectors produced by importing the reference itself in the build
container (tests/golden/make_golden.py, run with /root/reference mounted) and
against the doctest values in the reference's util/ndcg.py:15-27,54-65.
``oracle/_ref`` (git-ignored, built by ``oracle/build_ref.py`` where /root/reference exists) holds the reference's own
hot-path modules byte-compiled to sourceless .pyc files: the unmodified reference, runnable on the GPU box's host cores.
"""
