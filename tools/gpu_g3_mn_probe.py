"""Probe of gemm3x's in-place MN-major operand path: out = A . B^T with B stored transposed, structured inputs, error map."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fancyrec_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for (m, n, k, at, bt) in [(128, 128, 32, False, True), (128, 128, 32, True, False), (128, 128, 64, True, True), (512, 3072, 512, True, True)]:
        rs = np.random.RandomState(1)
        a = rs.standard_normal((m, k)).astype(np.float32)
        b = rs.standard_normal((n, k)).astype(np.float32)
        want = a.astype(np.float64) @ b.astype(np.float64).T
        ad = torch.from_numpy(np.ascontiguousarray(a.T) if at else a).to(dev)
        bd = torch.from_numpy(np.ascontiguousarray(b.T) if bt else b).to(dev)
        got = ops.matmul3x(ad, bd, a_transposed=at, b_transposed=bt).cpu().numpy().astype(np.float64)
        err = np.abs(got - want)
        print("m %d n %d k %d at %d bt %d: max err %.3e  |want| %.2f  zeros %d / %d" % (m, n, k, at, bt, err.max(), np.abs(want).max(), int((got == 0).sum()), got.size))
        if err.max() > 1e-3:
            bad = err > 1e-3
            print("   bad rows", np.unique(np.nonzero(bad)[0])[:40], "bad cols", np.unique(np.nonzero(bad)[1])[:40])
            # which permutation of B columns / A rows would explain it?
            if not at and bt and n <= 128:
                g = got[:, :]
                # got[:, j] should equal want[:, perm[j]] for some perm
                perm = [int(np.argmin(np.abs(want - g[:, j:j + 1]).sum(0))) for j in range(n)]
                print("   col perm", perm[:64])
            if at and not bt and m <= 128:
                perm = [int(np.argmin(np.abs(want - got[i:i + 1, :]).sum(1))) for i in range(m)]
                print("   row perm", perm[:64])


if __name__ == "__main__":
    main()
