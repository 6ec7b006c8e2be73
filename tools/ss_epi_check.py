"""Epilogue variants of the gather-GEMM against the CUDA-core cross-check: LN / affine / none x GELU / ReLU / none, odd widths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fullysparsefusion_b200 import ops
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
bad = 0
for rows in (50, 300, 5000):
    for cin, cout in ((128, 128), (1024, 128), (768, 1024), (32, 146), (128, 3), (128, 33), (16, 32), (133, 128)):
        for norm, act, bias in (("ln", "gelu", False), ("ln", None, True), ("affine", "relu", False), (None, "gelu", True), (None, None, True)):
            if norm == "ln" and cout > 256:
                continue
            a = ops.empty_rows(rows, cin, dev); a.copy_(T(rng.standard_normal((rows, cin)).astype(np.float32)))
            w = T((rng.standard_normal((cout, cin)) / np.sqrt(cin)).astype(np.float32))
            b = T(rng.standard_normal(cout).astype(np.float32)) if bias else None
            nw = T(rng.uniform(0.5, 1.5, cout).astype(np.float32)) if norm else None
            nb = T(rng.standard_normal(cout).astype(np.float32)) if norm else None
            pw = ops.gemm_prepack(w, keep_raw=True)
            kw = dict(bias=b, norm=norm, norm_w=nw, norm_b=nb, eps=1e-3, act=act)
            got = ops.gather_gemm(a, pw, **kw)
            want = ops.gather_gemm(a, pw, simt=True, **kw)
            err = float((got - want).abs().max() / want.abs().max())
            rowerr = (got - want).abs().max(1).values / want.abs().max()
            nbad = int((rowerr > 1e-4).sum())
            if err > 1e-4:
                bad += 1
                first = int(torch.nonzero(rowerr > 1e-4)[0])
                print(f"BAD rows={rows} {cin}x{cout} norm={norm} act={act} bias={bias}: err {err:.2e}, {nbad} bad rows, first {first}", flush=True)
print("EPI CHECK", "OK" if bad == 0 else f"{bad} failures")
