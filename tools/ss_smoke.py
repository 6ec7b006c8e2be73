"""Quick smoke of the gather-GEMM through ops.gather_gemm: a few Linear / conv shapes against the CUDA-core cross-check
(fsfb_gather_gemm_simt).  Prints one line per case; run under `timeout` during kernel bring-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fullysparsefusion_b200 import ops, _capi
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cases = [(1, 8, 16, 1), (128, 32, 128, 1), (1000, 131, 128, 1), (5000, 128, 33, 1), (333, 10, 131, 1), (200, 768, 1024, 1),
         (40000, 128, 128, 1), (3000, 64, 64, 27), (20000, 128, 128, 27), (1500, 256, 256, 27)]
for rows, cin, cout, koff in cases:
    a = T(rng.standard_normal((rows, cin)).astype(np.float32))
    w = T((rng.standard_normal((koff, cout, cin)) / np.sqrt(cin * max(1, koff // 4))).astype(np.float32))
    b = T(rng.standard_normal(cout).astype(np.float32))
    nbr = None
    if koff > 1:
        n = rng.integers(0, rows, (koff, rows)).astype(np.int32)
        n[rng.random((koff, rows)) < 0.7] = -1
        n[13] = np.arange(rows)
        nbr = T(n)
    pw = ops.gemm_prepack(w if koff > 1 else w[0], keep_raw=True)
    got = ops.gather_gemm(a, pw, nbr=nbr, bias=b, act="relu")
    torch.cuda.synchronize()
    want = ops.gather_gemm(a, pw, nbr=nbr, bias=b, act="relu", simt=True)
    err = float((got - want).abs().max() / want.abs().max())
    print(f"rows={rows} cin={cin} cout={cout} koff={koff}: max err / max = {err:.2e}", flush=True)
    assert err < 1e-4, err
cnt = __import__("ctypes").c_uint(0)
_capi.check(_capi.load().fsfb_gemm_f16_overflows(__import__("ctypes").byref(cnt)), "overflows")
print("overflow launches:", cnt.value)
print("SMOKE OK")
