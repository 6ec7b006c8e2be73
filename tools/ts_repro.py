import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fullysparsefusion_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
m, koff, cin, cout = 300, 27, 64, 32
a = torch.randn(m, cin, device=dev, generator=g)
w = torch.randn(koff, cout, cin, device=dev, generator=g) * 0.05
nbr = torch.randint(-1, m, (koff, m), device=dev, generator=g, dtype=torch.int32)
pw = ops.gemm_prepack(w, keep_raw=True)
out = ops.gather_gemm(a, pw, nbr=nbr)
torch.cuda.synchronize()
ref = ops.gather_gemm(a, pw, nbr=nbr, simt=True)
print("max err", (out - ref).abs().max().item())
