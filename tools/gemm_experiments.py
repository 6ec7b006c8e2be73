"""Timing experiments on the gather-GEMM kernel (B200): isolates A gather, W copy, MMA via FSFB_GEMM_DEBUG."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from fullysparsefusion_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    res = {}
    for name, (m, koff, cin, cout, dens) in {"conv160k_128": (160000, 27, 128, 128, 0.37), "lin300k_131_128": (300000, 1, 131, 128, 1.0),
                                             "lin300k_128_128": (300000, 1, 128, 128, 1.0), "conv40k_256": (40000, 27, 256, 256, 0.37)}.items():
        a = torch.randn(m, cin, device=dev, generator=g)
        w = torch.randn(koff, cout, cin, device=dev, generator=g) * 0.05
        pw = ops.gemm_prepack(w)
        nbr = None
        if koff > 1:
            # spatially coherent neighbours: row r's offset-k neighbour is near r
            base = torch.arange(m, device=dev, dtype=torch.int64)[None, :] + torch.randint(-2000, 2000, (koff, 1), device=dev, generator=g)
            nbr = base.clamp(0, m - 1).to(torch.int32)
            nbr[torch.rand(koff, m, device=dev, generator=g) > dens] = -1
            nbr[13] = torch.arange(m, device=dev, dtype=torch.int32)
            nbr = nbr.contiguous()
        out = torch.empty(m, cout, device=dev)
        for _ in range(3):
            ops.gather_gemm(a, pw, nbr=nbr, act="relu", out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gather_gemm(a, pw, nbr=nbr, act="relu", out=out)
        e1.record()
        torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) / 10, 4)
    print(json.dumps(res))
else:
    for dbg in [int(x) for x in os.environ.get('FSFB_DBG_LIST', '0').split(',')]:
        env = dict(os.environ, FSFB_GEMM_DEBUG=str(dbg))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("debug", dbg, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
