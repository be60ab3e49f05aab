"""A/B probe (same box, alternating) for dense pair-kernel options.  Usage: python tools/ab_dense.py key v0 v1 [reps]
PREC=x3 in the environment runs the split-precision mode (hi + lo planes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag4dyg_b200 import _lib, engine
key, v0, v1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda"); n_pool, d, qs, k = 10_000_000, 768, 8192, 10
g = torch.Generator(device=dev).manual_seed(1)
x3 = os.environ.get("PREC", "") == "x3"
prec = engine.PREC_BF16X3 if x3 else engine.PREC_BF16
hi = torch.empty((n_pool, d), dtype=torch.bfloat16, device=dev)
lo = torch.empty((n_pool, d), dtype=torch.bfloat16, device=dev) if x3 else None
for a in range(0, n_pool, 500_000):
    pl = engine.dense_prepare(torch.randn((500_000, d), generator=g, device=dev), prec)
    hi[a:a + 500_000] = pl.hi
    if x3: lo[a:a + 500_000] = pl.lo
pool = engine.DensePlanes(hi, lo, d, d, prec)
pt = torch.rand(n_pool, generator=g, device=dev) * 110
q = engine.dense_prepare(torch.randn((qs, d), generator=g, device=dev), prec)
qt = torch.rand(qs, generator=g, device=dev) * 110
ws = torch.empty(_lib.load().r4d_dense_topk_workspace_bytes(qs, n_pool, k), dtype=torch.uint8, device=dev)
def run(steps):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(steps): engine.dense_topk(q, pool, k, engine.DENSE_COS_DECAY, qt, pt, 1e-4, workspace=ws)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / steps
for r in range(reps):
    for v in (v0, v1):
        _lib.set_option(key, v); run(2); ms = run(6 if x3 else 15)
        print(f"rep {r} {key}={v}: {ms:.2f} ms/step  {(3 if x3 else 1)*2*d*qs*n_pool/ms/1e9:.0f} TFLOP/s issued", flush=True)
