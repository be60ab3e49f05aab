import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch.distributed as dist
import bench
from rag4dyg_b200 import _lib
from rag4dyg_b200.jaccard_pool import JaccardPool
rank, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if "RANK" in os.environ: dist.init_process_group("nccl", device_id=dev)
pi, po = bench.synth_sets(100000, 1, 2.2); qi, qo = bench.synth_sets(20000, 2, 2.2)
pool = JaccardPool.from_csr(pi, po, bench.V_BITS, dev)
dq, do = qi.to(dev), qo.to(dev)
pool.topk(dq, do, 10); torch.cuda.synchronize()
print(rank, "prev", _lib.set_option("kernel_timing", 1), flush=True)
print(rank, "read0", _lib.profile_read("jaccard_postings"), flush=True)
for i in range(3): pool.topk(dq, do, 10)
torch.cuda.synchronize()
print(rank, "read1", _lib.profile_read("jaccard_postings"), _lib.last_error(), flush=True)
