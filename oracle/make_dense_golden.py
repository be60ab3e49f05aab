#!/usr/bin/env python
"""Generate the dense-scorer golden fixture (BASELINE config 3) by running the reference's OWN code in this container.

TEST INFRASTRUCTURE ONLY.  Imports /root/reference with sys.modules stubs for packages that are absent here
(SURVEY.md section 8c), builds the reference `models.modeling_rag.GPT2LMHeadModel` exactly as
scripts/train_retriever/train_retriever_UCI_13.sh configures it (n_layer=4, n_head=2, n_embed=512, seed 42, random
init — no checkpoint exists offline), embeds the UCI_13 pool (histories) and test queries with the mean-over-padded-
length rule of train/train_retriever.py:414-423,430-432, then evaluates the literal scoring lines :433-438 on CPU and
the reference writer save_index_score (:357-368, argsort forced stable).  Also runs get_train_query_time.py for the
pool time vector.  Output: tests/golden/dense_UCI13.npz (+ sha256 of the reference .gen files in the manifest).
"""
import hashlib
import json
import os
import runpy
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# per-dataset retriever configuration: scripts/train_retriever/train_retriever_{UCI_13,hepth,dialog}.sh
CONFIGS = {
    "UCI_13": dict(T="12", n_layer=4, n_head=2, n_embed=512, block=512, pool_rows=None, query_rows=None),
    "hepth": dict(T="11", n_layer=12, n_head=2, n_embed=256, block=1024, pool_rows=1500, query_rows=128),
    "dialog": dict(T="15", n_layer=2, n_head=2, n_embed=256, block=1024, pool_rows=1500, query_rows=128),
}


def main(ds="UCI_13"):
    cfg = CONFIGS[ds]
    T = cfg["T"]
    scratch = tempfile.mkdtemp(prefix="r4d_dense_golden_")
    os.makedirs(os.path.join(scratch, "resources", ds), exist_ok=True)
    shutil.copytree(os.path.join(REF, "resources", ds, T), os.path.join(scratch, "resources", ds, T))
    if ds == "hepth":
        shutil.copy(os.path.join(REF, "resources", "hepth", "node_features.npy"), os.path.join(scratch, "resources", "hepth"))
        torch.Tensor.cuda = lambda self, *a, **k: self   # utils/tokenizer.py:62 hard-codes .cuda(); no GPU here
    shutil.copytree(os.path.join(REF, "vocabs"), os.path.join(scratch, "vocabs"))
    os.chdir(scratch)

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, k):
            return _Any()
    stub("boto3")
    bc = stub("botocore")
    bc.config = stub("botocore.config", Config=_Any)
    bc.exceptions = stub("botocore.exceptions", ClientError=Exception)
    stub("tensorboardX", SummaryWriter=_Any)
    stub("ipdb", set_trace=lambda: None)
    stub("wandb", init=_Any(), log=_Any(), finish=_Any(), login=_Any())
    tg = stub("torch_geometric")
    tg.nn = stub("torch_geometric.nn", GCNConv=_Any, global_mean_pool=_Any, GATConv=_Any, SAGEConv=_Any,
                 global_max_pool=_Any, global_add_pool=_Any)
    tg.utils = stub("torch_geometric.utils", from_networkx=_Any, to_undirected=_Any)
    tg.data = stub("torch_geometric.data", Data=_Any, Batch=_Any)
    import transformers
    if not hasattr(transformers, "AdamW"):
        transformers.AdamW = torch.optim.AdamW
    sys.path.insert(0, REF)

    from models.modeling_rag import GPT2LMHeadModel
    from models import GPT2Config
    from transformers import GPT2Tokenizer
    from utils.tokenizer import get_model_tokenizer
    import train.train_retriever as tr
    from torch.nn.utils.rnn import pad_sequence

    args = types.SimpleNamespace(model_type="gpt2", config_name=None, model_name_or_path=None, cache_dir=None,
                                 n_head=cfg["n_head"], n_layer=cfg["n_layer"], n_embed=cfg["n_embed"], eta=0.8,
                                 gamma=0.4, beta=0.0, timestamp=T, dataset=ds, device="cpu",
                                 node_feat_file=os.path.join("resources", "hepth", "node_features.npy"))
    torch.manual_seed(42)
    model, tokenizer, _, args = get_model_tokenizer(args, {"gpt2": (GPT2Config, GPT2LMHeadModel, GPT2Tokenizer)})
    model.eval()

    def read(p):
        with open(p, encoding="utf-8") as f:
            return [ln for ln in f.read().splitlines() if len(ln) > 0 and not ln.isspace()]

    def embed(lines):
        # dataloader/retriever.py:23 (batch_encode_plus was removed in transformers 5: tokenizer(...) gives the same ids)
        ids = tokenizer(lines, add_special_tokens=True, max_length=cfg["block"], truncation="longest_first")["input_ids"]
        out = []
        for b0 in range(0, len(ids), 32):                      # per_gpu_eval_batch_size default 32, SequentialSampler
            batch = [torch.tensor(x, dtype=torch.long) for x in ids[b0:b0 + 32]]
            inputs = pad_sequence(batch, batch_first=True, padding_value=tokenizer.pad_token_id)
            with torch.no_grad():
                _, h = model(input_ids=inputs)                 # train/train_retriever.py:419
                out.append(torch.mean(h, dim=1))               # :420 mean over the PADDED length
        return torch.cat(out, dim=0)

    base = os.path.join("resources", ds, T)
    train_lines = [ln.split("<|pre|>")[0].strip() for ln in read(os.path.join(base, "train.link_prediction"))]  # :51
    test_lines = read(os.path.join(base, "test.link_prediction"))
    if cfg["pool_rows"]:      # larger datasets: a prefix keeps the fixture small; scoring parity is per pair
        train_lines, test_lines = train_lines[:cfg["pool_rows"]], test_lines[:cfg["query_rows"]]
    train_embeddings = embed(train_lines)
    test_embeddings = embed(test_lines)

    # literal scoring block :433-438 per eval batch of 32, on CPU; the reference writer with stable argsort
    orig_argsort = np.argsort
    np.argsort = lambda a, axis=-1, kind=None, order=None, **kw: orig_argsort(a, axis=axis, kind="stable", order=order)
    rows, steps = [], 0
    os.makedirs("out", exist_ok=True)
    for b0 in range(0, test_embeddings.shape[0], 32):
        h_egos = test_embeddings[b0:b0 + 32]
        h_egos_norm = h_egos / h_egos.norm(dim=1, keepdim=True)
        train_embeddings_norm = train_embeddings / train_embeddings.norm(dim=1, keepdim=True)
        dot_products = torch.matmul(h_egos_norm, train_embeddings_norm.t())
        dot_products = (dot_products + 1) / 2
        arr = dot_products.detach().cpu().numpy()
        tr.save_index_score(arr, "out/test_index.gen", "out/test_score.gen", steps)
        rows.append(arr)
        steps += 1
    np.argsort = orig_argsort
    ref_scores = np.concatenate(rows, axis=0)

    pool_time = np.zeros(0, dtype=np.float32)
    if ds == "UCI_13":
        # pool query times: unmodified get_train_query_time.py UCI_13 12
        argv = sys.argv
        sys.argv = ["get_train_query_time.py", "UCI_13", "12"]
        runpy.run_path(os.path.join(REF, "get_train_query_time.py"), run_name="__main__")
        sys.argv = argv
        pool_time = torch.load(os.path.join("resources", "UCI_13_train_query_time.pt")).numpy()

    def sha(p):
        return hashlib.sha256(open(p, "rb").read()).hexdigest()
    tag = {"UCI_13": "UCI13"}.get(ds, ds)
    np.savez_compressed(os.path.join(GOLD, f"dense_{tag}.npz"), pool_emb=train_embeddings.numpy(),
                        query_emb=test_embeddings.numpy(), ref_scores=ref_scores, pool_time=pool_time)
    man_path = os.path.join(GOLD, "manifest.json")
    man = json.load(open(man_path))
    man[f"dense_{tag}"] = {"model": f"reference models.modeling_rag.GPT2LMHeadModel, {cfg['n_layer']} layers/{cfg['n_head']} heads/"
                                    f"{cfg['n_embed']}-d, seed 42, random init; pool rows {len(train_lines)}, queries {len(test_lines)}",
                          "pool": list(train_embeddings.shape), "queries": list(test_embeddings.shape),
                          "test_index.gen": {"sha256": sha("out/test_index.gen"), "bytes": os.path.getsize("out/test_index.gen")},
                          "test_score.gen": {"sha256": sha("out/test_score.gen"), "bytes": os.path.getsize("out/test_score.gen")},
                          "score_range": [float(ref_scores.min()), float(ref_scores.max())],
                          "pool_time_range": [float(pool_time.min()), float(pool_time.max())] if pool_time.size else None,
                          "torch": torch.__version__, "transformers": transformers.__version__}
    json.dump(man, open(man_path, "w"), indent=1, sort_keys=True)
    print("dense golden written:", train_embeddings.shape, test_embeddings.shape, ref_scores.min(), ref_scores.max())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "UCI_13")
