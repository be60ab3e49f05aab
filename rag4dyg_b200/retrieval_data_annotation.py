"""Drop-in for the reference's retrieval-pool annotation stage, computed on a B200 through libr4d.so.

Same surface as /root/reference/retrieval_data_annotation.py:
  * CLI      python retrieval_data_annotation.py <dataset> <timestep> <threshold>      (:111-113)
  * functions co_occurrence_ratio, get_input_seq, get_output_seq, get_inout_list, occurrence_matrix,
              save_train_annotation, save_index_score, save_score_file_train           (same positional arguments)
  * files    resources/<ds>/<T>/train_retrieval/{train,val,test}_{index,score}.retrieval and
             resources/train_generator/<ds>/<T>/train_gt_topk/train_{index,score}.gen  (:117-135)

What differs, on purpose:
  * every `np.argsort(-x)` of the reference is unstable and therefore machine-dependent on ties;
    this engine always emits the canonical order (score descending, pool index ascending) == kind='stable'.
  * all scoring / ranking / mining runs on the GPU; there is no CPU path (a missing library or device raises).
The random negative per positive still comes from the legacy global numpy RNG, called exactly as the reference calls
it (:79), so seeding `np.random.seed(s)` before `main()` reproduces a seeded reference run byte for byte.
"""
import os
import sys

import numpy as np
import torch

from . import engine, set_encoder, writers

_DEVICE = "cuda"


# ------------------------------------------------------------------------------------------------ parsing
def _between(text, start, end):
    # text after the first `start` (up to a second one, as str.split would cut), then up to the first `end`
    parts = text.split(start)
    if len(parts) < 2:
        raise IndexError(f"marker {start!r} not found")
    return parts[1].split(end)[0]


def get_input_seq(seq):
    """History tokens: ego id, `<|timeN|>` markers and neighbour ids, duplicates kept (reference :17-20)."""
    return [tok for tok in _between(seq, "<|history|>", "<|endofhistory|>").split(" ") if tok != ""]


def get_output_seq(seq):
    """Label (y) tokens; anything containing 'time' is dropped (reference :22-26)."""
    return [tok for tok in _between(seq, "<|pre|>", "<|endofpre|>").split(" ") if tok != "" and "time" not in tok]


def get_inout_list(data, gt):
    """(history lists, label lists) for parallel line lists (reference :28-34)."""
    in_list = [get_input_seq(data[n]) for n in range(len(data))]
    out_list = [get_output_seq(gt[n]) for n in range(len(data))]
    return in_list, out_list


# ------------------------------------------------------------------------------------------------ scoring
def _encode_pair(target, source):
    uni = set_encoder.Universe().add(source).add(target)
    same = target is source
    p = set_encoder.encode_sequences(source, uni, _DEVICE)
    q = p if same else set_encoder.encode_sequences(target, uni, _DEVICE)
    return q, p


def occurrence_matrix(target, source):
    """All-pairs Jaccard, float64 [len(target), len(source)] on the host (reference :36-41), computed by
    r4d_bitset_encode + r4d_jaccard_full."""
    if len(target) == 0 or len(source) == 0:
        return np.zeros((len(target), len(source)))
    q, p = _encode_pair(target, source)
    _, score = engine.jaccard_full(q, p, zero_diag=False, want_score=True)
    return score.cpu().numpy()


def co_occurrence_ratio(seq_i, seq_j):
    """|A & B| / |A | B| of two token lists; 0 when either is empty/None (reference :5-15)."""
    if type(seq_j) is not list:
        seq_j = [seq_j]
    if seq_i is None or seq_j is None:
        return 0
    if len(seq_i) == 0 or len(seq_j) == 0:
        return 0
    return float(occurrence_matrix([list(seq_i)], [seq_j])[0, 0])


def _to_device_f64(m):
    if isinstance(m, torch.Tensor):
        return m.to(device=_DEVICE, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(m, dtype=np.float64)).to(_DEVICE)


def _replay_choice(sizes):
    """np.random.choice(a) for consecutive arrays of the given lengths, on numpy's legacy GLOBAL RandomState: returns
    the chosen positions and leaves the global stream exactly where the reference's per-triplet calls (:79) would
    (r4d_mt19937_choice_replay; one native loop instead of one Python call per triplet)."""
    import ctypes
    from . import _lib
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    if sizes.size and int(sizes.min()) <= 0:
        raise ValueError("'a' cannot be empty unless no samples are taken")   # np.random.choice([]) in the reference
    st = np.random.get_state()
    if st[0] != "MT19937":
        raise _lib.R4DError(f"unsupported global bit generator {st[0]!r}")
    key = np.array(st[1], dtype=np.uint32)
    pos = ctypes.c_int32(int(st[2]))
    pick = np.empty(sizes.size, dtype=np.int32)
    _lib.check(_lib.load().r4d_mt19937_choice_replay(key.ctypes.data, ctypes.addressof(pos), sizes.ctypes.data, sizes.size,
                                                     pick.ctypes.data), "r4d_mt19937_choice_replay")
    np.random.set_state((st[0], key, pos.value, st[3], st[4]))
    return pick


def _write_triplets(save_file, save_file_score, n_rows, pos_rows, pos_cols, pos_scores, neg, n_neg, neg_scores):
    """Host tail of save_train_annotation: dialog truncation, np.random.choice replay, text output (:72-85).
    pos_* list the positives in row-major order; neg / neg_scores [n, neg_num], n_neg [n]."""
    is_dialog = "dialog" in dataset  # noqa: F821  module global set by main(), exactly like the reference (:73)
    pos_rows = np.asarray(pos_rows, dtype=np.int64)
    if is_dialog:                    # pos_indices[:4]: the first four positives of every row
        starts = np.searchsorted(pos_rows, np.arange(n_rows + 1))
        keep = (np.arange(pos_rows.size) - starts[pos_rows]) < 4
        pos_rows, pos_cols, pos_scores = pos_rows[keep], np.asarray(pos_cols)[keep], np.asarray(pos_scores)[keep]
    pick = _replay_choice(np.asarray(n_neg)[pos_rows])          # one np.random.choice per written triplet, in file order
    neg_i = np.asarray(neg)[pos_rows, pick]
    s_neg = np.asarray(neg_scores)[pos_rows, pick]
    writers.write_int_rows(save_file, np.stack([pos_rows, np.asarray(pos_cols, dtype=np.int64), neg_i.astype(np.int64)], axis=1))
    writers.write_triplet_scores(save_file_score, pos_rows, pos_scores, s_neg)
    print("Number of original instances:", n_rows)
    print("Number of positive samples:", int(pos_rows.size))


def _device_sampled_triplets(out_d, pos, neg, n_neg, save_file, save_file_score, seed):
    """sampler="counter": the whole triplet list is produced on the device (r4d_triplet_sample) and written by the
    native formatters.  Deterministic in `seed`, NOT the numpy stream."""
    n = out_d.shape[0]
    rows, cols = pos[:, 0].contiguous(), pos[:, 1].contiguous()
    if "dialog" in dataset:  # noqa: F821  (:73-74) keep the first 4 positives of every row
        start = torch.searchsorted(rows, torch.arange(n + 1, device=rows.device))
        keep = (torch.arange(rows.numel(), device=rows.device) - start[rows]) < 4
        rows, cols = rows[keep].contiguous(), cols[keep].contiguous()
    row_start = torch.searchsorted(rows, torch.arange(n + 1, device=rows.device)).contiguous()
    choice = engine.triplet_sample(rows, row_start, neg, n_neg, seed)
    if bool((choice < 0).any()):
        raise ValueError("'a' cannot be empty unless no samples are taken")  # np.random.choice([]) in the reference
    s_pos = out_d[rows, cols]
    s_neg = out_d[rows, choice.long()]
    trip = torch.stack([rows.int(), cols.int(), choice], dim=1).cpu().numpy()
    writers.write_int_rows(save_file, trip)
    sp, sn = s_pos.cpu().numpy(), s_neg.cpu().numpy()
    with open(save_file_score, "w") as g:
        for t in range(trip.shape[0]):
            g.write(f"{trip[t, 0]} {writers.fmt_str(sp[t])} {writers.fmt_str(sn[t])}\n")
    print("Number of original instances:", n)
    print("Number of positive samples:", trip.shape[0])


def _mine_and_write(out_d, in_d, save_file, save_file_score, threshold, neg_num, sampler="numpy", seed=0):
    n = out_d.shape[0]
    n_pos, neg, n_neg = engine.triplet_mine(out_d, in_d, threshold, neg_num)
    pos = torch.nonzero(out_d > threshold)  # row-major == np.where per row in ascending order (:54)
    if sampler == "counter":
        _device_sampled_triplets(out_d, pos, neg, n_neg, save_file, save_file_score, seed)
        return n_pos
    pos_scores = out_d[pos[:, 0], pos[:, 1]]
    neg_scores = torch.gather(out_d, 1, neg.clamp(min=0).to(torch.int64))
    pos_h = pos.cpu().numpy()
    _write_triplets(save_file, save_file_score, n, pos_h[:, 0], pos_h[:, 1], pos_scores.cpu().numpy(),
                    neg.cpu().numpy(), n_neg.cpu().numpy(), neg_scores.cpu().numpy())
    return n_pos


def save_train_annotation(scores_matrix, scores_matrix_in, save_file, save_file_score, threshold=0.8, neg_num=5,
                          sampler="numpy", seed=0):
    """Positive / negative triplets for the retriever (reference :43-85).  Matrices may be numpy or CUDA tensors.
    sampler="numpy" replays the reference's np.random.choice stream (parity); "counter" samples on the device."""
    _mine_and_write(_to_device_f64(scores_matrix), _to_device_f64(scores_matrix_in), save_file, save_file_score,
                    threshold, neg_num, sampler, seed)


def save_index_score(score_matrix, save_index_file, save_score_file):
    """Full descending ranking + raw score rows (reference :88-93)."""
    s = _to_device_f64(score_matrix)
    order = engine.rank_rows(s)
    writers.write_int_rows(save_index_file, order.cpu().numpy())
    writers.write_float_rows(save_score_file, s.cpu().numpy(), writers.fmt_str)


def save_score_file_train(scores_matrix, save_file_index, save_file_score, topk=10):
    """Top-k indices and scores per row (reference :97-103)."""
    s = _to_device_f64(scores_matrix)
    k = min(topk, s.shape[1])
    ts, ti = engine.topk_rows(s, k)
    writers.write_int_rows(save_file_index, ti.cpu().numpy())
    writers.write_float_rows(save_file_score, ts.cpu().numpy(), writers.fmt_str)


# ------------------------------------------------------------------------------------------------ CLI
def _read_lines(path):
    with open(path, "r") as f:
        return [line for line in f.read().splitlines() if (len(line) > 0 and not line.isspace())]


class _StageTimer:
    """Wall time per stage (device work is synchronised at stage boundaries only when timing is requested)."""

    def __init__(self, enabled):
        import time
        self.enabled, self.t, self.clock, self.stages = enabled, time.perf_counter(), time.perf_counter, {}

    def mark(self, name):
        if self.enabled:
            torch.cuda.synchronize()
            now = self.clock()
            self.stages[name] = self.stages.get(name, 0.0) + now - self.t
            self.t = now


def annotate(dataset_name, timestamp, threshold, neg_num=5, topk=10, timing=None):
    """The whole stage on the device (reference __main__, :109-200): bitsets stay in HBM, the train top-k comes
    from the fused scorer+top-K kernel (no [N, N] ranking), only what is written to disk crosses to the host.
    timing: optional dict that receives wall seconds per stage (parse / encode / score / mine+triplets / rank / write)."""
    global dataset
    dataset = dataset_name
    tm = _StageTimer(timing is not None)
    save_path = os.path.join("./resources/", dataset, str(timestamp), "train_retrieval")
    os.makedirs(save_path, exist_ok=True)
    save_path_gen = os.path.join("./resources/train_generator", dataset, str(timestamp), "train_gt_topk")
    os.makedirs(save_path_gen, exist_ok=True)

    base = os.path.join("resources", dataset, timestamp)
    train_data = _read_lines(os.path.join(base, "train.link_prediction"))
    test_data = _read_lines(os.path.join(base, "test.link_prediction"))
    test_gt = _read_lines(os.path.join(base, "test_gt.link_prediction"))
    val_data = _read_lines(os.path.join(base, "val.link_prediction"))
    val_gt = _read_lines(os.path.join(base, "val_gt.link_prediction"))

    train_in, train_out = get_inout_list(train_data, train_data)
    _, test_out = get_inout_list(test_data, test_gt)
    _, val_out = get_inout_list(val_data, val_gt)
    tm.mark("read+parse")

    # subsystem 1: set encoder.  One universe for label sets (train/test/val), one for history sets.
    uni_out = set_encoder.Universe().add(train_out).add(test_out).add(val_out)
    uni_in = set_encoder.Universe().add(train_in)
    b_train_out = set_encoder.encode_sequences(train_out, uni_out, _DEVICE)
    b_test_out = set_encoder.encode_sequences(test_out, uni_out, _DEVICE)
    b_val_out = set_encoder.encode_sequences(val_out, uni_out, _DEVICE)
    b_train_in = set_encoder.encode_sequences(train_in, uni_in, _DEVICE)
    tm.mark("universe+csr+encode")

    # subsystem 2 + triplets: positives and hard / fill negatives straight from the OUT and IN bitsets (both diagonals
    # zeroed as :172-173); no [N, N] matrix exists at any point
    m = engine.triplet_mine_bits(b_train_out, b_train_in, threshold, neg_num, zero_diag=True)
    tm.mark("triplet mining from bitsets (GPU)")
    _write_triplets(os.path.join(save_path, "train_index.retrieval"), os.path.join(save_path, "train_score.retrieval"),
                    b_train_out.n_rows, m["pos_row"].cpu().numpy(), m["pos_col"].cpu().numpy(),
                    writers.jaccard_scores_f64(m["pos_inter"].cpu().numpy(), m["pos_union"].cpu().numpy()),
                    m["neg"].cpu().numpy(), m["n_neg"].cpu().numpy(),
                    writers.jaccard_scores_f64(m["neg_inter"].cpu().numpy(), m["neg_union"].cpu().numpy()))
    tm.mark("numpy RNG replay + write triplets")

    for name, b in (("test", b_test_out), ("val", b_val_out)):
        _, s = engine.jaccard_full(b, b_train_out, zero_diag=False)
        tm.mark("jaccard val/test matrices (GPU)")
        order = engine.rank_rows(s)
        tm.mark("full ranking val/test (GPU)")
        # the rankings and scores become text in HBM (r4d_format_rows_device); only the files' bytes cross to the host
        writers.write_int_rows_device(os.path.join(save_path, f"{name}_index.retrieval"), order)
        writers.write_float_rows_device(os.path.join(save_path, f"{name}_score.retrieval"), s, writers.fmt_str)
        tm.mark("text assembly (GPU) + D2H + write val/test files")

    # subsystem 4: fused scorer + top-K for the generator's train_gt_topk (never ranks the [N, N] matrix)
    k = min(topk, b_train_out.n_rows)
    t_inter, t_union, t_idx = engine.jaccard_topk(b_train_out, b_train_out, k, zero_diag=True)
    writers.write_int_rows(os.path.join(save_path_gen, "train_index.gen"), t_idx.cpu().numpy())
    writers.write_float_rows(os.path.join(save_path_gen, "train_score.gen"),
                             writers.jaccard_scores_f64(t_inter.cpu().numpy(), t_union.cpu().numpy()), writers.fmt_str)
    tm.mark("fused top-K train_gt_topk (GPU) + write")
    if timing is not None:
        timing.update(tm.stages)
    print("Done!")


def main(argv=None):
    argv = sys.argv if argv is None else argv
    dataset_name = argv[1]
    timestamp = argv[2]
    threshold = float(argv[3])
    annotate(dataset_name, timestamp, threshold)


if __name__ == "__main__":
    main()
