"""Drop-in for the reference's query-time extraction (get_train_query_time.py): one float32 per pool
sequence = `ts` of the ego node's last interaction strictly before its final observed time step (or, when there is
none, its last interaction in that step), divided by the dataset's time scale.  These are the `p_time` values of the
dense scorer's exp(-lambda*|dt|) epilogue (train/train_retriever.py:50-55, :290-291).

Host code: the reference spends ~11 s in per-sequence pandas filters (`get_query_time`, :17-25); here the edge list is
sorted once and every ego id is answered by two binary searches.  The same rule can be applied to val/test queries
(the reference defines times for the train pool only).
"""
import os

import numpy as np
import pandas as pd
import torch

# get_train_query_time.py:47-54
SCALES = {"UCI_13": 3600 * 24, "hepth": 3600 * 24 * 30, "dialog": 1, "wikiv2": 3600 * 24, "enron": 1, "reddit": 1}


class EdgeTimes:
    """Undirected interaction list grouped by node (get_train_query_time.py:7-15 adds the reversed edges)."""

    def __init__(self, csv_path):
        data = pd.read_csv(csv_path, usecols=["u", "i", "ts", "timestamp"])
        u = np.concatenate([data["u"].to_numpy(), data["i"].to_numpy()])
        ts = np.concatenate([data["ts"].to_numpy(), data["ts"].to_numpy()]).astype(np.float64)
        step = np.concatenate([data["timestamp"].to_numpy(), data["timestamp"].to_numpy()])
        order = np.lexsort((ts, u))
        self.u, self.ts, self.step = u[order], ts[order], step[order]

    def query_time(self, node, timestamp):
        """get_query_time (:17-25): rows of `node` with step <= T-2; last ts before the node's max step, else the
        last ts within it."""
        lo, hi = np.searchsorted(self.u, node, "left"), np.searchsorted(self.u, node, "right")
        keep = self.step[lo:hi] <= int(timestamp) - 2
        ts, step = self.ts[lo:hi][keep], self.step[lo:hi][keep]
        if ts.size == 0:
            raise IndexError("single positional indexer is out-of-bounds")  # what the reference's .iloc[-1] raises
        max_step = step.max()
        before = ts[step < max_step]
        return before.max() if before.size else ts[step == max_step].max()


def ego_ids(lines):
    """queryID = int(seq.split('<|history|>')[1].split(' ')[1])  (get_train_query_time.py:36)."""
    return [int(seq.split("<|history|>")[1].split(" ")[1]) for seq in lines]


def get_query_time_all(data_name, timestamp, root=".", data_file="train.link_prediction", save=True):
    """float32 tensor [N] for the sequences of `data_file`; saved as resources/<ds>_train_query_time.pt like :41."""
    base = os.path.join(root, "resources", data_name, str(timestamp))
    edges = EdgeTimes(os.path.join(base, f"ml_{data_name}.csv"))
    with open(os.path.join(base, data_file)) as f:
        lines = [ln for ln in f.read().splitlines() if len(ln) > 0 and not ln.isspace()]
    times = [edges.query_time(q, timestamp) / SCALES[data_name] for q in ego_ids(lines)]
    out = torch.tensor(times, dtype=torch.float)
    if save:
        torch.save(out, os.path.join(root, "resources", data_name + "_train_query_time.pt"))
    return out


def main(argv=None):
    import sys
    argv = sys.argv if argv is None else argv
    get_query_time_all(argv[1], argv[2])


if __name__ == "__main__":
    main()
