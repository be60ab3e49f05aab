"""Host half of the set encoder: token lists -> CSR of bit positions.

The reference compares tokens as *strings* inside Python sets (retrieval_data_annotation.py:12-13): history sets hold
the ego id, neighbour ids and `<|timeN|>` markers (get_input_seq keeps them, :17-20).  Any injective token -> bit
position map reproduces the set algebra exactly, so the universe is simply "every distinct token seen", numbered in
first-appearance order; with the shipped data that is the dataset vocabulary plus the T+1 time tokens.
The device half (scatter-OR + popcount) is r4d_bitset_encode.
"""
import itertools

import numpy as np
import torch

from . import engine


def _flatten(seqs):
    """(flat token list, row lengths int64 [n]) of list[list[token]] with co_occurrence_ratio's argument quirks."""
    rows = [_as_list(seq) for seq in seqs]
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    return list(itertools.chain.from_iterable(rows)), lens


class Universe:
    """Injective token -> bit-position map shared by every list that will be compared (first-appearance order)."""

    def __init__(self):
        self.pos = {}

    def add(self, seqs):
        flat, _ = _flatten(seqs)
        pos = self.pos
        # dict.fromkeys: the distinct tokens in order of appearance, at C speed; the Python loop only sees DISTINCT tokens
        for tok in dict.fromkeys(flat):
            if tok not in pos:
                pos[tok] = len(pos)
        return self

    @property
    def n_bits(self):
        return max(1, len(self.pos))


def _as_list(seq):
    # co_occurrence_ratio wraps a non-list seq_j in a list (:6-9) and treats None as empty (:10)
    if seq is None:
        return []
    if type(seq) is not list:
        return [seq]
    return seq


def to_csr(seqs, universe):
    """list[list[token]] -> (bit_pos int32 [nnz], row_off int64 [n+1]) numpy arrays (duplicates kept).
    The per-token dict lookups run inside map() / np.fromiter (no Python-level loop body)."""
    flat, lens = _flatten(seqs)
    row_off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=row_off[1:])
    if not flat:
        return np.zeros(0, dtype=np.int32), row_off
    return np.fromiter(map(universe.pos.__getitem__, flat), dtype=np.int32, count=len(flat)), row_off


def encode_csr(bit_pos, row_off, n_bits, device="cuda"):
    """Host CSR (numpy or CPU tensors, ideally pinned) -> BitsetMatrix on `device` (H2D + r4d_bitset_encode)."""
    bp = torch.as_tensor(bit_pos, dtype=torch.int32)
    ro = torch.as_tensor(row_off, dtype=torch.int64)
    if bp.numel() == 0:
        bp = torch.zeros(1, dtype=torch.int32)  # keep a valid device pointer for empty inputs
    return engine.encode_bitsets(bp.to(device, non_blocking=True), ro.to(device, non_blocking=True), n_bits)


def encode_sequences(seqs, universe, device="cuda"):
    bit_pos, row_off = to_csr(seqs, universe)
    return encode_csr(bit_pos, row_off, universe.n_bits, device)
