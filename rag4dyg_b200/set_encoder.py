"""Host half of the set encoder: token lists -> CSR of bit positions.

The reference compares tokens as *strings* inside Python sets (retrieval_data_annotation.py:12-13): history sets hold
the ego id, neighbour ids and `<|timeN|>` markers (get_input_seq keeps them, :17-20).  Any injective token -> bit
position map reproduces the set algebra exactly, so the universe is simply "every distinct token seen", numbered in
first-appearance order; with the shipped data that is the dataset vocabulary plus the T+1 time tokens.
The device half (scatter-OR + popcount) is r4d_bitset_encode.
"""
import numpy as np
import torch

from . import engine


class Universe:
    """Injective token -> bit-position map shared by every list that will be compared."""

    def __init__(self):
        self.pos = {}

    def add(self, seqs):
        pos = self.pos
        for seq in seqs:
            for tok in _as_list(seq):
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
    """list[list[token]] -> (bit_pos int32 [nnz], row_off int64 [n+1]) numpy arrays (duplicates kept)."""
    pos = universe.pos
    row_off = np.zeros(len(seqs) + 1, dtype=np.int64)
    flat = []
    for i, seq in enumerate(seqs):
        s = _as_list(seq)
        flat.extend(pos[t] for t in s)
        row_off[i + 1] = len(flat)
    return np.asarray(flat, dtype=np.int32), row_off


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
