"""CPU: r4d_mt19937_choice_replay (a HOST function of libr4d.so) reproduces numpy's legacy global
np.random.choice(a) call by call — same picks, same generator state afterwards — which is what lets the drop-in write
train_index.retrieval byte-identically to a seeded reference run (retrieval_data_annotation.py:79) without one Python
call per triplet."""
import numpy as np
import pytest

from rag4dyg_b200 import build
from rag4dyg_b200 import retrieval_data_annotation as rda


def setup_module(module):
    build.build_lib()


@pytest.mark.parametrize("seed", [0, 1, 2024])
def test_replay_equals_numpy_choice_and_leaves_the_same_state(seed):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, 9, size=4000).astype(np.int32)
    sizes[::7] = 1                                     # one-element arrays consume no randomness
    np.random.seed(seed)
    np.random.random(5)                                # mid-block position
    ref = np.array([int(np.random.choice(np.arange(m))) for m in sizes])
    tail_ref = np.random.randint(0, 1 << 30, size=4)
    np.random.seed(seed)
    np.random.random(5)
    got = rda._replay_choice(sizes)
    tail = np.random.randint(0, 1 << 30, size=4)
    assert np.array_equal(ref, got) and np.array_equal(tail_ref, tail)


def test_replay_crosses_state_regeneration_and_large_arrays():
    np.random.seed(3)
    sizes = np.full(3000, 1000, dtype=np.int32)        # > 624 draws: the MT state is regenerated several times
    sizes[5] = 2 ** 20 + 3
    ref = np.array([int(np.random.choice(np.arange(m))) for m in sizes])
    np.random.seed(3)
    assert np.array_equal(ref, rda._replay_choice(sizes))


def test_empty_candidate_list_raises_like_numpy():
    with pytest.raises(ValueError):
        rda._replay_choice(np.array([3, 0, 2], dtype=np.int32))
    assert rda._replay_choice(np.zeros(0, dtype=np.int32)).size == 0
