"""rag4dyg_b200 — B200-native (sm_100a) query-by-pool scoring + top-K for RAG4DyG.

Only the hot path lives here: `csrc/` (hand-written CUDA + the C ABI of include/r4d.h), `engine` (torch-tensor front
end of that ABI) and host-side mirrors of the reference's interface for the path
(`retrieval_data_annotation`, `dense_retrieval`, `sharded`).  No CPU fallback exists anywhere in this package.
"""
from . import _lib  # noqa: F401
from ._lib import R4DError  # noqa: F401

__version__ = "0.1.0"
