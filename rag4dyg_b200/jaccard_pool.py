"""Retrieval pool resident in HBM + the top-K call a user makes against it (Jaccard scorer).

`JaccardPool` holds the state of one pool (shard): the bitsets (storage of record; the bitset kernels serve dense sets
and full matrices), the cardinalities and — when the universe fits (n_bits <= 65 535) — the postings index that the
sparse-set top-K path walks (r4d_postings_build).  It replaces what the reference rebuilds for every pair: `set(seq_j)`
of every pool sample inside co_occurrence_ratio (retrieval_data_annotation.py:12-13).

`topk` is the device-resident call; `HostTopK` is the same call for HOST buffers: the step's query id lists arrive in
pinned host memory, results land in pinned host memory; device->host copies run on a copy stream and overlap the
scoring of the next row range / the next step.
"""
import torch

from . import _lib, engine
from ._lib import R4DError

POSTINGS_MAX_BITS = 65535


class JaccardPool:
    def __init__(self, bits, pool_base=0, postings="auto"):
        """bits: BitsetMatrix of the pool (shard) on the GPU.  postings: "auto" | True | False."""
        self.bits = bits
        self.pool_base = int(pool_base)
        self.index = None
        want = postings is True or (postings == "auto" and bits.n_bits <= POSTINGS_MAX_BITS)
        if want:
            try:
                self.index = engine.build_postings(bits)
            except R4DError:
                if postings is True:
                    raise
        self._ws = None

    @classmethod
    def from_csr(cls, bit_pos, row_off, n_bits, device="cuda", pool_base=0, postings="auto"):
        """Host CSR id lists (duplicates allowed) -> pool state in HBM (H2D + set encoder + postings build)."""
        from . import set_encoder
        return cls(set_encoder.encode_csr(bit_pos, row_off, n_bits, device), pool_base, postings)

    @property
    def n_rows(self):
        return self.bits.n_rows

    @property
    def device(self):
        return self.bits.bits.device

    def workspace(self, nq, k, relay=False):
        """relay=True adds the staging room that lets packed lists leave for pinned host memory in whole blocks."""
        lib = _lib.load()
        need = (lib.r4d_jaccard_topk_postings_workspace_bytes(nq) if self.index is not None
                else lib.r4d_jaccard_topk_workspace_bytes(nq, self.n_rows, k))
        if relay and self.index is not None:
            need += lib.r4d_jaccard_topk_postings_relay_bytes(nq, k)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
        return self._ws

    def topk(self, q_ids, q_off, k, zero_diag=False, query_base=0, out=None, q_nnz=None):
        """Queries as CSR id lists on the device -> (inter, union, idx) int32 [nq, k], canonical order, GLOBAL indices.
        Sparse path (postings) when the index exists, else set encoder + bitset kernels.  q_off may be a row range of a
        larger CSR (offsets stay absolute into q_ids); q_nnz then tells how many ids the range holds (sizing hint)."""
        nq = q_off.numel() - 1
        if self.index is not None:
            return engine.jaccard_topk_postings(q_ids, q_off, self.index, k, zero_diag=zero_diag, query_base=query_base,
                                                pool_base=self.pool_base, workspace=self.workspace(nq, k), out=out,
                                                q_nnz=q_nnz)
        q = engine.encode_bitsets(q_ids, q_off, self.bits.n_bits)
        r = engine.jaccard_topk(q, self.bits, k, zero_diag=zero_diag, query_base=query_base, pool_base=self.pool_base,
                                workspace=self.workspace(nq, k))
        if out is not None:
            for o, t in zip(out, r):
                o.copy_(t)
            return out
        return r

    def topk_packed(self, q_ids, q_off, k, zero_diag=False, query_base=0, out=None, q_nnz=None):
        """`topk` with packed results: (pair = inter << 16 | |pool set| [nq, k], idx [nq, k], q_card = |query set| [nq]),
        8 bytes per entry instead of 12 (postings path only; `engine.unpack_topk` gives the three planes back)."""
        if self.index is None:
            raise R4DError("topk_packed needs the postings path (a pool with an index)")
        nq = q_off.numel() - 1
        to_host = out is not None and not out[0].is_cuda
        return engine.jaccard_topk_postings_packed(q_ids, q_off, self.index, k, zero_diag=zero_diag, query_base=query_base,
                                                   pool_base=self.pool_base, workspace=self.workspace(nq, k, relay=to_host),
                                                   out=out, q_nnz=q_nnz)


class GraphTopK:
    """One JaccardPool.topk call captured in a CUDA graph: replay() re-runs the whole launch sequence (counter memset +
    scoring kernels) with a single graph launch.  For short steps (a few thousand queries per GPU) the Python / ctypes
    launch path (~100 us) is longer than the kernels; the graph makes the step GPU-bound again.  The query buffers and
    the outputs are fixed device tensors: refresh their CONTENTS (copy_) between replays, never rebind them."""

    def __init__(self, pool, q_ids, q_off, k, zero_diag=False, query_base=0, out=None):
        if pool.index is None:
            raise R4DError("GraphTopK needs the postings path (a pool with an index)")
        nq = q_off.numel() - 1
        self.pool, self.q_ids, self.q_off, self.k = pool, q_ids, q_off, int(k)
        self.out = out if out is not None else tuple(torch.empty((nq, self.k), dtype=torch.int32, device=pool.device)
                                                     for _ in range(3))
        pool.workspace(nq, self.k)
        with torch.cuda.device(pool.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up outside the capture: lazy module load, shared-memory opt-ins
                pool.topk(q_ids, q_off, self.k, zero_diag=zero_diag, query_base=query_base, out=self.out)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                pool.topk(q_ids, q_off, self.k, zero_diag=zero_diag, query_base=query_base, out=self.out)
            # the captured launches point into the pool's workspace: keep THAT buffer alive even if the pool later moves
            # to a larger one (a bigger step, the relay room of host-bound packed lists)
            self._workspace = pool._ws

    def replay(self):
        self.graph.replay()
        return self.out


class HostTopK:
    """Host-buffer front end of JaccardPool.topk: submit(q_ids, q_off) enqueues H2D -> top-K -> results in pinned host
    memory and returns a ticket; result(ticket) waits for that step only; `depth` steps may be in flight.

    direct=True (default, postings path): the top-K kernel stores its [Q, K] lists STRAIGHT into the pinned host
    buffers (pinned memory is device-addressable under unified addressing), so the device->host transfer is spread over
    the kernel's run time as posted PCIe writes — there is no separate copy, and a step costs H2D + scoring.
    direct=False: results land in HBM and are copied out on a copy stream, in `chunks` row ranges so that the copy of
    one range overlaps the scoring of the next (also the path for pools without a postings index).
    packed=True (direct mode): the lists cross PCIe in the packed form of r4d_jaccard_topk_postings_packed — 8 bytes per
    entry instead of 12; result() then returns (pair, idx, q_card) and `HostTopK.unpack` turns them into the three
    planes on the host (score = inter / union is the same rational either way)."""

    def __init__(self, pool, k, max_queries, max_ids, depth=2, chunks=1, direct=True, packed=False):
        self.pool, self.k, self.depth, self.chunks = pool, int(k), int(depth), max(1, int(chunks))
        self.direct = bool(direct) and pool.index is not None
        self.packed = bool(packed)
        if self.packed and not self.direct:
            raise R4DError("HostTopK: packed results need direct mode (a pool with a postings index)")
        dev = pool.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(self.depth):
            self.slots.append({
                "ids": torch.empty((max(1, max_ids),), dtype=torch.int32, device=dev),
                "off": torch.empty((max_queries + 1,), dtype=torch.int64, device=dev),
                "out": None if self.direct else tuple(torch.empty((max_queries, self.k), dtype=torch.int32, device=dev)
                                                      for _ in range(3)),
                "host": ((torch.empty((max_queries, self.k), dtype=torch.int32).pin_memory(),
                          torch.empty((max_queries, self.k), dtype=torch.int32).pin_memory(),
                          torch.empty((max_queries,), dtype=torch.int32).pin_memory()) if self.packed else
                         tuple(torch.empty((max_queries, self.k), dtype=torch.int32).pin_memory() for _ in range(3))),
                "scored": [torch.cuda.Event() for _ in range(self.chunks)], "done": torch.cuda.Event(), "nq": 0, "busy": False,
            })
        self.step = 0

    def submit(self, q_ids, q_off, zero_diag=False, query_base=0):
        """q_ids int32 [nnz], q_off int64 [nq+1]: CPU tensors (pinned for an asynchronous copy)."""
        s = self.slots[self.step % self.depth]
        if s["busy"]:
            raise R4DError("HostTopK: more than `depth` steps in flight; collect result() first")
        nq, nnz = q_off.numel() - 1, q_ids.numel()
        if nq > s["off"].numel() - 1 or nnz > s["ids"].numel():
            raise R4DError("HostTopK: step larger than the buffers this object was created with")
        ids, off = s["ids"][:max(nnz, 1)], s["off"][:nq + 1]
        if nnz:
            ids[:nnz].copy_(q_ids, non_blocking=True)
        off.copy_(q_off, non_blocking=True)
        if self.direct:
            call = self.pool.topk_packed if self.packed else self.pool.topk
            call(ids, off, self.k, zero_diag=zero_diag, query_base=query_base, out=tuple(h[:nq] for h in s["host"]), q_nnz=nnz)
            s["done"].record()
        else:
            n_chunks = min(self.chunks, max(1, nq // 4096))            # small steps are not worth splitting
            bounds = [nq * c // n_chunks for c in range(n_chunks + 1)]
            for c in range(n_chunks):
                a, b = bounds[c], bounds[c + 1]
                # row range [a, b): the offsets stay absolute into `ids`, so only the offset and output views move
                self.pool.topk(ids, off[a:b + 1], self.k, zero_diag=zero_diag, query_base=query_base + a,
                               out=tuple(o[a:b] for o in s["out"]), q_nnz=int(q_off[b]) - int(q_off[a]))
                s["scored"][c].record()
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(s["scored"][c])
                    for h, o in zip(s["host"], s["out"]):
                        h[a:b].copy_(o[a:b], non_blocking=True)
            with torch.cuda.stream(self.copy_stream):
                s["done"].record()
        s["nq"], s["busy"] = nq, True
        ticket = self.step
        self.step += 1
        return ticket

    def result(self, ticket):
        s = self.slots[ticket % self.depth]
        s["done"].synchronize()
        s["busy"] = False
        if not self.direct:
            # the compute stream may reuse this slot's device buffers only after the copy has read them
            torch.cuda.current_stream().wait_event(s["done"])
        return tuple(h[:s["nq"]] for h in s["host"])

    @staticmethod
    def unpack(result):
        """(pair, idx, q_card) of a packed step -> (inter, union, idx) int32 [nq, k]."""
        return engine.unpack_topk(*result)

    def bytes_per_step(self, nq, nnz):
        return nnz * 4 + (nq + 1) * 8, (nq * self.k * 8 + nq * 4) if self.packed else nq * self.k * 12
