"""Row-sharded candidate screening with the Tanimoto cross-kernel (BASELINE.json configs[3]).

Caller side of the hot path: the reference screens a held-out library one candidate per call
(notebooks/Bayesian Optimisation Over Molecules.ipynb cell 5 -> gauche/gp.py:169-226, i.e.
``K(x*, X) @ alpha``).  Here each rank (one process per GPU) owns a contiguous block of the
candidate rows, streams ``K(X*_block, X_train)`` through the tcgen05 Gram kernel into ping-pong
buffers, reduces each block to posterior-mean scores, keeps a local top-k, and the ranks exchange
ONLY their k ``(score, global index)`` records in one all-gather (NCCL over NVLink; gloo on CPU
in the tests).  The train operand is replicated (100k x 2048 bits = 205 MB as u8).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

from . import _lib
from .io import PackedLibrary
from .packed import PackedFingerprints, from_packed_library, pack, padded_k, stream_ptr

# dense host blocks gk_host_pack_bits can read
_HOST_DTYPES = {torch.float32: _lib.GK_F32, torch.float64: _lib.GK_F64, torch.uint8: _lib.GK_U8}


def bind_host_to_gpu_numa(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs lookup), so pinned staging
    buffers are first-touched on GPU-local memory.  With one process per GPU on a two-socket host this keeps
    every rank's H2D stream off the inter-socket link.  Returns the node, or None if it cannot be determined."""
    import os

    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_bounds(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of ``rank``: sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(n_total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rowdot(K: torch.Tensor, alpha: torch.Tensor, bias: float = 0.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``scores[i] = bias + sum_j K[i, j] * alpha[j]`` (fp32) via ``gk_rowdot_f32``."""
    assert K.is_cuda and K.dtype == torch.float32 and K.ndim == 2 and K.stride(1) == 1
    alpha = alpha.to(device=K.device, dtype=torch.float32).contiguous()
    n, m = K.shape
    assert alpha.numel() == m
    if out is None:
        out = torch.empty((n,), dtype=torch.float32, device=K.device)
    with torch.cuda.device(K.device):
        rc = _lib.load().gk_rowdot_f32(K.data_ptr(), K.stride(0) if n > 1 else max(m, 1), n, m,
                                       alpha.data_ptr(), float(bias), out.data_ptr(), stream_ptr(K.device))
    _lib.check(rc, "gk_rowdot_f32")
    return out


def topk(scores: torch.Tensor, k: int, index_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact deterministic top-k (descending, ties to the smaller index) via ``gk_topk_f32``."""
    assert scores.is_cuda and scores.dtype == torch.float32 and scores.ndim == 1 and scores.is_contiguous()
    n = scores.numel()
    k = min(k, n)
    lib = _lib.load()
    ws_bytes = lib.gk_topk_workspace(n, k)
    if ws_bytes < 0:
        raise ValueError(f"top-k supports 1 <= k <= 16384, got {k}")
    ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.int64, device=scores.device)
    vals = torch.empty((k,), dtype=torch.float32, device=scores.device)
    idx = torch.empty((k,), dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        rc = lib.gk_topk_f32(scores.data_ptr(), n, k, index_base, vals.data_ptr(), idx.data_ptr(),
                             ws.data_ptr(), ws_bytes, stream_ptr(scores.device))
    _lib.check(rc, "gk_topk_f32")
    return vals, idx


def merge_topk(vals: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Final selection over the gathered ``[world, k_local]`` records (tiny: host-side plumbing).
    Ranks own ascending index blocks and each list is already (score desc, index asc), so a stable
    descending sort of the concatenation keeps ties in global-index order on every rank."""
    v = vals.reshape(-1)
    i = idx.reshape(-1)
    valid = i >= 0
    v = torch.where(valid, v, torch.full_like(v, float("-inf")))
    order = torch.sort(v, descending=True, stable=True).indices[:k]
    return v[order], i[order]


def gather_topk(vals: torch.Tensor, idx: torch.Tensor, k: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """ONE all-gather of every rank's local top-k records — 12 bytes each: (score fp32, global index int64) packed
    as three int32 words — the path's only collective, then the identical merge on every rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_topk(vals, idx, k)
    world = dist.get_world_size(group)
    rec = torch.empty((k, 3), dtype=torch.int32, device=vals.device)
    have = vals.numel()
    rec[:have, 0] = vals.to(torch.float32).view(torch.int32)
    rec[:have, 1:] = idx.to(torch.int64).view(torch.int32).view(have, 2)
    if have < k:      # pad so that every rank contributes the same record count (a rank may own < k candidates)
        rec[have:, 0] = torch.tensor(float("-inf"), dtype=torch.float32).view(torch.int32)
        rec[have:, 1:] = -1
    all_rec = torch.empty((world * k * 3,), dtype=torch.int32, device=vals.device)   # flat: accepted by nccl and gloo
    dist.all_gather_into_tensor(all_rec, rec.view(-1), group=group)
    all_rec = all_rec.view(world * k, 3)
    all_v = all_rec[:, 0].contiguous().view(torch.float32)
    all_i = all_rec[:, 1:].contiguous().view(torch.int64).view(-1)
    return merge_topk(all_v.view(world, k), all_i.view(world, k), k)


class TanimotoScreener:
    """Posterior-mean screening of a candidate shard against a replicated train set.

    ``train``: dense ``[M, D]`` CUDA tensor or ``PackedFingerprints``; ``alpha``: ``[M]`` weights
    (``K_train^{-1}(y - mean)``, produced by the untouched GPyTorch solve); ``bias``: constant mean.
    """

    def __init__(self, train, alpha: torch.Tensor, bias: float = 0.0, block_rows: int = 16384,
                 engine: str = "auto", fused: bool = True, host_pack="auto", host_threads: int = 0):
        self.train = train if isinstance(train, PackedFingerprints) else pack(train)
        if self.train.kind == "real":
            raise ValueError("screening runs on bit / count fingerprints")
        self.device = self.train.device
        self.alpha = alpha.to(device=self.device, dtype=torch.float32).contiguous()
        self.bias = float(bias)
        self.block_rows = int(block_rows)
        self.engine = engine
        self.fused = fused          # K·alpha inside the Gram epilogue: the Gram block is never written
        self._buffers = None
        self._partials = None
        # host-resident candidates: "auto" | True | False (see _host_block); threads 0 = cpu_count / LOCAL_WORLD_SIZE
        self.host_pack = host_pack
        self.host_threads = int(host_threads)
        self._host_stage = None
        self._lib_stage = None
        self._dense_stage = None
        self._host_pack_choice = None if host_pack == "auto" else bool(host_pack)
        self._hp = {"n": 0, "t_pack_row": None, "t_dense_row": None, "dense_ev": None, "choice": None, "decisions": 0}
        self.host_pack_stats = None
        self.host_bytes_copied = 0
        self.last_operand_kind = None   # gk_tanimoto_rowdot operand kind of the most recent fused launch (0 u8, 1 e2m1, 2 bits)

    def _gram_buffers(self, rows: int):
        m = self.train.rows
        if self._buffers is None or self._buffers[0].shape[0] < rows:
            self._buffers = [torch.empty((rows, m), dtype=torch.float32, device=self.device) for _ in range(2)]
        return self._buffers

    def _scores_fused(self, cand: PackedFingerprints, out: torch.Tensor) -> torch.Tensor:
        """One launch per row block: tcgen05 Gram tiles are multiplied by alpha in the epilogue; only
        per-row partial sums leave the SM (``gk_tanimoto_rowdot``).  Bits: e2m1 operands on kind::mxf4
        (``engine="bx"``: the packed bit words themselves, expanded inside the SM); counts: u8 on kind::i8."""
        lib = _lib.load()
        binary = cand.kind == "binary" and self.train.kind == "binary"
        m = self.train.rows
        if binary and self.engine == "bx":      # packed bit words expanded inside the SM: no e2m1 copy of the library in HBM
            kind, a_all, b_op, kb = 2, cand.bits, self.train.bits, cand.kp // 8
            lda, ldb = (a_all.stride(0) if cand.rows > 1 else kb // 4) * 4, (b_op.stride(0) if m > 1 else kb // 4) * 4
        elif binary and self.engine == "bxh":   # hybrid: candidates as (permuted) bit words -> tensor memory, train set as e2m1
            kind, a_all, b_op, kb = 3, cand.bitsp(), self.train.q4(), cand.k4_bytes
            lda, ldb = (a_all.stride(0) if cand.rows > 1 else kb // 16), kb
        elif binary:
            kind, a_all, b_op, kb = 1, cand.q4(), self.train.q4(), cand.k4_bytes
            lda = ldb = kb
        else:
            kind, a_all, b_op, kb = 0, cand.q8, self.train.q8, cand.kp
            lda = ldb = kb
        slabs = lib.gk_tanimoto_rowdot_slabs(m, kind)
        rows_max = min(self.block_rows, cand.rows)
        ldp = (rows_max + 3) // 4 * 4
        if self._partials is None or self._partials.shape[0] < slabs or self._partials.shape[1] < ldp:
            self._partials = torch.empty((slabs, ldp), dtype=torch.float32, device=self.device)
        part = self._partials
        stream = stream_ptr(self.device)
        with torch.cuda.device(self.device):
            for r0 in range(0, cand.rows, self.block_rows):
                r1 = min(cand.rows, r0 + self.block_rows)
                if kind == 3:
                    rc = lib.gk_tanimoto_rowdot_hybrid(
                        a_all[r0:r1].data_ptr(), lda, cand.rowsq[r0:r1].data_ptr(), r1 - r0,
                        b_op.data_ptr(), ldb, self.train.rowsq.data_ptr(), m, kb // 16, kb,
                        self.alpha.data_ptr(), 1e-6, self.bias, part.data_ptr(), part.stride(0),
                        out[r0:r1].data_ptr(), stream)
                else:
                    rc = lib.gk_tanimoto_rowdot(
                        a_all[r0:r1].data_ptr(), lda, cand.rowsq[r0:r1].data_ptr(), r1 - r0,
                        b_op.data_ptr(), ldb, self.train.rowsq.data_ptr(), m, kb, kind,
                        self.alpha.data_ptr(), 1e-6, self.bias, part.data_ptr(), part.stride(0),
                        out[r0:r1].data_ptr(), stream)
                _lib.check(rc, "gk_tanimoto_rowdot")
        self.last_operand_kind = kind
        return out

    # ------------------------------------------------------------------ candidates that live in host memory
    def _host_threads(self) -> int:
        import os
        if not self.host_threads:
            local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
            self.host_threads = max(1, (os.cpu_count() or 1) // max(local, 1))
        return int(self.host_threads)

    HOST_PROBE_PERIOD = 256      # blocks between re-measurements of the two host routes

    def _host_route(self, mode) -> Tuple[bool, Optional[str]]:
        """Which way the next dense 0/1 host block travels: (host-pack?, what to time).  ``host_pack="auto"`` MEASURES
        both routes on real blocks of the running job — so contention between the ranks of a node (host cores and
        DRAM bandwidth for the packer, the shared PCIe / host-memory path for dense copies) is part of the numbers —
        and re-measures every HOST_PROBE_PERIOD blocks: blocks 0-1 are packed on the host (the second, warm, call is
        timed), block 2 is shipped dense with CUDA events around its copy, block 3 decides."""
        if mode is True or mode is False:
            return bool(mode), None
        hp = self._hp
        c = hp["n"] % self.HOST_PROBE_PERIOD
        first = hp["n"] < self.HOST_PROBE_PERIOD
        hp["n"] += 1
        if c == 0:
            return True, (None if first else "pack")      # very first call: thread pool start-up, untouched pages
        if c == 1:
            return True, "pack"
        if c == 2:
            return False, "dense"
        if c == 3 and hp["dense_ev"] is not None:
            e0, e1, rows = hp["dense_ev"]
            e1.synchronize()
            hp["dense_ev"] = None
            hp["t_dense_row"] = e0.elapsed_time(e1) * 1e-3 / max(rows, 1)
            if hp["t_pack_row"] is not None:
                hp["choice"] = hp["t_pack_row"] <= hp["t_dense_row"]
                hp["decisions"] += 1
                self._host_pack_choice = hp["choice"]
                self.host_pack_stats = {
                    "pack_us_per_krow": hp["t_pack_row"] * 1e9, "dense_copy_us_per_krow": hp["t_dense_row"] * 1e9,
                    "threads": self._host_threads(), "chosen": "host-pack" if hp["choice"] else "dense",
                    "decisions": hp["decisions"], "method": "both routes timed on live blocks (contended), re-probed every "
                                                            f"{self.HOST_PROBE_PERIOD} blocks"}
        return (True if hp["choice"] is None else hp["choice"]), None

    def _host_block(self, block: torch.Tensor):
        """Bring one dense host block ``[rows <= block_rows, D]`` to the device as ``PackedFingerprints``.

        A dense fp32 block moves at PCIe speed (~50 GB/s: 2.7 ms for 16 384 x 2048); packing it to 256 B per
        molecule on the host cores first (``gk_host_pack_bits``: thread pool + AVX-512 / AVX2, ~120 GB/s on 16 cores)
        leaves 4 MB to copy.  ``host_pack="auto"`` measures both on live blocks (``_host_route``); a block that is
        not exactly 0/1 is always shipped dense.  Input preparation only: the Gram runs on the GPU."""
        import ctypes, time
        rows, dim = block.shape
        code = _HOST_DTYPES.get(block.dtype)
        mode = self.host_pack
        usable = (code is not None and block.stride(1) == 1 and rows > 0 and self.train.kind == "binary")
        use_pack, probe = self._host_route(mode) if usable else (False, None)
        if usable and use_pack:
            kp = padded_k(dim)
            words = kp // 32
            st = self._host_stage
            if st is None or st["words"] != words or st["rows"] < rows:
                cap = max(rows, min(self.block_rows, 1 << 20))
                # per staging slot: pinned [bits | popcounts] in ONE buffer (one H2D copy), its device twin, an event
                st = {"words": words, "rows": cap, "i": 0,
                      "pin": [torch.empty((cap * (words + 1),), dtype=torch.int32).pin_memory() for _ in range(2)],
                      "dev": [torch.empty((cap * (words + 1),), dtype=torch.int32, device=self.device) for _ in range(2)],
                      "ev": [torch.cuda.Event(), torch.cuda.Event()], "used": [False, False]}
                self._host_stage = st
            i = st["i"]
            st["i"] = i ^ 1
            if st["used"][i]:
                st["ev"][i].synchronize()           # the copy that last read this pinned buffer has finished
            flag = ctypes.c_int32(0)
            pin = st["pin"][i]
            pop_ptr = pin.data_ptr() + rows * words * 4       # popcounts right behind the rows*words bit words
            t0 = time.perf_counter()
            rc = _lib.load().gk_host_pack_bits(block.data_ptr(), code, rows, dim, block.stride(0),
                                               pin.data_ptr(), words, pop_ptr,
                                               ctypes.byref(flag), self._host_threads())
            t_pack = time.perf_counter() - t0
            _lib.check(rc, "gk_host_pack_bits")
            if probe == "pack":
                t_row = t_pack / rows
                prev = self._hp["t_pack_row"]
                # two timed calls per probe period: keep the better one of the current period
                self._hp["t_pack_row"] = t_row if (prev is None or self._hp["n"] % self.HOST_PROBE_PERIOD == 1) else min(prev, t_row)
            if not flag.value:
                with torch.cuda.device(self.device):
                    n_i32 = rows * (words + 1)
                    dev = st["dev"][i][:n_i32]
                    dev.copy_(pin[:n_i32], non_blocking=True)            # bits and popcounts in one copy
                    st["ev"][i].record()
                    st["used"][i] = True
                    self.host_bytes_copied += n_i32 * 4
                    return PackedFingerprints.from_bits(dev[:rows * words].view(torch.uint8).view(rows, words * 4), dim,
                                                        clean=True, popcnt=dev[rows * words:])
        # dense route: the copy runs on a side stream into one of two device buffers, so the H2D of block i+1 overlaps
        # the kernels of block i (scores() enqueues and returns; `_dense_done` releases the buffer afterwards)
        self.host_bytes_copied += block.numel() * block.element_size()
        ds = self._dense_stage
        if ds is None or ds["dtype"] != block.dtype or ds["dim"] != dim or ds["rows"] < rows:
            cap = max(rows, min(self.block_rows, 1 << 20))
            ds = {"dtype": block.dtype, "dim": dim, "rows": cap, "i": 0, "stream": torch.cuda.Stream(device=self.device),
                  "dev": [torch.empty((cap, dim), dtype=block.dtype, device=self.device) for _ in range(2)],
                  "copied": [torch.cuda.Event(enable_timing=True) for _ in range(2)],
                  "start": [torch.cuda.Event(enable_timing=True) for _ in range(2)],
                  "consumed": [torch.cuda.Event(), torch.cuda.Event()], "used": [False, False], "busy": None}
            self._dense_stage = ds
        i = ds["i"]
        ds["i"] = i ^ 1
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            with torch.cuda.stream(ds["stream"]):
                if ds["used"][i]:
                    ds["stream"].wait_event(ds["consumed"][i])       # the kernels that read this buffer have finished
                dev = ds["dev"][i][:rows]
                ds["start"][i].record(ds["stream"])
                dev.copy_(block, non_blocking=True)
                ds["copied"][i].record(ds["stream"])
            cur.wait_event(ds["copied"][i])
        ds["used"][i] = True
        ds["busy"] = i
        if probe == "dense":
            self._hp["dense_ev"] = (ds["start"][i], ds["copied"][i], rows)
        return pack(dev)

    def _dense_done(self) -> None:
        """The kernels reading the dense staging buffer of the last host block are enqueued: mark it reusable."""
        ds = self._dense_stage
        if ds is not None and ds["busy"] is not None:
            ds["consumed"][ds["busy"]].record(torch.cuda.current_stream(self.device))
            ds["busy"] = None

    def _scores_library(self, lib: PackedLibrary, out: Optional[torch.Tensor]) -> torch.Tensor:
        """Stream a (memory-mapped) ``io.PackedLibrary`` shard: per block the packed rows are copied into one of two
        pinned staging buffers (the only host work: a memcpy of 256 B per molecule, page faults of the map included),
        sent to the device and scored; the copy of block i+1 overlaps the kernels of block i."""
        import numpy as np

        n = lib.rows
        if out is None:
            out = torch.empty((n,), dtype=torch.float32, device=self.device)
        nb, nc = lib.bits.shape[1], lib.n_counts
        width = nb + nc
        rows_max = max(1, min(self.block_rows, n))
        st = self._lib_stage
        if st is None or st["width"] != width or st["rows"] < rows_max:
            st = {"width": width, "rows": rows_max, "i": 0, "used": [False, False],
                  "pin": [torch.empty((rows_max, width), dtype=torch.uint8).pin_memory() for _ in range(2)],
                  "dev": [torch.empty((rows_max, width), dtype=torch.uint8, device=self.device) for _ in range(2)],
                  "ev": [torch.cuda.Event(), torch.cuda.Event()]}
            self._lib_stage = st
        for r0, blk in lib.blocks(self.block_rows):
            i = st["i"]
            st["i"] = i ^ 1
            if st["used"][i]:
                st["ev"][i].synchronize()            # the H2D copy that last read this pinned buffer has finished
            rows = blk.rows
            pin = st["pin"][i][:rows]
            host = pin.numpy()
            np.copyto(host[:, :nb], blk.bits)
            if nc:
                np.copyto(host[:, nb:], blk.counts)
            with torch.cuda.device(self.device):
                dev = st["dev"][i][:rows]
                dev.copy_(pin, non_blocking=True)
                st["ev"][i].record()
                st["used"][i] = True
                self.host_bytes_copied += rows * width
                cand = from_packed_library(dev[:, :nb], lib.n_bits, dev[:, nb:] if nc else None)
            self.scores(cand, out=out[r0:r0 + rows])
        return out

    def scores(self, candidates, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``K(candidates, train) @ alpha + bias`` for this rank's shard, streamed in row blocks.
        ``candidates``: a CUDA tensor, ``PackedFingerprints``, a dense HOST tensor (see ``_host_block``) or an
        ``io.PackedLibrary`` (packed wire / on-disk format, possibly memory-mapped)."""
        from .ops import _launch_tanimoto, resolve_engine

        if isinstance(candidates, PackedLibrary):
            return self._scores_library(candidates, out)

        if isinstance(candidates, torch.Tensor) and not candidates.is_cuda:
            if candidates.ndim != 2:
                raise ValueError(f"scores expects [rows, D] candidates, got shape {tuple(candidates.shape)}")
            n = candidates.shape[0]
            if out is None:
                out = torch.empty((n,), dtype=torch.float32, device=self.device)
            for r0 in range(0, n, self.block_rows):       # host packing / H2D of block i+1 overlaps the GPU work of block i
                r1 = min(n, r0 + self.block_rows)
                self.scores(self._host_block(candidates[r0:r1]), out=out[r0:r1])
                self._dense_done()
            return out
        cand = candidates if isinstance(candidates, PackedFingerprints) else pack(candidates)
        if cand.kind == "real":
            raise ValueError("screening runs on bit / count fingerprints")
        n = cand.rows
        if out is None:
            out = torch.empty((n,), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        if self.fused:
            return self._scores_fused(cand, out)
        kind = "binary" if (cand.kind == "binary" and self.train.kind == "binary") else "count"
        eng = resolve_engine("tanimoto", kind, self.engine)
        lib = _lib.load()
        bufs = self._gram_buffers(min(self.block_rows, n))
        stream = stream_ptr(self.device)
        with torch.cuda.device(self.device):
            for bi, r0 in enumerate(range(0, n, self.block_rows)):
                r1 = min(n, r0 + self.block_rows)
                kbuf = bufs[bi & 1][: r1 - r0]
                _launch_tanimoto(lib, eng, cand.rows_slice(r0, r1), self.train, kbuf, _lib.GK_F32, 1e-6, stream)
                rowdot(kbuf, self.alpha, self.bias, out=out[r0:r1])
        return out

    def screen(self, candidates, k: int, index_base: int = 0, group=None):
        """Local scores -> local top-k -> one all-gather -> identical global top-k on every rank."""
        s = self.scores(candidates)
        if s.numel() == 0:
            vals = torch.empty((0,), dtype=torch.float32, device=self.device)
            idx = torch.empty((0,), dtype=torch.int64, device=self.device)
        else:
            vals, idx = topk(s, k, index_base)
        return gather_topk(vals, idx, k, group)


def screen_sharded(score_fn: Callable[[int, int], torch.Tensor], n_total: int, k: int, group=None):
    """Backend-agnostic driver used by the multi-process tests: ``score_fn(start, stop)`` returns the
    scores of this rank's candidate rows; selection + exchange are the production code paths
    (``merge_topk`` / ``gather_topk``), with the local top-k taken by a stable sort when the scores
    live on the CPU (gloo tests) and by ``gk_topk_f32`` on CUDA."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, stop = shard_bounds(n_total, world, rank)
    s = score_fn(start, stop)
    if s.numel() == 0:
        vals, idx = s.new_empty((0,)), torch.empty((0,), dtype=torch.int64, device=s.device)
    elif s.is_cuda:
        vals, idx = topk(s.contiguous(), k, start)
    else:
        order = torch.sort(s, descending=True, stable=True)
        vals, idx = order.values[:k], order.indices[:k] + start
    return gather_topk(vals, idx, k, group)
