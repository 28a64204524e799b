#!/usr/bin/env python
"""bench.py — Tanimoto cross-kernel throughput (Gpairs/s @2048-bit) on 1..8 B200.

Workload (BASELINE.json configs[3], the configuration the metric's "1/2/4/8 B200" refers to): BO candidate
screening, 100 000 train x 1 000 000 candidate 2048-bit fingerprints, candidate rows sharded over the GPUs.
One STEP = one pass of the hot path over one candidate row block per GPU:
K(X*_block [16 384 x 2048], X_train [100k x 2048]) -> fp32 Gram block written to HBM once.  Weak scaling: every
rank processes its own block per step; no data-path collective.

One JSON line (rank 0).  Keys beyond the base contract:
  value / roofline   the Gram kernel (tcgen05 kind::mxf4 on packed e2m1 operands resident in HBM, Gram written to HBM);
                     HBM-bound: 4 B/pair + operands once vs the measured copy bandwidth.  --engine bx: the bit-expanding
                     variant (packed bits all the way into the SM)
  sustained          the same launch repeated for >= 2 s with its own clock record
  roofline_fused     the kernel the e2e path runs (K·alpha fused into the epilogue, Gram never written): tensor-bound,
                     issued MAC-ops vs the MEASURED kind::mxf4 peak (tools/umma_microbench.cu, profiles/)
  e2e                the public call TanimotoScreener.scores(host fp32 block) -> scores on the host, >= 8 distinct
                     pinned host blocks in rotation; variants for the other host formats; host_bound = what the host
                     memory system permits for the dense-fp32 format (measured, all ranks concurrently)
  strong_scaling     the REAL config: 1 000 000 candidates (packed wire format, device resident) x 100k train sharded
                     over the ranks, fused scores + per-rank top-k + ONE all-gather inside the timed region
  strong_scaling_cfg5  configs[4]: K_xz of 1 000 000 fingerprints x 4096 inducing points, rows sharded over the ranks, + K_zz
  configs            (N = 1) cfg2 4200^2 fp32 Gram, cfg3 MinMax 50k^2, cfg5 K_xz shard + K_zz: CUDA-event times
  cpu_baseline       the reference's torch-CPU algorithm (oracle/torch_port.py) on the host cores, same inputs

`--impl reference` times that CPU path alone with the same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D_BITS = 2048
N_TRAIN = 100_000
N_CAND_TOTAL = 1_000_000
METRIC = "tanimoto_gram_gpairs_per_s_2048bit"
UNIT = "Gpairs/s"
# kind::mxf4.block_scale / kind::i8 dense rate of one B200 of this pool, measured by tools/umma_microbench.cu with a
# bare tcgen05.mma loop on all 148 SMs for 2.4 s (profiles/r2_umma_microbench_mxf4_peak.log): burst == sustained
MXF4_PEAK_TOPS = 8912.0
I8_PEAK_TOPS = 4590.0
# ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch of the 16384 x 100000 Gram (profiles/)
NCU_TRAFFIC = {"bxh": 6.710e9, "mxf4x1": 6.929e9, "bx": 6.607e9, "umma2": 8.174e9}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--block-rows", type=int, default=16384, help="candidate rows per GPU per step")
    p.add_argument("--train-rows", type=int, default=N_TRAIN)
    p.add_argument("--engine", default="bxh", choices=["bxh", "bx", "mxf4x1", "umma2", "umma2x1", "popc", "u8"],
                   help="Gram engine of the `value` leg (bxh = the library's default for bit fingerprints); the fused "
                        "screening legs use the library's default fused kernel unless --engine bx")
    p.add_argument("--cpu-sample-rows", type=int, default=16384, help="candidate rows of the CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-strong", action="store_true", help="skip the 1M-candidate strong-scaling leg")
    p.add_argument("--no-configs", action="store_true", help="skip the cfg2 / cfg3 / cfg5 legs (run at N = 1 only)")
    p.add_argument("--no-sustained", action="store_true")
    p.add_argument("--candidates", type=int, default=N_CAND_TOTAL, help="library size of the strong-scaling leg")
    p.add_argument("--topk", type=int, default=1000)
    p.add_argument("--no-numa-bind", action="store_true", help="do not pin each rank to its GPU's NUMA node")
    return p.parse_args()


def workload_config(args, world):
    return {
        "workload": "cfg4 BO screening cross-kernel: %d train x (%d candidates/GPU/step) x %d-bit, row-sharded"
                    % (args.train_rows, args.block_rows, D_BITS),
        "train_rows": args.train_rows, "block_rows": args.block_rows, "bits": D_BITS,
        "out_dtype": "fp32", "engine": args.engine, "parallelism": f"row-shard x{world}",
        "l2_policy": "operands + output of one step (26 MB packed train + %.1f GB Gram) exceed the 126 MB L2; no flush needed"
                     % (args.block_rows * args.train_rows * 4 / 1e9),
        "data": "synthetic ECFP-like bits (gauche_b200/synthetic.py), seeds 4/5",
    }


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        # samples under load = power within 40 % of the maximum seen
        if sm:
            thr = 0.6 * max(pw) if pw else 0
            loaded = [s for s, p in zip(sm, pw) if p >= thr] or sm
            return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                    "samples": len(sm), "samples_under_load": len(loaded), "reasons": sorted(reasons)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_pass(cand_cpu, train_cpu, alpha_cpu, chunk=1024):
    """Reference algorithm on the host: batch_tanimoto_sim block by block + K·alpha."""
    import torch
    from oracle.torch_port import tanimoto_sim_torch

    scores = torch.empty(cand_cpu.shape[0], dtype=torch.float32)
    for r0 in range(0, cand_cpu.shape[0], chunk):
        K = tanimoto_sim_torch(cand_cpu[r0:r0 + chunk], train_cpu)
        scores[r0:r0 + chunk] = K @ alpha_cpu
    return scores


def time_cpu_baseline(args, rows, train, cand, alpha, gpu_gram_head=None, gpu_scores=None):
    """Times the reference algorithm on the host on the SAME inputs the GPU arm used (copied back), and —
    since both arms then hold results for identical inputs — reports their agreement."""
    import torch
    from oracle.torch_port import tanimoto_sim_torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_reference_pass(cand[:256], train, alpha)  # warm-up
    t0 = time.perf_counter()
    scores = cpu_reference_pass(cand[:rows], train, alpha)
    dt = time.perf_counter() - t0
    out = {"value": rows * args.train_rows / dt / 1e9, "unit": UNIT, "cores": torch.get_num_threads(),
           "kind": "port", "seconds": dt,
           "sample": f"{rows} candidates x {args.train_rows} train x {D_BITS} bits fp32 (the GPU arm's own inputs), "
                     f"oracle/torch_port.py (same ATen ops as the reference) in 1024-row blocks + K@alpha"}
    parity = {}
    if gpu_gram_head is not None:
        k_cpu = tanimoto_sim_torch(cand[:gpu_gram_head.shape[0]], train)
        parity["gram_rows_compared"] = int(gpu_gram_head.shape[0])
        parity["gram_bit_identical"] = bool(torch.equal(k_cpu, gpu_gram_head))
        parity["gram_max_abs_diff"] = float((k_cpu - gpu_gram_head).abs().max())
    if gpu_scores is not None:
        # fp64 evaluation of the same sum on a sample: which fp32 arm is closer to the truth?
        ns = min(256, rows)
        truth = (tanimoto_sim_torch(cand[:ns].double(), train.double()) @ alpha.double())
        denom = float(truth.abs().max()) or 1.0
        parity["scores_max_rel_diff"] = float((scores - gpu_scores[:rows]).abs().max()) / (float(scores.abs().max()) or 1.0)
        parity["fp64_truth_rows"] = ns
        parity["gpu_fused_err_vs_fp64"] = float((gpu_scores[:ns].double() - truth).abs().max()) / denom
        parity["cpu_fp32_err_vs_fp64"] = float((scores[:ns].double() - truth).abs().max()) / denom
    out["parity_vs_gpu_arm"] = parity
    return out


def run_reference_arm(args):
    import torch
    from gauche_b200.synthetic import ecfp_bits

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = 1024  # bounded sample of one step (the full 16384-row step would take minutes per step)
    train = ecfp_bits(args.train_rows, D_BITS, seed=4)
    cand = ecfp_bits(rows * (args.steps + args.warmup), D_BITS, seed=5)
    alpha = torch.randn(args.train_rows, generator=torch.Generator().manual_seed(6)) * 1e-3
    for w in range(args.warmup):
        cpu_reference_pass(cand[w * rows:(w + 1) * rows], train, alpha)
    t0 = time.perf_counter()
    for s in range(args.steps):
        i = args.warmup + s
        cpu_reference_pass(cand[i * rows:(i + 1) * rows], train, alpha)
    dt = time.perf_counter() - t0
    val = rows * args.train_rows * args.steps / dt / 1e9
    cfg = workload_config(args, int(os.environ.get("WORLD_SIZE", "1")))
    sample = f"each step = {rows} of the {args.block_rows} candidate rows x {args.train_rows} train (bounded sample)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from gauche_b200 import _lib, ops
    from gauche_b200.io import PackedLibrary
    from gauche_b200.ops import _launch_tanimoto
    from gauche_b200.packed import PackedFingerprints, pack, stream_ptr
    from gauche_b200.screening import TanimotoScreener, bind_host_to_gpu_numa, gather_topk, shard_bounds, topk
    from gauche_b200.synthetic import ecfp_bits, ecfp_counts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = None if args.no_numa_bind else bind_host_to_gpu_numa(local_rank)   # GPU-local pinned buffers
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None   # streams while operands are generated and packed
    lib = _lib.load()
    R, M = args.block_rows, args.train_rows
    # ---- resident operands: replicated train set, this rank's candidate block
    train = pack(ecfp_bits(M, D_BITS, seed=4, device=dev))
    cand_dense = ecfp_bits(R, D_BITS, seed=5 + 1000 * rank, device=dev)
    cand = pack(cand_dense)
    assert train.kind == "binary" and cand.kind == "binary"
    alpha = (torch.randn(M, generator=torch.Generator().manual_seed(6)) * 1e-3).to(dev)
    gram = [torch.empty((R, M), dtype=torch.float32, device=dev) for _ in range(2)]
    stream = stream_ptr(dev)
    pairs_per_step = R * M

    def gram_step(i):
        _launch_tanimoto(lib, args.engine, cand, train, gram[i & 1], _lib.GK_F32, 1e-6, stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed_ms(fn, steps):
        """CUDA-event time of `steps` back-to-back calls on the current stream, bracketed by barriers."""
        sync_all()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            fn(i)
            ev[i + 1].record()
        sync_all()
        return ev[0].elapsed_time(ev[-1]), [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]

    # ---- leg 1: kernel-resident throughput (the clock sampler runs from start-up to the end of the timed legs)
    for i in range(max(args.warmup, 3)):
        gram_step(i)
    total_ms, per_launch_ms = timed_ms(gram_step, args.steps)
    total_ms_max = max_over_ranks(total_ms)
    value = pairs_per_step * world * args.steps / (total_ms_max * 1e-3) / 1e9
    clocks = sampler.stop() if sampler is not None else None

    # ---- leg 1b: the same launch for >= 2 s (power-capped regime), own clock record
    sustained = None
    if not args.no_sustained:
        s2 = ClockSampler(local_rank) if rank == 0 else None
        n_sus = max(args.steps, int(2000.0 / max(total_ms / args.steps, 1e-3)) + 1)
        sus_ms, _ = timed_ms(gram_step, n_sus)
        sus_ms = max_over_ranks(sus_ms)
        sustained = {"value": pairs_per_step * world * n_sus / (sus_ms * 1e-3) / 1e9, "unit": UNIT, "steps": n_sus,
                     "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus,
                     "clocks": s2.stop() if s2 is not None else None}

    # pack cost (reported separately; operands are packed once per library, not per step)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pack(cand_dense).kind
    e0.record()
    pack(cand_dense)
    e1.record()
    torch.cuda.synchronize()
    pack_ms = e0.elapsed_time(e1)

    # ---- leg 2: the fused screening kernel on device-resident packed operands (what e2e launches per block)
    screener = TanimotoScreener(train, alpha, bias=0.0, block_rows=R, engine="bx" if args.engine == "bx" else "auto")
    fused_out = torch.empty((R,), dtype=torch.float32, device=dev)
    for _ in range(3):
        screener.scores(cand, out=fused_out)
    fused_ms, fused_launch_ms = timed_ms(lambda i: screener.scores(cand, out=fused_out), args.steps)
    fused_ms = max_over_ranks(fused_ms)
    fused_val = pairs_per_step * world * args.steps / (fused_ms * 1e-3) / 1e9
    fused_tops = 2.0 * D_BITS * pairs_per_step / (statistics.mean(fused_launch_ms) * 1e-3) / 1e12
    roofline_fused = {
        "bound": "tensor", "achieved": fused_tops, "peak": MXF4_PEAK_TOPS, "unit": "TOP/s (e2m1, kind::mxf4)",
        "frac": fused_tops / MXF4_PEAK_TOPS, "traffic": None, "value_gpairs_per_s": fused_val,
        "kernel": "gram_tc_kernel<1, KindMxf4, TanimotoRowdotF32x2, %s> + reduce_partials_kernel (both inside the timed launch)"
                  % ("true" if screener.last_operand_kind == 2 else "false"),
        "algorithmic_per_unit": "4096 MAC-ops per pair (2 * 2048 bits); 0 B/pair written (K·alpha leaves the SM as one float per row and column-tile role)",
        "peak_source": "kind::mxf4.block_scale issue rate measured with a bare tcgen05.mma loop on 148 SMs for 2.4 s, "
                       "tools/umma_microbench.cu -> profiles/r2_umma_microbench_mxf4_peak.log (8912 TOP/s burst and sustained), of measured",
    }

    # ---- leg 3: end to end through the public screening API, host candidate buffers
    e2e = None
    if not args.no_e2e:
        NBLK = 8                                           # distinct pinned host blocks in rotation (1 GB of fp32)
        host_blocks = []
        for b in range(NBLK):
            blk = ecfp_bits(R, D_BITS, seed=5 + 1000 * rank + 17 * b, device=dev) if b else cand_dense
            host_blocks.append(blk.cpu().pin_memory())
            del blk
        host_scores = torch.empty((R,), dtype=torch.float32).pin_memory()
        scr = TanimotoScreener(train, alpha, bias=0.0, block_rows=R, engine=screener.engine)

        # headline e2e: the call a user makes — scores(host fp32 block) -> scores on the host; the library decides how the
        # block reaches the GPU (host-side bit packing + 4 MB copy, or the dense 134 MB copy: both are timed on live blocks)
        def e2e_steps(n_steps):
            last = None
            for i in range(n_steps):
                s = scr.scores(host_blocks[i % NBLK])    # host packing of block i+1 overlaps the GPU work of block i
                host_scores.copy_(s, non_blocking=True)  # D2H of the step's result
                last = s
            if last is not None:
                vals, idx = topk(last, min(args.topk, R), index_base=rank * R)
                gather_topk(vals, idx, min(args.topk, R))    # the path's only collective
            torch.cuda.synchronize()

        e2e_steps(max(6, args.warmup))                       # also lets the library measure and pick its host route
        sync_all()
        bytes0 = scr.host_bytes_copied
        t0 = time.perf_counter()
        e2e_steps(args.steps)
        sync_all()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        h2d_per_step = (scr.host_bytes_copied - bytes0) // max(args.steps, 1)
        e2e_val = pairs_per_step * world * args.steps / e2e_s / 1e9
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d_per_step) * world,
               "d2h_bytes_per_step": R * 4 * world,
               "api": "TanimotoScreener.scores(pinned host fp32 block [16384 x 2048]) -> scores on host, 8 distinct blocks in "
                      "rotation; the 134 MB block is bit-packed on the host cores (gk_host_pack_bits, thread pool + AVX-512/AVX2) "
                      "or copied dense, whichever the library MEASURED faster on live blocks; K·alpha fused into the Gram "
                      "epilogue (every pair still computed)",
               "host_pack": scr.host_pack_stats, "ms_per_step": e2e_s / args.steps * 1e3, "host_numa_node_rank0": numa_node,
               "host_threads_per_rank": scr._host_threads()}

        # what the dense-fp32 host format permits: all ranks stream their blocks through the host packer at once
        # (reads 8 KB per candidate from host DRAM) -> aggregate host read bandwidth -> Gpairs/s bound
        forced = TanimotoScreener(train, alpha, bias=0.0, block_rows=R, host_pack=True)
        forced.scores(host_blocks[0]); forced.scores(host_blocks[1])
        sync_all()
        import ctypes
        st = forced._host_stage
        flag = ctypes.c_int32(0)
        t0 = time.perf_counter()
        for b in range(NBLK):
            blk = host_blocks[b]
            lib.gk_host_pack_bits(blk.data_ptr(), _lib.GK_F32, R, D_BITS, blk.stride(0), st["pin"][0].data_ptr(), st["words"],
                                  st["pin"][0].data_ptr() + R * st["words"] * 4, ctypes.byref(flag), forced._host_threads())
        t_pack_all = time.perf_counter() - t0
        host_gbs = sum_over_ranks(NBLK * R * D_BITS * 4 / t_pack_all / 1e9)
        e2e["host_bound"] = {"host_read_gbs_all_ranks": host_gbs, "value": host_gbs * 1e9 / (D_BITS * 4) * M / 1e9, "unit": UNIT,
                             "note": "aggregate rate at which the box's host cores stream dense fp32 candidates (8 KB each) "
                                     "through gk_host_pack_bits, all ranks concurrently; e2e from this host format cannot exceed it"}

        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def timed_variant(host_bufs, dev_bufs, make_arg):
            """same double-buffered H2D / scores / D2H loop for another host format (explicit copies on a copy stream)"""
            nb = len(host_bufs)

            def h2d_v(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i & 1])
                    dev_bufs[i & 1].copy_(host_bufs[i % nb], non_blocking=True)
                    copied[i & 1].record(copy_stream)

            def steps_v(n_steps):
                cur = torch.cuda.current_stream(dev)
                for b in range(2):
                    consumed[b].record(cur)
                h2d_v(0)
                for i in range(n_steps):
                    if i + 1 < n_steps:
                        h2d_v(i + 1)
                    cur.wait_event(copied[i & 1])
                    sc = scr.scores(make_arg(dev_bufs[i & 1]))
                    consumed[i & 1].record(cur)
                    host_scores.copy_(sc, non_blocking=True)
                torch.cuda.synchronize()

            steps_v(3)
            sync_all()
            t0v = time.perf_counter()
            steps_v(args.steps)
            sync_all()
            return max_over_ranks(time.perf_counter() - t0v)

        def variant(seconds, h2d, note):
            return {"value": pairs_per_step * world * args.steps / seconds / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": R * 4 * world, "note": note, "ms_per_step": seconds / args.steps * 1e3}

        dev_blocks = [torch.empty_like(cand_dense) for _ in range(2)]
        td = timed_variant(host_blocks, dev_blocks, lambda t_: t_)
        e2e["dense_h2d_variant"] = variant(td, R * D_BITS * 4, "explicit double-buffered H2D of the dense fp32 block on a copy stream, then scores(device block)")
        del dev_blocks

        # candidates delivered in the packed wire format (gauche_b200.io: 256 B per molecule instead of 8 KB of fp32)
        wire = [torch.from_numpy(np.packbits(hb.numpy().astype(np.uint8), axis=1, bitorder="little")).pin_memory()
                for hb in host_blocks]
        dev_wire = [torch.empty((R, D_BITS // 8), dtype=torch.uint8, device=dev) for _ in range(2)]
        tw = timed_variant(wire, dev_wire, lambda t: PackedFingerprints.from_bits(t, D_BITS))
        e2e["packed_wire_variant"] = variant(tw, R * (D_BITS // 8), "host block in the packed wire format (np.packbits little-endian rows) -> PackedFingerprints.from_bits -> same fused scores; no expansion pass: the kernel reads the bits")
        libs = [PackedLibrary(w.numpy(), D_BITS) for w in wire]
        for i in range(3):
            scr.scores(libs[i % NBLK])
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.steps):
            host_scores.copy_(scr.scores(libs[i % NBLK]), non_blocking=True)
        sync_all()
        tl = max_over_ranks(time.perf_counter() - t0)
        e2e["packed_library_variant"] = variant(tl, R * (D_BITS // 8), "TanimotoScreener.scores(io.PackedLibrary block): the library stages the packed rows through its own pinned double buffer")

        host_u8 = [hb.to(torch.uint8).pin_memory() for hb in host_blocks[:4]]
        dev_u8 = [torch.empty((R, D_BITS), dtype=torch.uint8, device=dev) for _ in range(2)]
        tu = timed_variant(host_u8, dev_u8, lambda t: t)
        e2e["u8_host_variant"] = variant(tu, R * D_BITS, "host block as dense uint8 0/1 rows -> same TanimotoScreener.scores call")
        del host_u8, dev_u8, wire, dev_wire, libs, host_blocks

    # ---- leg 4: the real config as strong scaling — the whole candidate library sharded over the ranks
    strong = None
    if not args.no_strong:
        n_total = args.candidates
        start, stop = shard_bounds(n_total, world, rank)
        CH = 62_500                                           # chunk g is always built from seed 1000 + g: same library for every N
        chunks = []
        for g in range(start // CH, (stop + CH - 1) // CH):
            g0, g1 = g * CH, min(n_total, (g + 1) * CH)
            dense = ecfp_bits(g1 - g0, D_BITS, seed=1000 + g, device=dev)
            lo, hi = max(start, g0) - g0, min(stop, g1) - g0
            wire = pack(dense[lo:hi]).bits.view(torch.uint8)                 # the packed wire format, device resident
            chunks.append(PackedFingerprints.from_bits(wire.view(hi - lo, D_BITS // 8), D_BITS))
            del dense
        scr4 = TanimotoScreener(train, alpha, bias=0.0, block_rows=R, engine=screener.engine)
        scores4 = torch.empty((stop - start,), dtype=torch.float32, device=dev)
        k4 = min(args.topk, max(stop - start, 1))
        g_ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
        result = {}

        def screen_once(_i=0):
            off = 0
            for p in chunks:
                scr4.scores(p, out=scores4[off:off + p.rows])
                off += p.rows
            v, i = topk(scores4, k4, index_base=start)
            g_ev[0].record()
            result["top"] = gather_topk(v, i, k4)
            g_ev[1].record()

        for _ in range(2):
            screen_once()
        n4 = max(3, min(args.steps, 20))
        ms4, per4 = timed_ms(screen_once, n4)
        ms4 = max_over_ranks(ms4) / n4
        gather_ms = max_over_ranks(g_ev[0].elapsed_time(g_ev[1]))
        vals4, idx4 = result["top"]
        strong = {"candidates": n_total, "train_rows": M, "n_gpus": world, "time_to_topk_ms": ms4,
                  "value": n_total * M / (ms4 * 1e-3) / 1e9, "unit": UNIT, "scaling": "strong", "steps": n4, "topk": k4,
                  "allgather_ms": gather_ms, "allgather_share": gather_ms / ms4 if ms4 > 0 else None,
                  "top1": {"score": float(vals4[0]), "index": int(idx4[0])},
                  "topk_index_checksum": int((idx4 * torch.arange(1, k4 + 1, device=dev)).sum().item()),
                  "note": "whole job: every rank scores its contiguous shard of the packed-bit library (device resident, 256 B per "
                          "molecule) with the fused kernel, selects its top-k, ONE all-gather of 12-byte records, identical merge on "
                          "every rank; barrier-to-barrier CUDA-event time, max over ranks; the checksum is the same for every N"}
        del chunks, scores4

    # ---- leg 4b: cfg5 (sparse GP, 1 M fingerprints x 4096 inducing points) sharded the same way: K_xz rows of this rank
    #      + the replicated K_zz, Grams written to HBM through the public kernel class
    cfg5 = None
    if not args.no_strong:
        import gauche_b200 as gb
        from gauche_b200.packed import pack_cache as _pc
        _pc.max_entries = max(_pc.max_entries, 16)                            # 8 row blocks + Z stay cached
        n5 = args.candidates
        s5, e5 = shard_bounds(n5, world, rank)
        z5 = ecfp_bits(4096, D_BITS, seed=6, device=dev)                      # frozen inducing rows, replicated
        kern5 = gb.TanimotoKernel()
        CH5 = 125_000
        x5_chunks = [ecfp_bits(min(CH5, e5 - c0), D_BITS, seed=2000 + c0 // CH5, device=dev) for c0 in range(s5, e5, CH5)]

        def cfg5_once(_i=0):
            for xc in x5_chunks:                                               # K_xz shard, 125 000-row blocks (2 GB each)
                kern5.forward(xc, z5)
            kern5.forward(z5, z5)                                              # K_zz, computed redundantly per rank

        for _ in range(2):
            cfg5_once()
        n5r = max(3, min(args.steps, 10))
        ms5, _ = timed_ms(cfg5_once, n5r)
        ms5 = max_over_ranks(ms5) / n5r
        cfg5 = {"rows": n5, "inducing": 4096, "n_gpus": world, "time_ms": ms5, "value": (n5 * 4096 + world * 4096 * 4096) / (ms5 * 1e-3) / 1e9,
                "unit": UNIT, "scaling": "strong",
                "note": "TanimotoKernel.forward(X_shard, Z) in 125 000-row blocks (operand cache hits after the first pass) + "
                        "TanimotoKernel.forward(Z, Z) on every rank; fp32 Grams written to HBM; no collective (the SGPR "
                        "reductions belong to the untouched GPyTorch solve)"}
        del x5_chunks, z5

    # ---- leg 5: the other BASELINE.json configs (single GPU): CUDA-event times through the public API
    configs = None
    if world == 1 and not args.no_configs:
        configs = {}
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

        def event_ms(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        import gauche_b200 as gb
        # cfg2: Lipophilicity-sized exact-GP Gram, 4200 x 4200 x 2048 bits, fp32 (operand cache hit: the parameter-free
        # kernel is re-evaluated on the same train tensor every L-BFGS closure)
        x2 = ecfp_bits(4200, D_BITS, seed=2, device=dev)
        gb.batch_tanimoto_sim(x2, x2)
        ms = event_ms(lambda: gb.batch_tanimoto_sim(x2, x2), reps=50)
        bytes2 = 4200 * 4200 * 4 + 2 * 4200 * 256
        configs["cfg2_gram_4200_fp32"] = {"ms": ms, "gpairs_per_s": 4200 * 4200 / ms / 1e6, "hbm_gbs": bytes2 / ms / 1e6,
                                          "frac_of_hbm": bytes2 / ms / 1e6 / hbm_peak,
                                          "note": "public call batch_tanimoto_sim(x, x), cached operands; 128 x 192 tiles: 33 x 22 = 726 tiles on 148 SMs (4.9 waves) — launch / wave-quantisation bound"}
        del x2
        # cfg5: sparse-GP shard, K_xz 125 000 x 4096 and K_zz 4096^2
        lib5 = ecfp_bits(4096 + 125_000, D_BITS, seed=6, device=dev)
        z5, x5 = lib5[:4096], lib5[4096:]
        gb.batch_tanimoto_sim(x5, z5); gb.batch_tanimoto_sim(z5, z5)
        ms_xz = event_ms(lambda: gb.batch_tanimoto_sim(x5, z5), reps=20)
        ms_zz = event_ms(lambda: gb.batch_tanimoto_sim(z5, z5), reps=50)
        bytes5 = 125_000 * 4096 * 4 + (125_000 + 4096) * 256
        configs["cfg5_kxz_125000x4096_fp32"] = {"ms": ms_xz, "gpairs_per_s": 125_000 * 4096 / ms_xz / 1e6,
                                                "hbm_gbs": bytes5 / ms_xz / 1e6, "frac_of_hbm": bytes5 / ms_xz / 1e6 / hbm_peak,
                                                "note": "one rank's share of K(X, Z) at 8 GPUs (1M x 4096 inducing points), Gram written to HBM"}
        configs["cfg5_kzz_4096_fp32"] = {"ms": ms_zz, "gpairs_per_s": 4096 * 4096 / ms_zz / 1e6}
        del lib5, z5, x5
        # cfg3: MinMax on 50 000 count fingerprints (block-sparse thermometer planes on kind::mxf4)
        x3 = ecfp_counts(50_000, D_BITS, seed=3, device=dev)
        k3 = ops.minmax_similarity(x3, x3)
        del k3
        ms3 = event_ms(lambda: ops.minmax_similarity(x3, x3), reps=5)
        from gauche_b200.packed import pack_cache
        p3 = pack_cache.get(x3)
        planes = p3.maxval
        kb = planes * p3.k4_bytes // 128
        sym3 = ops.last_launch["minmax"] == "gk_minmax_gram_mxf4_sym"     # upper-triangle tiles + mirrored stores
        pair_rows3 = ops.MINMAX_SYM_PAIR_MIN_ROWS if sym3 else ops.MINMAX_PAIR_MIN_ROWS
        cg3 = 2 if (ops.MINMAX_PAIRS and 50_000 >= pair_rows3) else 1      # CTA pairs: 256-row tiles
        occ_a = p3.occupancy(planes, 128 * cg3).view(-1, (kb + 63) // 64)
        occ_b = p3.occupancy(planes, 224).view(-1, (kb + 63) // 64)
        ks = torch.arange(kb, device=dev)
        bits_a = ((occ_a[:, ks // 64] >> (ks % 64)) & 1).double()           # [row tiles, k-blocks]
        bits_b = ((occ_b[:, ks // 64] >> (ks % 64)) & 1).double()
        bits_a[:, 0] = 1.0                                                  # k-block 0 is always issued
        bits_b[:, 0] = 1.0
        per_tile = bits_a @ bits_b.t()                                      # k-blocks issued per (row tile, column tile)
        ti = torch.arange(occ_a.shape[0], device=dev)[:, None] * (128 * cg3)
        tj = (torch.arange(occ_b.shape[0], device=dev)[None, :] + 1) * 224
        live = (tj > ti) if sym3 else torch.ones_like(per_tile, dtype=torch.bool)   # tiles not strictly below the diagonal
        issued = float((per_tile * live).sum().item())
        dense_blocks = occ_a.shape[0] * occ_b.shape[0] * kb
        tops3 = issued * 128.0 * cg3 * 224.0 * 256.0 * 2.0 / (ms3 * 1e-3) / 1e12
        configs["cfg3_minmax_50000_fp32"] = {
            "ms": ms3, "gpairs_per_s": 50_000 * 50_000 / ms3 / 1e6, "thermometer_planes": planes, "cta_group": cg3,
            "kernel": ops.last_launch["minmax"], "symmetric": bool(sym3), "live_tile_fraction": float(live.double().mean().item()),
            "issued_kblock_fraction": issued / dense_blocks, "issued_tops": tops3, "frac_of_mxf4_peak": tops3 / MXF4_PEAK_TOPS,
            "hbm_write_gbs": 50_000 * 50_000 * 4 / ms3 / 1e6, "frac_of_hbm": 50_000 * 50_000 * 4 / ms3 / 1e6 / hbm_peak,
            "note": "public call minmax_similarity(x, x) incl. the allocation of the 10 GB result; operands (thermometer planes, "
                    "occupancy maps) cached; x1 is x2 -> only tiles touching or above the diagonal are computed (the MinMax Gram is "
                    "bitwise symmetric), every chunk stored as computed and mirrored; issued_* count the MMAs actually issued"}
        del x3, p3, occ_a, occ_b
        pack_cache.clear()
        # the baselines a GPU user has without this library, on the bench's own cfg4 block (dense fp32 0/1 tensors resident in
        # HBM): the reference's algorithm as the same torch ops on the GPU (tanimoto_kernel.py:36-46: matmul + two norms + the
        # elementwise epilogue), and cuBLASLt int8 (torch._int_mm) + that epilogue.  Informational; not the reference arm.
        try:
            torch.cuda.empty_cache()
            t_dense = ecfp_bits(M, D_BITS, seed=4, device=dev)

            def torch_ops():
                dot = cand_dense @ t_dense.t()
                n1 = (cand_dense ** 2).sum(-1, keepdim=True)
                n2 = (t_dense ** 2).sum(-1, keepdim=True)
                return ((dot + 1e-6) / (1e-6 + n1 + n2.t() - dot)).clamp_min_(0)

            torch_ops()
            ms_t = event_ms(torch_ops, reps=3)
            c8, t8 = cand_dense.to(torch.int8), t_dense.to(torch.int8).t().contiguous()
            n1_, n2_ = cand_dense.sum(1, keepdim=True), t_dense.sum(1, keepdim=True).t()

            def int_mm():
                dot = torch._int_mm(c8, t8).float()
                return ((dot + 1e-6) / (1e-6 + n1_ + n2_ - dot)).clamp_min_(0)

            int_mm()
            ms_i = event_ms(int_mm, reps=3)
            configs["gpu_baselines_cfg4_block"] = {
                "torch_ops_on_gpu": {"ms": ms_t, "gpairs_per_s": pairs_per_step / ms_t / 1e6},
                "cublaslt_int8_plus_torch_epilogue": {"ms": ms_i, "gpairs_per_s": pairs_per_step / ms_i / 1e6},
                "note": "same 16384 x 100000 x 2048 block as `value`, dense fp32 0/1 inputs already in HBM; torch %s" % torch.__version__}
            del t_dense, c8, t8
            torch.cuda.empty_cache()
        except Exception as exc:      # informational leg: never fails the bench
            configs["gpu_baselines_cfg4_block"] = {"unavailable": repr(exc)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from live CUDA-event timings
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak = json.load(open(peaks_path))["hbm_gbs"]
        hbm_src = "hbm_gbs of MEASURED_PEAKS.json (copy bandwidth), of measured"
    else:
        hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md, of fallback"
    avg_launch_ms = statistics.mean(per_launch_ms)
    sec = avg_launch_ms * 1e-3
    tensor_tops = 2.0 * D_BITS * pairs_per_step / sec / 1e12                # algorithmic MAC ops (2 per bit pair)
    measured_traffic = NCU_TRAFFIC.get(args.engine) if (R, M) == (16384, 100000) else None
    if args.engine in ("bxh", "bx", "mxf4x1"):
        # operand rows in HBM: packed bits (256 B) | e2m1 nibbles (1 KB); the hybrid kernel reads candidates as bits, train as e2m1
        a_row = D_BITS // 2 if args.engine == "mxf4x1" else D_BITS // 8
        b_row = D_BITS // 8 if args.engine == "bx" else D_BITS // 2
        bytes_per_launch = 4.0 * pairs_per_step + R * a_row + M * b_row
        achieved = bytes_per_launch / sec / 1e9
        kname = {"bxh": "gram_tc_kernel<1, KindMxf4T<192>, TanimotoF32x2, 3> (hybrid: A bits -> TMEM, B e2m1 by TMA)",
                 "bx": "gram_tc_kernel<1, KindMxf4, TanimotoF32x2, 1> (bit-expanding)",
                 "mxf4x1": "gram_tc_kernel<1, KindMxf4, TanimotoF32x2, 0>"}[args.engine]
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": measured_traffic,
                "kernel": kname,
                "algorithmic_per_unit": "4 B/pair written (fp32 Gram) + each operand row read once (%d B per candidate row, %d B per train row)" % (a_row, b_row),
                "peak_source": hbm_src,
                "tensor": {"achieved_tops": tensor_tops, "peak_mxf4_tops": MXF4_PEAK_TOPS, "frac_of_mxf4_peak": tensor_tops / MXF4_PEAK_TOPS,
                           "note": "4096 MAC-ops/pair against the measured kind::mxf4 rate (profiles/r2_umma_microbench_mxf4_peak.log)"}}
    elif args.engine.startswith("umma"):
        roof = {"bound": "tensor", "achieved": tensor_tops, "peak": I8_PEAK_TOPS, "unit": "TOP/s (int8)",
                "frac": tensor_tops / I8_PEAK_TOPS, "traffic": measured_traffic,
                "kernel": "gram_tc_kernel<%d, KindI8, TanimotoF32x2, false>" % (2 if args.engine == "umma2" else 1),
                "algorithmic_per_unit": "4096 int8-ops/pair (2*D) ; 4 B/pair written",
                "peak_source": "kind::i8 issue rate measured by tools/umma_microbench.cu (profiles/r2_umma_microbench_mxf4_peak.log), of measured",
                "hbm_write_gbs": 4.0 * pairs_per_step / sec / 1e9,
                "hbm_write_frac_of_measured": 4.0 * pairs_per_step / sec / 1e9 / hbm_peak}
    else:
        achieved = 4.0 * pairs_per_step / sec / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "kernel": "gram_words_kernel", "algorithmic_per_unit": "4 B/pair written", "peak_source": hbm_src}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rows = min(args.cpu_sample_rows, R)
        gram_step(0)
        torch.cuda.synchronize()
        gram_head = gram[0][:512].cpu()                      # 512 x M Gram rows of the GPU arm
        gpu_scores = screener.scores(cand).cpu()
        train_np = np.unpackbits(train.bits.cpu().numpy().view("uint8"), axis=1, bitorder="little")[:, :D_BITS]
        train_cpu = torch.from_numpy(train_np.astype("float32"))   # the dense train set rebuilt from the packed bits
        cpu = time_cpu_baseline(args, rows, train_cpu, cand_dense.cpu(), alpha.cpu(), gram_head, gpu_scores)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("e2m1 bits (exact f32 accumulation)" if args.engine in ("bxh", "bx", "mxf4x1") else "u8 (exact s32 accumulation)") + ", fp32 epilogue",
        "data": "synthetic", "config": workload_config(args, world),
        "roofline": roof, "roofline_fused": roofline_fused, "sustained": sustained, "cpu_baseline": cpu, "e2e": e2e,
        "strong_scaling": strong, "strong_scaling_cfg5": cfg5, "configs": configs, "clocks": clocks,
        "gpu_launches": args.steps, "pack_ms_per_block": pack_ms,
        "per_launch_ms": {"mean": avg_launch_ms, "min": min(per_launch_ms), "max": max(per_launch_ms)},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
