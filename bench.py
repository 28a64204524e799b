#!/usr/bin/env python
"""bench.py — Tanimoto cross-kernel throughput (Gpairs/s @2048-bit) on 1..8 B200.

Workload (BASELINE.json configs[3], the configuration the metric's "1/2/4/8 B200" refers to):
BO candidate screening, 100 000 train x 1 000 000 candidate 2048-bit fingerprints, candidate
rows sharded over the GPUs.  One STEP = one pass of the hot path over one candidate row block
per GPU: K(X*_block [R x 2048], X_train [100k x 2048]) -> fp32 Gram block written to HBM once.
Weak scaling: every rank processes its own R-row block per step; no data-path collective.

  value    Gram-kernel throughput, packed operands resident in HBM (CUDA events, max over ranks)
  e2e      same metric through the public screening API with HOST candidate buffers: per step
           H2D of the dense fp32 block from pinned memory, pack, Gram, K·alpha, D2H of the scores
  roofline tcgen05 int8 pipe: 2*D int8-ops per pair vs 2 x measured bf16 dense peak
  cpu_baseline  the reference's torch-CPU algorithm (oracle/torch_port.py) on the host cores

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
METRIC = "tanimoto_gram_gpairs_per_s_2048bit"
UNIT = "Gpairs/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--block-rows", type=int, default=16384, help="candidate rows per GPU per step")
    p.add_argument("--train-rows", type=int, default=N_TRAIN)
    p.add_argument("--engine", default="bx", choices=["bx", "mxf4x1", "umma2", "umma2x1", "popc", "u8"])
    p.add_argument("--cpu-sample-rows", type=int, default=16384, help="candidate rows of the CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--unfused-e2e", action="store_true", help="e2e writes the Gram block and re-reads it for K·alpha")
    p.add_argument("--topk", type=int, default=1000)
    p.add_argument("--no-numa-bind", action="store_true", help="do not pin each rank to its GPU's NUMA node")
    return p.parse_args()


def workload_config(args, world):
    return {
        "workload": "cfg4 BO screening cross-kernel: %d train x (%d candidates/GPU/step) x %d-bit, row-sharded"
                    % (args.train_rows, args.block_rows, D_BITS),
        "train_rows": args.train_rows, "block_rows": args.block_rows, "bits": D_BITS,
        "out_dtype": "fp32", "engine": args.engine, "parallelism": f"row-shard x{world}",
        "l2_policy": "operands+output per step (205 MB u8 train + %.1f GB Gram) exceed the 126 MB L2; no flush needed"
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
        denom = float(scores.abs().max()) or 1.0
        parity["scores_max_rel_diff"] = float((scores - gpu_scores[:rows]).abs().max()) / denom
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

    import torch
    import torch.distributed as dist

    from gauche_b200 import _lib
    from gauche_b200.ops import _launch_tanimoto
    from gauche_b200.packed import pack, stream_ptr
    from gauche_b200.screening import TanimotoScreener, gather_topk, topk
    from gauche_b200.synthetic import ecfp_bits

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from gauche_b200.screening import bind_host_to_gpu_numa
    numa_node = None if args.no_numa_bind else bind_host_to_gpu_numa(local_rank)   # GPU-local pinned buffers
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None   # streams while operands are generated and packed
    lib = _lib.load()
    R, M = args.block_rows, args.train_rows
    # ---- resident operands: replicated train set, this rank's candidate block(s)
    train_dense = ecfp_bits(M, D_BITS, seed=4, device=dev)
    train = pack(train_dense)
    del train_dense
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

    # ---- kernel-resident throughput (the clock sampler runs from start-up to the end of the e2e region)
    for i in range(args.warmup):
        gram_step(i)
    sync_all()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        gram_step(i)
        ev[i + 1].record()
    sync_all()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = pairs_per_step * world * args.steps / (total_ms_max * 1e-3) / 1e9

    # pack cost (reported separately; operands are packed once per library, not per step)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pack(cand_dense)
    e0.record()
    pack(cand_dense)
    e1.record()
    torch.cuda.synchronize()
    pack_ms = e0.elapsed_time(e1)

    # ---- end to end through the public screening API, host candidate buffers
    e2e = None
    if not args.no_e2e:
        screener = TanimotoScreener(train, alpha, bias=0.0, block_rows=R, engine=args.engine, fused=not args.unfused_e2e)
        host_blocks = [cand_dense.cpu().pin_memory() for _ in range(2)]
        host_scores = torch.empty((R,), dtype=torch.float32).pin_memory()
        dev_blocks = [torch.empty_like(cand_dense) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        # headline e2e: the call a user makes — scores(host fp32 block) -> scores; the library decides how the block
        # reaches the GPU (host-side bit packing on the box's cores + 4 MB copy, or the dense 134 MB copy: it times
        # the first block and keeps the faster)
        def e2e_steps(n_steps):
            last = None
            for i in range(n_steps):
                s = screener.scores(host_blocks[i & 1])      # host packing of block i+1 overlaps the GPU work of block i
                host_scores.copy_(s, non_blocking=True)  # D2H of the step's result
                last = s
            if last is not None:
                vals, idx = topk(last, min(args.topk, R), index_base=rank * R)
                gather_topk(vals, idx, min(args.topk, R))    # the path's only collective
            torch.cuda.synchronize()

        e2e_steps(max(2, args.warmup))
        sync_all()
        bytes0 = screener.host_bytes_copied
        t0 = time.perf_counter()
        e2e_steps(args.steps)
        sync_all()
        e2e_s = time.perf_counter() - t0
        h2d_per_step = (screener.host_bytes_copied - bytes0) // max(args.steps, 1)
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_val = pairs_per_step * world * args.steps / float(t.item()) / 1e9
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d_per_step) * world,
               "d2h_bytes_per_step": R * 4 * world,
               "api": "TanimotoScreener.scores(pinned host fp32 block [16384 x 2048]) -> scores on host; the 134 MB block is "
                      "bit-packed on the host cores (gk_host_pack_bits, thread pool + AVX2) when that beats the dense PCIe copy "
                      "(decided by timing the first block), then 4.3 MB (bits + row popcounts) are copied; "
                      + ("Gram block written then K·alpha" if args.unfused_e2e else "K·alpha fused into the Gram epilogue (every pair still computed)"),
               "host_pack": screener.host_pack_stats,
               "ms_per_step": float(t.item()) / args.steps * 1e3, "host_numa_node_rank0": numa_node}

        def timed_variant(host_bufs, dev_bufs, make_arg):
            """same double-buffered H2D / scores / D2H loop as the fp32 leg for another host format"""
            def h2d_v(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i & 1])
                    dev_bufs[i & 1].copy_(host_bufs[i & 1], non_blocking=True)
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
                    sc = screener.scores(make_arg(dev_bufs[i & 1]))
                    consumed[i & 1].record(cur)
                    host_scores.copy_(sc, non_blocking=True)
                torch.cuda.synchronize()

            steps_v(3)
            sync_all()
            t0v = time.perf_counter()
            steps_v(args.steps)
            sync_all()
            tv = torch.tensor([time.perf_counter() - t0v], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            return float(tv.item())

        # the dense fp32 block copied as is (what round-1 measured as e2e before host packing): PCIe bound
        td = timed_variant(host_blocks, dev_blocks, lambda t_: t_)
        e2e["dense_h2d_variant"] = {
            "value": pairs_per_step * world * args.steps / td / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": R * D_BITS * 4 * world, "d2h_bytes_per_step": R * 4 * world,
            "note": "explicit double-buffered H2D of the dense fp32 block on a copy stream, then scores(device block)",
            "ms_per_step": td / args.steps * 1e3}

        # same API, candidates delivered in the packed-bit wire format (256 B per molecule instead of
        # 8 KB of fp32): H2D of the packed block + device-side expansion inside the timed region
        from gauche_b200.packed import PackedFingerprints
        import numpy as np
        wire_np = np.packbits(cand_dense.cpu().numpy().astype(np.uint8), axis=1, bitorder="little")
        host_wire = [torch.from_numpy(wire_np).pin_memory() for _ in range(2)]
        dev_wire = [torch.empty((R, D_BITS // 8), dtype=torch.uint8, device=dev) for _ in range(2)]
        tw = timed_variant(host_wire, dev_wire, lambda t: PackedFingerprints.from_bits(t, D_BITS))
        e2e["packed_wire_variant"] = {
            "value": pairs_per_step * world * args.steps / tw / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": R * (D_BITS // 8) * world, "d2h_bytes_per_step": R * 4 * world,
            "note": "host block as np.packbits(bitorder=little) uint8 rows -> PackedFingerprints.from_bits -> same fused scores",
            "ms_per_step": tw / args.steps * 1e3}

        # same API, the dense block held as uint8 0/1 on the host (what RDKit -> numpy gives before the
        # reference's cast to float): 2 KB per molecule
        host_u8 = [cand_dense.to(torch.uint8).cpu().pin_memory() for _ in range(2)]
        dev_u8 = [torch.empty((R, D_BITS), dtype=torch.uint8, device=dev) for _ in range(2)]
        tu = timed_variant(host_u8, dev_u8, lambda t: t)
        e2e["u8_host_variant"] = {
            "value": pairs_per_step * world * args.steps / tu / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": R * D_BITS * world, "d2h_bytes_per_step": R * 4 * world,
            "note": "host block as dense uint8 0/1 rows -> same TanimotoScreener.scores call",
            "ms_per_step": tu / args.steps * 1e3}

    clocks = sampler.stop() if sampler is not None else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tcgen05 int8 Gram), from live CUDA-event timings
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    long_region = total_ms_max >= 1000.0   # the 1 kW power cap bites after ~0.3-1 s of dense tensor work
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
        if long_region:
            bf16_sustained = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
            peak_src = "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 dense = 2x bf16 rate; timed region >= 1 s, power-capped), of measured"
        else:
            bf16_sustained = peaks["bf16_tflops"]
            peak_src = "2 x bf16_tflops (burst) of MEASURED_PEAKS.json (int8 dense = 2x bf16 rate; timed region < 1 s), of measured"
        hbm_peak = peaks["hbm_gbs"]
    else:
        bf16_sustained, hbm_peak = (1400.0 if long_region else 1590.0), 6650.0
        peak_src = "2 x bf16 fallback of B200_PROFILING.md (%s), of fallback" % ("sustained" if long_region else "burst")
    avg_launch_ms = statistics.mean(per_launch_ms)
    sec = avg_launch_ms * 1e-3
    ops_per_launch = 2.0 * D_BITS * pairs_per_step                # algorithmic MAC ops (2 per bit pair)
    tensor_tops = ops_per_launch / sec / 1e12
    # ncu --set full capture of this exact launch (profiles/r1_ncu_full_*_fullsize_16384x100000.txt):
    # dram__bytes_read.sum + dram__bytes_write.sum per launch
    measured_traffic = {"umma2": 8.174e9, "mxf4x1": 6.928e9}.get(args.engine) if (R, M) == (16384, 100000) else None
    if args.engine in ("bx", "mxf4x1"):
        # packed-e2m1 operands on kind::mxf4: tensor time (4096 ops / 4x bf16 peak) and HBM write time
        # (4 B / pair) are within 3 % of each other; HBM is the tighter one and the one reported
        kbytes = D_BITS // 2 if args.engine == "mxf4x1" else D_BITS // 8     # operand row in HBM: e2m1 nibbles | packed bits
        bytes_per_launch = 4.0 * pairs_per_step + (R + M) * kbytes
        achieved = bytes_per_launch / sec / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": measured_traffic, "kernel": "gram_tc_kernel<1, KindMxf4, TanimotoF32x2, %s>" % ("true (bit-expanding)" if args.engine == "bx" else "false"),
                "algorithmic_per_unit": "4 B/pair written (fp32 Gram) + each 1 KB e2m1 operand row read once",
                "peak_source": "hbm_gbs of MEASURED_PEAKS.json (copy bandwidth), of measured",
                "tensor": {"achieved_tops": tensor_tops, "peak_fp4_tops": 4.0 * bf16_sustained,
                           "frac_of_fp4_peak": tensor_tops / (4.0 * bf16_sustained),
                           "frac_of_int8_peak": tensor_tops / (2.0 * bf16_sustained),
                           "note": "4096 MAC-ops/pair; fp4 dense = 4x, int8 dense = 2x the measured bf16 rate; " + peak_src},
                "limiter": "operand reads and Gram writes serialise chip-wide: time/pair fits operand_bytes/21.3 TB/s + 4 B/7.2 TB/s = 868 Gpairs/s at 2048 bit (fingerprint-length sweep, profiles/r1_store_stream_ceiling.log); same rate with all epilogue math removed"}
        # empirical bound of this kernel's traffic mix (128 x 224 tiles: (128 + 224) operand rows of kbytes per tile)
        sm_in_bytes_per_pair = (128 + 224) * kbytes / (128.0 * 224.0)
        mix_bound = 1.0 / (sm_in_bytes_per_pair / 21.3e12 + 4.0 / 7.2e12) / 1e9
        roof["empirical_mix_bound"] = {"gpairs_per_s_per_gpu": mix_bound, "frac": value / world / mix_bound,
                                       "model": "per pair: bytes into the SMs / 21.3 TB/s + 4 B written / 7.2 TB/s (no overlap chip-wide); the two constants are FITTED to this kernel at 2048 and 4096 bits (profiles/r1_store_stream_ceiling.log), so ~1.0 here says the model holds, not that a hardware peak is reached"}
    elif args.engine.startswith("umma"):
        peak = 2.0 * bf16_sustained
        roof = {"bound": "tensor", "achieved": tensor_tops, "peak": peak, "unit": "TOP/s (int8)",
                "frac": tensor_tops / peak, "traffic": measured_traffic,
                "kernel": "gram_tc_kernel<%d, KindI8, TanimotoF32x2, false>" % (2 if args.engine == "umma2" else 1),
                "algorithmic_per_unit": "4096 int8-ops/pair (2*D) ; 4 B/pair written",
                "peak_source": peak_src,
                "hbm_write_gbs": 4.0 * pairs_per_step / sec / 1e9,
                "hbm_write_frac_of_measured": 4.0 * pairs_per_step / sec / 1e9 / hbm_peak}
    else:
        bytes_per_launch = 4.0 * pairs_per_step
        achieved = bytes_per_launch / sec / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "kernel": "gram_words_kernel", "algorithmic_per_unit": "4 B/pair written",
                "peak_source": "hbm_gbs of MEASURED_PEAKS.json"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rows = min(args.cpu_sample_rows, R)
        gram_step(0)
        torch.cuda.synchronize()
        gram_head = gram[0][:512].cpu()                      # 512 x M Gram rows of the GPU arm
        from gauche_b200.screening import TanimotoScreener as _Scr
        gpu_scores = _Scr(train, alpha, bias=0.0, block_rows=R, engine=args.engine).scores(cand).cpu()
        train_cpu = (train.bits.cpu().numpy().view("uint8"))  # rebuild the dense train set from the packed bits
        import numpy as np
        train_cpu = torch.from_numpy(np.unpackbits(train_cpu, axis=1, bitorder="little")[:, :D_BITS].astype("float32"))
        cpu = time_cpu_baseline(args, rows, train_cpu, cand_dense.cpu(), alpha.cpu(), gram_head, gpu_scores)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("e2m1 bits (exact f32 accumulation)" if args.engine in ("bx", "mxf4x1") else "u8 (exact s32 accumulation)") + ", fp32 epilogue",
        "data": "synthetic", "config": workload_config(args, world),
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": args.steps, "pack_ms_per_block": pack_ms,
        "per_launch_ms": {"mean": avg_launch_ms, "min": min(per_launch_ms), "max": max(per_launch_ms)},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
