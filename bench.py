#!/usr/bin/env python
"""bench.py -- brand x post pairs scored + ranked per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on host cores

Workload (BASELINE.json configs[1], per GPU; posts shard across GPUs => weak scaling):
    1 000 brands x 1 000 000 posts, 2048-d visual + 1024-d text fp32 embeddings, A = 2000 aspects,
    score + top-100 + NDCG@10/50 (+ recall@1/5/10, MedR, MeanR from the same integers).
One step = one full evaluation pass: brand embed -> post finalisation (per-branch l2norm, concat, row
l2norm, bf16) -> fused tcgen05 score + top-k GEMM -> rank statistics -> host float64 aggregation
(multi-GPU: + all-gather of the candidate lists and merge).  Synthetic data, seeded
(seed = 20261018 + 1000*config + rank).  `value` has the inputs resident in HBM and runs the K steps as a
pipelined stream of evaluations (fancyrec_b200/pipeline.py: the device is never idle -- the packed statistics of step t
are copied to pinned memory asynchronously and aggregated on the host while step t+1 runs; the region starts and ends
with an idle device, so pipeline fill and drain are inside it); `latency_ms` is one isolated synchronous step.  `e2e` starts from pinned
HOST buffers every step (H2D inside the timed region) and ends with the metrics on the host.
`extra` carries the other BASELINE.json configs: c3 (loss tile), c4 (10 k brands x 20 M posts, top-1000, strong-sharded
over the ranks), c5 (video pooling + 5 k x 5 M evaluation), the exact-AUC evaluation, and for N > 1 `sharded_check`
(a 128-brand slab recomputed on ONE GPU from the gathered operands must equal the NCCL-merged result).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "brand_post_pairs_scored_and_ranked_per_sec"
UNIT = "pairs/s"
CFG = dict(nb=1000, np_per_gpu=1000000, dv=2048, dt=1024, aspects=2000, k=100, signal=0.05)
SEED0 = 20261018 + 1000 * 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--posts-per-gpu", type=int, default=CFG["np_per_gpu"])
    ap.add_argument("--brands", type=int, default=CFG["nb"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--cpu-sample-posts", type=int, default=200000)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("FRX_CLOCK_MS", "20")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived in [t0, t1] (perf_counter); all samples when none did."""
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [l for (t, l) in self.lines if t0 is not None and t0 <= t <= t1]
        window = "timed region"
        if not inside:
            inside, window = [l for (_, l) in self.lines], "warm-up + timed region (timed region shorter than the sampling period)"
        for line in inside:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm), window=window)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------
# synthetic workload (generated on the device, chunk-wise)
# ---------------------------------------------------------------------------------------------
def make_workload(dev, rank, nb, n_local, cfg):
    g = torch.Generator(device=dev).manual_seed(SEED0)          # brand tables identical on every rank
    a, d = cfg["aspects"], cfg["dv"] + cfg["dt"]
    w = torch.randn((nb + 1, a), generator=g, device=dev)
    e = torch.randn((a, d), generator=g, device=dev)
    from fancyrec_b200 import ops
    brand = ops.brand_embed(w, e, nb=nb)
    g = torch.Generator(device=dev).manual_seed(SEED0 + 1 + rank)
    labels = (torch.randperm(n_local, generator=g, device=dev) % nb).to(torch.int32)
    visual = torch.empty((n_local, cfg["dv"]), device=dev)
    text = torch.empty((n_local, cfg["dt"]), device=dev)
    bv = brand[:, :cfg["dv"]] / brand[:, :cfg["dv"]].norm(dim=1, keepdim=True)
    bt = brand[:, cfg["dv"]:] / brand[:, cfg["dv"]:].norm(dim=1, keepdim=True)
    chunk = 65536
    for lo in range(0, n_local, chunk):
        hi = min(n_local, lo + chunk)
        lab = labels[lo:hi].long()
        visual[lo:hi] = torch.randn((hi - lo, cfg["dv"]), generator=g, device=dev) + \
            cfg["signal"] * (cfg["dv"] ** 0.5) * bv[lab]
        text[lo:hi] = torch.randn((hi - lo, cfg["dt"]), generator=g, device=dev) + \
            cfg["signal"] * (cfg["dt"] ** 0.5) * bt[lab]
    return w, e, labels, visual, text


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
# our kernels per step (profiles/r02_launches.csv): brand_embed = split_rows + split_transpose + 3xTF32 GEMM | 2 x finalize |
# sample pass: dense score + row k-th select | main: fused score + merge | label_stats, decode_best, rank_from_topk |
# missing_thresholds, score_count (returns at once unless a first positive is missing), pack_rank_stats |
# sharded: merge of the gathered lists + reduce_shard_stats
def launches_per_step(world):
    return 3 + 2 + 2 + 2 + 3 + 3 + (2 if world > 1 else 0)


def config_dict(nb, n_local, world, cfg, numa_cores=None):
    d = cfg["dv"] + cfg["dt"]
    return {"workload": "configs[1]: %d brands x %d posts per GPU, 2048-d visual + 1024-d text fp32 "
                        "embeddings, A=2000 aspects, score + top-100 + NDCG@10/50 (+recall@k, MedR)" % (nb, n_local),
            "brands": nb, "posts_per_gpu": n_local, "posts_total": n_local * world, "dim": d, "k": cfg["k"],
            "sharding": "posts x %d, one all-gather of top-k lists" % world,
            "cache": "inputs larger than L2 (12.3 GB fp32 + 6.1 GB bf16 per step vs 126 MB)",
            "seed": SEED0, "host_cores_bound_to_gpu_numa_node": numa_cores}


def barrier_sync(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, dev, world):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def _claim_stdout():
    """Route fd 1 to stderr for the rest of the process (NCCL prints its banner / INFO lines on stdout) and
    return a writer for the ONE JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return lambda text: os.write(saved, (text + "\n").encode())


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPU cores NVML reports as local to its GPU, so that the pinned host buffers of the
    e2e leg are allocated on the GPU's own NUMA node (8 ranks sharing one socket's memory halve the H2D rate)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        try:
            bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def timed_pipeline(pipe, inputs, steps, world, dev):
    """K evaluations through the pipeline; idle device on both sides of the timed region; every result collected
    (host aggregation included) before the closing event.  Returns (ms total max over ranks, last result, t0, t1)."""
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    t0 = time.perf_counter()
    beg.record()
    prev, result = None, None
    for _ in range(steps):
        tk = pipe.submit(*inputs)
        if prev is not None:
            result = pipe.result(prev)
        prev = tk
    result = pipe.result(prev)
    end.record()
    barrier_sync(world)
    t1 = time.perf_counter()
    return max_over_ranks(beg.elapsed_time(end), dev, world), result, t0, t1


def run_ours(args):
    emit = _claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        # NCCL_DEBUG is left as the caller set it (the driver reads the communicator's rank count from NCCL's log);
        # _claim_stdout() already keeps NCCL's lines off the JSON line's file descriptor
        torch.distributed.init_process_group("nccl", device_id=dev)
    from fancyrec_b200 import _lib, ops, pipeline, ranking
    lib = _lib.load()
    cfg = dict(CFG)
    nb, n_local = args.brands, args.posts_per_gpu
    d = cfg["dv"] + cfg["dt"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                  # nvidia-smi needs a few 100 ms to start reporting: start before the workload is built
    w, e, labels, visual, text = make_workload(dev, rank, nb, n_local, cfg)
    overlap = os.environ.get("FRX_OVERLAP", "0") == "1"   # measured slower on B200: off
    pipe = pipeline.EvalPipeline(dev, nb, n_local, cfg["dv"], cfg["dt"], k=cfg["k"], n_posts_total=n_local * world,
                                 want_auc=False, overlap=overlap)
    inputs = (w, e, visual, text, labels)
    pk = peaks()
    warmup = max(args.warmup, 3)

    for _ in range(warmup):
        result = pipe.result(pipe.submit(*inputs))
    barrier_sync(world)

    # ---- timed region: K steps, device resident inputs (12.3 GB of fp32 inputs per step >> 126 MB L2)
    lib.frx_probe_enable(1)
    profile_range = bool(os.environ.get("FRX_PROFILE_RANGE"))   # ncu --profile-from-start off
    if profile_range:
        torch.cuda.profiler.start()
    ms_total, result, t_start, t_end = timed_pipeline(pipe, inputs, args.steps, world, dev)
    if profile_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    buf = (torch.zeros(4096, dtype=torch.float32)).numpy()
    n_probe = lib.frx_probe_read(buf.ctypes.data, 4096)
    lib.frx_probe_enable(0)
    # one probed launch per step: the main fused score + top-k kernel
    topk_ms = float(np.mean(buf[:n_probe])) if n_probe else float("nan")
    ms_step = ms_total / args.steps
    pairs = float(nb) * float(n_local) * world
    value = pairs / (ms_step * 1e-3)
    flops = 2.0 * nb * n_local * d
    achieved = flops / (topk_ms * 1e-3) / 1e12

    # ---- the same K steps one at a time (submit, wait for the result): the latency of an isolated evaluation,
    # and the contraction kernel alone on the device (no co-resident finalisation)
    lat_steps = max(3, min(args.steps, 10))
    lib.frx_probe_enable(1)
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    beg.record()
    for _ in range(lat_steps):
        result_sync = pipe.result(pipe.submit(*inputs))
    end.record()
    barrier_sync(world)
    latency_ms = max_over_ranks(beg.elapsed_time(end), dev, world) / lat_steps
    n_probe = lib.frx_probe_read(buf.ctypes.data, 4096)
    lib.frx_probe_enable(0)
    topk_alone_ms = float(np.mean(buf[:n_probe])) if n_probe else float("nan")
    assert all(float(a) == float(b) or (a != a and b != b) for a, b in zip(result_sync, result)), \
        "pipelined and synchronous results differ"              # (AUC is NaN in both: not part of configs[1])

    sharded = None
    if world > 1:
        sharded = sharded_check(pipe, labels, dev, world, rank, nb, n_local, d, cfg["k"])

    # ---- e2e: pinned host inputs -> H2D (chunked, overlapped with finalisation) -> metrics on host
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(pipe, w, e, labels, visual, text, args, dev, world, nb, n_local, cfg)
        except RuntimeError as ex:       # e.g. the host cannot pin 12.3 GB per rank
            e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    extra = {}
    if not args.no_extras:
        if world == 1:
            extra["auc"] = guarded(extra_auc, dev, nb, n_local, cfg, inputs, result)
        del pipe, inputs, w, e, labels, visual, text
        torch.cuda.empty_cache()
        if world == 1:
            extra["c3"] = guarded(extra_c3, dev, pk)
            extra["c5"] = guarded(extra_c5, dev, pk)
        extra["c4"] = guarded(extra_c4, dev, world, rank, pk)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.brands, args.cpu_sample_posts, cfg, steps=1)
        cpu["c1_reference"] = reference_c1()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(nb, n_local, world, cfg, numa_cores),
            "clocks": clocks,
            "gpu_launches": launches_per_step(world) * args.steps,
            "latency_ms": latency_ms,
            "schedule": ("pipelined, FRX_OVERLAP=1: finalisation of step t+1 on a side stream under the contraction of step t, "
                         "host aggregation of step t-1 meanwhile" if overlap else
                         "pipelined on one stream: async D2H of the packed statistics, host aggregation of step t-1 while "
                         "step t runs") +
                        "; idle device before and after the timed region; latency_ms = one isolated step",
            "roofline": roofline_block(pk, achieved, topk_ms, topk_alone_ms, ms_step, flops),
            "metrics_sample": {"MedR": float(result[0]), "MeanR": float(result[1]), "NDCG@10": float(result[3]),
                               "NDCG@50": float(result[4]), "r1": result[5], "r5": result[6], "r10": result[7]},
        }
        if sharded is not None:
            line["sharded_check"] = sharded["status"]
            line["sharded_check_detail"] = sharded
        if e2e is not None:
            line["e2e"] = e2e
        if extra:
            line["extra"] = extra
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def roofline_block(pk, achieved, topk_ms, topk_alone_ms, ms_step, flops):
    """The dominant kernel against the tensor roofline.  `frac` is taken against the BURST bf16 figure (the kernel runs
    for a few ms inside a step that alternates with HBM-bound work; the sustained figure was measured at a lower,
    power-settled clock and would flatter it).  `traffic`: DRAM bytes per launch from the committed ncu capture of THIS
    build (profiles/score_topk_traffic.json carries the library stamp it was taken with), else null."""
    traffic, traffic_note = None, "no ncu capture of this build committed"
    prof = os.path.join(ROOT, "profiles", "score_topk_traffic.json")
    if os.path.exists(prof):
        from fancyrec_b200 import _lib
        t = json.load(open(prof))
        if t.get("library_stamp") == _lib.source_hash():
            traffic, traffic_note = t.get("dram_bytes_per_launch"), t.get("source")
        else:
            traffic_note = "profiles/score_topk_traffic.json is from another build (stamp differs): not reported"
    return {"bound": "tensor", "kernel": "frx::score_kernel<MODE_TOPK> (tcgen05 bf16, fused top-k)",
            "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"],
            "frac_of_sustained_peak": achieved / pk["tf_sustained"],
            "peak_source": pk["source"] + ", burst bf16",
            "kernel_ms": topk_ms, "kernel_ms_alone": topk_alone_ms,
            "kernel_note": "kernel_ms: inside the pipelined (back-to-back, power-settled) steps; kernel_ms_alone: the same "
                           "launch in the isolated steps of latency_ms",
            "kernel_share_of_step": topk_ms / ms_step,
            "step_frac_of_sustained_peak": flops / (ms_step * 1e-3) / 1e12 / pk["tf_sustained"],
            "step_frac_of_burst_peak": flops / (ms_step * 1e-3) / 1e12 / pk["tf_burst"],
            "algorithmic_flops_per_launch": flops, "traffic": traffic, "traffic_source": traffic_note}


def guarded(fn, *a):
    """An extra must never cost the headline line: failures are reported in place."""
    try:
        torch.cuda.synchronize()
        out = fn(*a)
        torch.cuda.synchronize()
        return out
    except Exception as ex:  # noqa: BLE001
        torch.cuda.empty_cache()
        return {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}


def sharded_check(pipe, labels, dev, world, rank, nb, n_local, d, k):
    """Proof that the NCCL exchange + merge computes the right thing, inside the run: every rank's finalised operand and
    labels are gathered on rank 0, which recomputes a 128-brand slab of the WHOLE job on ONE GPU (no collective, no merge
    of shards) and compares top-k lists, positive counts, best positives, hit masks and first-positive ranks with the
    merged statistics of the last timed step."""
    import torch.distributed as dist
    from fancyrec_b200 import ops
    st = pipe.last_stats
    slot = (pipe.ticket - 1) % pipe.depth
    post_op = pipe.post_op[slot]
    slab = min(128, nb)
    big = torch.empty((world * n_local, post_op.shape[1]), dtype=post_op.dtype, device=dev) if rank == 0 else None
    lab_all = torch.empty(world * n_local, dtype=torch.int32, device=dev) if rank == 0 else None
    dist.gather(post_op, [big[r * n_local:(r + 1) * n_local] for r in range(world)] if rank == 0 else None, dst=0)
    dist.gather(labels, [lab_all[r * n_local:(r + 1) * n_local] for r in range(world)] if rank == 0 else None, dst=0)
    out = {"status": "ok", "brands_checked": slab, "posts": world * n_local}
    if rank == 0:
        brand_op = pipe.last_brand_op[:slab].contiguous()
        kk = st["topk_index"].shape[1]
        res = ops.score_topk(brand_op, big, kk, d=d, labels=lab_all)
        n_pos, best_s, best_i = ops.label_stats(lab_all, res["pos_score"], slab)
        hit, first = ops.rank_from_topk(res["index"], lab_all)
        checks = {"topk_index": torch.equal(res["index"], st["topk_index"][:slab]),
                  "topk_scores": torch.equal(res["scores"], st["topk_scores"][:slab]),
                  "n_pos": torch.equal(n_pos, st["n_pos"][:slab]),
                  "best_index": torch.equal(best_i, st["best_index"][:slab]),
                  "hit_mask": torch.equal(hit, st["hit_mask"][:slab]),
                  "first_in_list": torch.equal(first, st["first_in_list"][:slab])}
        out["checks"] = checks
        if not all(checks.values()):
            out["status"] = "MISMATCH"
        del big, lab_all
        torch.cuda.empty_cache()
    flag = torch.tensor([1 if out["status"] == "ok" else 0], device=dev)
    dist.broadcast(flag, src=0)
    out["status"] = "ok" if int(flag.item()) == 1 else "MISMATCH"
    return out


def run_e2e(pipe, w, e, labels, visual, text, args, dev, world, nb, n_local, cfg):
    """Same step, but the post embeddings and labels start in pinned HOST memory every step."""
    from fancyrec_b200 import ingest, ops, ranking, sharded
    dv, dt = cfg["dv"], cfg["dt"]
    d = dv + dt
    h_visual = torch.empty((n_local, dv), dtype=torch.float32, pin_memory=True)
    h_text = torch.empty((n_local, dt), dtype=torch.float32, pin_memory=True)
    h_labels = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    h_visual.copy_(visual); h_text.copy_(text); h_labels.copy_(labels)
    torch.cuda.synchronize()
    post_op = pipe.post_op[0]
    d_labels = torch.empty(n_local, dtype=torch.int32, device=dev)
    n_total = n_local * world

    def step():
        brand = ops.brand_embed(w, e, nb=nb)
        brand_op = ops.finalize_posts(brand, final_norm=True)[1]
        d_labels.copy_(h_labels, non_blocking=True)
        # pinned host rows -> chunked H2D on a copy stream, overlapped with the finalisation of the previous chunk
        ingest.finalize_from_host(h_visual, h_text, visual_norm=True, text_norm=True, final_norm=True,
                                  out_bf16=post_op, device=dev, chunk_posts=131072)
        st = sharded.sharded_rank_statistics(brand_op, post_op, d_labels, d, cfg["k"], n_total, workspace=pipe.workspace)
        stats = ranking.host_statistics(st, n_total, want_auc=False)
        return ranking.aggregate(stats, n_total, want_auc=False)

    for _ in range(2):
        step()
    steps = max(2, min(args.steps, 5))
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    beg.record()
    for _ in range(steps):
        step()
    end.record()
    barrier_sync(world)
    ms = max_over_ranks(beg.elapsed_time(end), dev, world) / steps
    pairs = float(nb) * float(n_local) * world
    h2d = n_local * (dv + dt) * 4 + n_local * 4
    d2h = nb * 5 * 8                      # the packed int64 [5, NB] statistics block
    return {"value": pairs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "api": "ops.brand_embed + ingest.finalize_from_host (pinned host -> device, chunked, overlapped) + "
                   "sharded.sharded_rank_statistics + ranking.aggregate: the call chain of evaluator.test_post_ranking "
                   "WITHOUT its exact-AUC sweep (configs[1] names score + top-100 + NDCG; extra.auc has the AUC-inclusive "
                   "evaluation), H2D-bound at the PCIe rate"}


# ---------------------------------------------------------------------------------------------
# extras: the other BASELINE.json configs, measured in the same run (not the headline value)
# ---------------------------------------------------------------------------------------------
def _event_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    beg.record()
    for _ in range(reps):
        fn()
    end.record()
    torch.cuda.synchronize()
    return beg.elapsed_time(end) / reps


def _graph_ms(fn, reps=50):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    beg.record()
    for _ in range(reps):
        g.replay()
    end.record()
    torch.cuda.synchronize()
    return beg.elapsed_time(end) / reps


def extra_auc(dev, nb, n_local, cfg, inputs, result_no_auc):
    """The evaluation the reference's test_post_ranking actually runs (AUC always on, evaluator.py:103-118): the fused
    pass also writes the scores, one streaming pass counts every (positive, negative) pair exactly."""
    from fancyrec_b200 import ops, pipeline, ranking
    w, e, visual, text, labels = inputs
    pipe = pipeline.EvalPipeline(dev, nb, n_local, cfg["dv"], cfg["dt"], k=cfg["k"], want_auc=True, overlap=False)
    for _ in range(2):
        res = pipe.result(pipe.submit(*inputs))
    ms_total, res, _, _ = timed_pipeline(pipe, inputs, 5, 1, dev)
    same = all(float(a) == float(b) for a, b in zip(res[:2] + res[3:], result_no_auc[:2] + result_no_auc[3:]))
    del pipe
    torch.cuda.empty_cache()

    def sync_call():          # the synchronous drop-in: finalised rows in, 8-tuple out (evaluator.test_post_ranking's body)
        brand = ops.brand_embed(w, e, nb=nb)
        posts = ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=False, want_f32=True,
                                   want_bf16=False)[0]
        return ranking.rank_posts(brand, posts, labels, k=cfg["k"], want_auc=True)[0]
    t_sync = _event_ms(sync_call, 3, warm=1)
    return {"ms_per_step_pipelined": ms_total / 5, "pairs_per_s": float(nb) * n_local / (ms_total / 5 * 1e-3),
            "test_post_ranking_ms": t_sync, "AUC": float(res[2]), "other_metrics_equal_no_auc_run": bool(same),
            "what": "exact AUC numerators (u64) for %d brands x %d posts on top of the headline step" % (nb, n_local)}


def extra_c3(dev, pk):
    """configs[2]: batch 512, D = 3072, hinge (TripletLoss) and contrastive loss, forward + backward through the nn.Module."""
    from types import SimpleNamespace
    from fancyrec_b200 import loss, loss_ctrs
    b, d = 512, 3072
    g = torch.Generator(device=dev).manual_seed(SEED0 + 3)
    ids = torch.randint(0, 51, (b,), generator=g, device=dev)
    brand = torch.randn((b, d), generator=g, device=dev, requires_grad=True)
    post = torch.randn((b, d), generator=g, device=dev, requires_grad=True)
    trip = loss.TripletLoss(margin=0.2, cost_style="sum").to(dev)
    opt = SimpleNamespace(cost_style="sum", queue_size=5120, common_embedding_size=d, no_queue=False, no_intra=False)
    con = loss_ctrs.ContrastiveLoss(opt).to(dev)

    def run(mod, *a):
        def f():
            brand.grad = None; post.grad = None
            mod(*a).backward()
        return f
    def fwd_only(mod, *a):
        def f():
            with torch.no_grad():
                mod(*a)
        return f
    t_ms = _event_ms(run(trip, ids, brand, post), 50, warm=5)
    t_fwd = _event_ms(fwd_only(trip, ids, brand, post), 50, warm=5)
    c_ms = _event_ms(run(con, brand, post), 50, warm=5)
    # device time of the same work with the host out of the way: the C entry point captured in a CUDA graph and replayed
    from fancyrec_b200 import ops
    bd, pd = brand.detach(), post.detach()
    t_dev = _graph_ms(lambda: ops.triplet_fwd_bwd(ids, bd, pd, 0.2, 0, True))
    t_dev_fwd = _graph_ms(lambda: ops.triplet_fwd_bwd(ids, bd, pd, 0.2, 0, False))
    c_dev = _graph_ms(lambda: ops.contrastive_fwd_bwd(bd, pd, con.queue, 0, False, 0.03, 0.8, 0, True))
    # bound: max(flops / tensor peak, bytes / HBM peak); 6 B^2 D flop (three B x B x D contractions), 4 B D fp32 arrays
    flop_us = 6.0 * b * b * d / (pk["tf_burst"] * 1e12) * 1e6
    byte_us = 4.0 * 4 * b * d / (pk["hbm"] * 1e9) * 1e6
    bound = max(flop_us, byte_us)
    q = opt.queue_size
    c_bound = max((6.0 * b * b * d + 4.0 * b * q * d) / (pk["tf_burst"] * 1e12) * 1e6,
                  4.0 * (4 * b * d + q * d) / (pk["hbm"] * 1e9) * 1e6)
    return {"batch": b, "dim": d, "triplet_fwd_bwd_us": t_ms * 1e3, "triplet_fwd_only_us": t_fwd * 1e3,
            "triplet_device_us": t_dev * 1e3, "triplet_device_fwd_only_us": t_dev_fwd * 1e3, "triplet_launches": 3,
            "triplet_bound_us": bound, "triplet_frac_of_bound": bound / (t_dev * 1e3),
            "contrastive_fwd_bwd_us": c_ms * 1e3, "contrastive_device_us": c_dev * 1e3, "contrastive_queue": q,
            "contrastive_bound_us": c_bound, "contrastive_frac_of_bound": c_bound / (c_dev * 1e3),
            "how": "*_us: nn.Module forward + .backward() from Python, CUDA events over 50 iterations (host-bound: "
                   "autograd.Function + ctypes + launches); *_device_us: the same C entry point captured in a CUDA graph "
                   "and replayed 50 times (device time only); frac_of_bound uses the device time"}


def _random_operand(n, d, gen, dev, brand_dir=None, labels=None, signal=0.05, chunk=65536):
    """Unit-norm bf16 rows [n, round_up(d, 64)] generated chunk-wise: N(0,1) + signal * sqrt(d) * brand direction."""
    from fancyrec_b200 import ops
    out = torch.empty((n, ops.round_up(d, 64)), dtype=torch.bfloat16, device=dev)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        x = torch.randn((hi - lo, d), generator=gen, device=dev)
        if brand_dir is not None:
            x += signal * d ** 0.5 * brand_dir[labels[lo:hi].long()]
        ops.finalize_posts(x, final_norm=True, out_bf16=out[lo:hi])
    return out


def _eval_once(brand_op, post_op, labels, d, k, n_total, ws, events=None):
    from fancyrec_b200 import ranking, sharded
    st = sharded.sharded_rank_statistics(brand_op, post_op, labels, d, k, n_total, workspace=ws, events=events)
    stats = ranking.host_statistics(st, n_total, want_auc=False)
    return ranking.aggregate(stats, n_total, want_auc=False), st


def extra_c5(dev, pk):
    """configs[4]: (a) 32 frames x 2048-d fp32 per post, mean-pool + L2 norm + bf16 from feature.bin-layout rows
    (HBM-bound); (b) the full 5 000 brands x 5 000 000 posts evaluation sweep (recall@k / MedR / MeanR / NDCG) on ONE GPU."""
    from fancyrec_b200 import ops
    g = torch.Generator(device=dev).manual_seed(SEED0 + 5)
    n_pool, f, dv = 100000, 32, 2048
    frames = torch.empty((n_pool * f, dv), device=dev)
    for lo in range(0, n_pool * f, 1 << 18):
        hi = min(n_pool * f, lo + (1 << 18))
        frames[lo:hi] = torch.relu(torch.randn((hi - lo, dv), generator=g, device=dev) * 0.5 + 0.3)
    row_ptr = torch.arange(n_pool + 1, device=dev, dtype=torch.int64) * f
    out = torch.empty((n_pool, dv), dtype=torch.bfloat16, device=dev)
    ms = _event_ms(lambda: ops.finalize_posts(frames, row_ptr=row_ptr, final_norm=True, out_bf16=out), 5, warm=2)
    pool_bytes = n_pool * (4.0 * f * dv + 2.0 * dv)
    pool = {"posts": n_pool, "frames_per_post": f, "ms": ms, "gb_per_s": pool_bytes / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": pool_bytes / (ms * 1e-3) / 1e9 / pk["hbm"], "posts_per_s": n_pool / (ms * 1e-3),
            "full_config_estimate_s": 5e6 / (n_pool / (ms * 1e-3))}
    del frames, out
    torch.cuda.empty_cache()
    nb, n, d, k = 5000, 5000000, 2048, 100
    brand = torch.randn((nb, d), generator=g, device=dev)
    bdir = brand / brand.norm(dim=1, keepdim=True)
    labels = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
    post_op = _random_operand(n, d, g, dev, bdir, labels)
    brand_op = ops.finalize_posts(brand, final_norm=True)[1]
    ws = torch.empty(_lib_ws(nb, n, d, k), dtype=torch.uint8, device=dev)
    res = [None]
    def one():
        res[0] = _eval_once(brand_op, post_op, labels, d, k, n, ws)[0]
    ms_eval = _event_ms(one, 3, warm=1)
    r = res[0]
    return {"pool": pool,
            "eval": {"brands": nb, "posts": n, "dim": d, "k": k, "ms": ms_eval, "pairs_per_s": float(nb) * n / (ms_eval * 1e-3),
                     "frac_of_burst_peak": 2.0 * nb * n * d / (ms_eval * 1e-3) / 1e12 / pk["tf_burst"],
                     "metrics": {"MedR": float(r[0]), "MeanR": float(r[1]), "NDCG@10": float(r[3]), "NDCG@50": float(r[4]),
                                 "r1": r[5], "r5": r[6], "r10": r[7]}}}


def _lib_ws(nb, n, d, k):
    from fancyrec_b200 import _lib
    return int(_lib.load().frx_score_topk_workspace_bytes(nb, n, d, max(k, 64)))


def extra_c4(dev, world, rank, pk):
    """configs[3]: 10 000 brands x 20 000 000 posts, D = 3072, top-1000, STRONG-sharded: the posts are split over the
    ranks (all 20 M on one GPU at N = 1: 123 GB of bf16 operand), fused GEMM-epilogue top-1000 per shard, one packed
    all-gather of the candidate lists + label statistics + labels, merge, rank statistics, float64 aggregation."""
    from fancyrec_b200 import ops, sharded
    nb, n_total, d, k = 10000, 20000000, 3072, 1000
    lo, hi = sharded.shard_bounds(n_total, world, rank)
    n_local = hi - lo
    need = n_local * d * 2 + _lib_ws(nb, n_local, d, k) + (8 << 30)
    free, _ = torch.cuda.mem_get_info()
    if free < need:
        return {"skipped": "needs %.0f GB of HBM per GPU, %.0f GB free" % (need / 2 ** 30, free / 2 ** 30)}
    g = torch.Generator(device=dev).manual_seed(SEED0 + 4)                # brands identical on every rank
    brand = torch.randn((nb, d), generator=g, device=dev)
    bdir = brand / brand.norm(dim=1, keepdim=True)
    brand_op = ops.finalize_posts(brand, final_norm=True)[1]
    g = torch.Generator(device=dev).manual_seed(SEED0 + 40 + rank)
    labels = ((torch.arange(lo, hi, device=dev) * 7919) % nb).to(torch.int32)      # 2 000 positives per brand, spread
    post_op = _random_operand(n_local, d, g, dev, bdir, labels)
    ws = torch.empty(_lib_ws(nb, n_local, d, k), dtype=torch.uint8, device=dev)
    _eval_once(brand_op, post_op, labels, d, k, n_total, ws)               # warm-up
    reps, ms_all, ev_last, res = 2, [], None, None
    for _ in range(reps):
        events = {}
        beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier_sync(world)
        beg.record()
        res, st = _eval_once(brand_op, post_op, labels, d, k, n_total, ws, events)
        end.record()
        barrier_sync(world)
        ms_all.append(max_over_ranks(beg.elapsed_time(end), dev, world))
        ev_last = events
    ms = min(ms_all)
    out = {"brands": nb, "posts_total": n_total, "posts_per_gpu": n_local, "dim": d, "k": k, "n_gpus": world,
           "scaling": "strong", "ms": ms, "ms_all": ms_all, "pairs_per_s": float(nb) * n_total / (ms * 1e-3),
           "frac_of_burst_peak_per_gpu": 2.0 * nb * n_total * d / world / (ms * 1e-3) / 1e12 / pk["tf_burst"],
           "metrics": {"MedR": float(res[0]), "MeanR": float(res[1]), "NDCG@10": float(res[3]), "NDCG@50": float(res[4]),
                       "r1": res[5], "r5": res[6], "r10": res[7]}}
    if world > 1 and ev_last and "exchange_begin" in ev_last:
        out["exchange_bytes"] = ev_last["exchange_bytes_received"]
        out["exchange_bytes_sent"] = ev_last["exchange_bytes_sent"]
        out["exchange_ms"] = ev_last["exchange_begin"].elapsed_time(ev_last["exchange_end"])
        out["merge_ms"] = ev_last["exchange_end"].elapsed_time(ev_last["merge_end"])
    return out


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's algorithm on the host cores
# ---------------------------------------------------------------------------------------------
def _cpu_workload(nb, n_posts, cfg, seed):
    rs = np.random.RandomState(seed)
    d = cfg["dv"] + cfg["dt"]
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    lab = (rs.permutation(n_posts) % nb).astype(np.int64)
    visual = rs.standard_normal((n_posts, cfg["dv"])).astype(np.float32)
    text = rs.standard_normal((n_posts, cfg["dt"])).astype(np.float32)
    bn = brand / np.linalg.norm(brand, axis=1, keepdims=True)
    visual += np.float32(cfg["signal"] * np.sqrt(d)) * bn[lab][:, :cfg["dv"]]
    text += np.float32(cfg["signal"] * np.sqrt(d)) * bn[lab][:, cfg["dv"]:]
    return brand, lab, visual, text


def _cpu_step(brand, lab, visual, text, k, pool, n_threads):
    """The reference's path for this workload (evaluator.py cal_sim + per-brand sort + NDCG + first
    positive rank; model.py per-branch l2norm + concat), restated in oracle/, all host threads."""
    from oracle import embed as oembed
    from oracle import ranking as oref
    from oracle.ndcg import ndcg_from_hits
    posts = oembed.finalize_posts(visual, text, True, True, False)
    scores = torch.mm(torch.from_numpy(oref.l2norm(brand)), torch.from_numpy(oref.l2norm(posts)).t()).numpy()
    nb, n_posts = scores.shape

    def one(rows):
        out = []
        for b in rows:
            order = oref.order_desc(scores[b])
            rel = lab[order] == b
            n_pos = int(rel.sum())
            out.append((b, n_pos, int(np.argmax(rel)) if n_pos else -1, rel[:50].copy(), order[:k].copy()))
        return out

    chunks = [range(i, nb, n_threads) for i in range(n_threads)]
    res = [r for part in pool.map(one, chunks) for r in part]
    first = [r[2] for r in res if r[1]]
    n10 = [ndcg_from_hits(r[3], r[1], 10, n_posts) for r in res if r[1]]
    n50 = [ndcg_from_hits(r[3], r[1], 50, n_posts) for r in res if r[1]]
    return np.floor(np.median(first)), np.average(n10), np.average(n50)


def cpu_baseline(nb, sample_posts, cfg, steps=1, warmup=0):
    from concurrent.futures import ThreadPoolExecutor
    n_threads = os.cpu_count() or 1
    torch.set_num_threads(n_threads)
    brand, lab, visual, text = _cpu_workload(nb, sample_posts, cfg, SEED0 + 7)
    with ThreadPoolExecutor(n_threads) as pool:
        for _ in range(warmup):
            _cpu_step(brand, lab, visual, text, cfg["k"], pool, n_threads)
        t0 = time.perf_counter()
        for _ in range(steps):
            _cpu_step(brand, lab, visual, text, cfg["k"], pool, n_threads)
        dt = (time.perf_counter() - t0) / steps
    return {"value": nb * sample_posts / dt, "unit": UNIT, "cores": n_threads, "kind": "port",
            "seconds_per_step": dt,
            "sample": "%d brands x %d posts (%.0f%% of one GPU's posts; pairs/s is per pair, no extrapolation), same dims / "
                      "k / metrics; oracle/ NumPy restatement of evaluator.py + torch.mm on %d host threads"
                      % (nb, sample_posts, 100.0 * sample_posts / CFG["np_per_gpu"], n_threads)}


def reference_c1():
    """BASELINE.json configs[0] on the UNMODIFIED reference (oracle/_ref: the reference's own evaluator.test_post_ranking
    and BrandAspects, byte-compiled from /root/reference by oracle/build_ref.py) in its documented CPU mode gpu = -1,
    i.e. a child process with CUDA_VISIBLE_DEVICES=-1 (README.md:64, bin/instance.sh:29-30): 50 brands x 10 000 posts,
    2048-d + 1024-d rows, A = 2000, full 8-tuple incl. AUC."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="-1")
    try:
        p = subprocess.run([sys.executable, "-c", "import json, bench; print('C1JSON' + json.dumps(bench._reference_c1_inproc()))"],
                           cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        for line in p.stdout.splitlines():
            if line.startswith("C1JSON"):
                return json.loads(line[6:])
        return {"kind": "reference", "unavailable": (p.stderr or p.stdout)[-200:]}
    except Exception as ex:  # noqa: BLE001
        return {"kind": "reference", "unavailable": str(ex)[:200]}


def _reference_c1_inproc():
    try:
        from oracle import build_ref
        ref = build_ref.load()
    except Exception as ex:  # noqa: BLE001
        return {"kind": "reference", "unavailable": str(ex)[:200]}
    from types import SimpleNamespace
    n_threads = os.cpu_count() or 1
    torch.set_num_threads(n_threads)
    nb, n_posts, dv, dt, a = 50, 10000, 2048, 1024, 2000
    d = dv + dt
    gen = torch.Generator().manual_seed(SEED0 - 1000)
    opt = SimpleNamespace(brand_num=nb, brand_aspect=a, common_embedding_size=d, dropout=0.5)
    enc = ref.model.BrandAspects(opt)
    with torch.no_grad():
        enc.brand_embeddings.weight.copy_(torch.randn(enc.brand_embeddings.weight.shape, generator=gen))
        enc.aspects_embeddings.copy_(torch.randn(enc.aspects_embeddings.shape, generator=gen))
    model = SimpleNamespace(brand_encoding=enc)
    labels = (torch.randperm(n_posts, generator=gen) % nb).long()
    visual = torch.randn((n_posts, dv), generator=gen)
    text = torch.randn((n_posts, dt), generator=gen)
    t0 = time.perf_counter()
    with torch.no_grad():
        posts = torch.cat((ref.model.l2norm(visual), ref.model.l2norm(text)), 1)      # model.py:208,302,483
        res = ref.evaluator.test_post_ranking(nb, "auc", model, posts, labels)
    dt_s = time.perf_counter() - t0
    return {"kind": "reference", "config": "configs[0]: %d brands x %d posts, D = %d, A = %d, gpu=-1" % (nb, n_posts, d, a),
            "value": nb * n_posts / dt_s, "unit": UNIT, "seconds": dt_s, "cores": n_threads,
            "metrics": {"MedR": float(res[0]), "AUC": float(res[2]), "NDCG@10": float(res[3])},
            "what": "unmodified evaluator.test_post_ranking (evaluator.py:85-143) + BrandAspects (model.py:406-428)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(CFG)
    t0 = time.perf_counter()
    steps, warmup = max(1, args.steps), max(args.warmup, 0)
    cpu = cpu_baseline(args.brands, args.cpu_sample_posts, cfg, steps=steps, warmup=warmup)
    cpu["c1_reference"] = reference_c1()
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": cpu["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.brands, args.posts_per_gpu, world, cfg),
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "each step = the reference's evaluation algorithm (oracle/ port, validated bit-exact against the reference "
                "in tests/) on a bounded sample of the configured workload: %s.  The unmodified reference itself "
                "(oracle/_ref) is timed on configs[0] in cpu_baseline.c1_reference; its per-brand Python loops are ~100x "
                "slower per pair than the vectorised port." % cpu["sample"],
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
