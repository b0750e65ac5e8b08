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
(seed = 20261018 + 1000*config + rank).  `value` has the inputs resident in HBM; `e2e` starts from pinned
HOST buffers every step (H2D inside the timed region) and ends with the metrics on the host.
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
    ap.add_argument("--cpu-sample-posts", type=int, default=20000)
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
class Evaluator:
    """One evaluation pass on this rank's shard (+ exchange when world > 1)."""

    def __init__(self, dev, world, rank, nb, n_local, cfg):
        from fancyrec_b200 import _lib, ops, ranking, sharded
        self.ops, self.ranking, self.sharded, self.lib = ops, ranking, sharded, _lib.load()
        self.dev, self.world, self.rank, self.nb, self.n_local, self.cfg = dev, world, rank, nb, n_local, cfg
        self.d = cfg["dv"] + cfg["dt"]
        self.n_total = n_local * world
        self.workspace = None
        self.launches = 0

    def step(self, w, e, labels, visual, text):
        ops = self.ops
        # brand side first (two small tensor-core launches), then the HBM-bound post finalisation.  Running the brand side
        # on a second stream was measured 0.5 ms/step SLOWER: its persistent GEMM cannot share SMs with the finalise blocks.
        brand = ops.brand_embed(w, e, nb=self.nb)                                      # split x2 + 3xTF32 GEMM
        brand_op = ops.finalize_posts(brand, final_norm=True)[1]                        # 1
        post_op = ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=True)[1]   # 1
        st = self.sharded.sharded_rank_statistics(brand_op, post_op, labels, self.d, self.cfg["k"], self.n_total,
                                                  workspace=self.workspace)
        self.workspace = st["workspace"]
        # our kernels per step (profiles/r01i_launches.csv): brand_embed = split_rows + split_transpose + 3xTF32 GEMM |
        # 2 x finalize | sample pass: dense score + row k-th select | main: fused score + merge | label_stats, decode_best,
        # rank_from_topk | missing_thresholds, score_count (returns at once unless a first positive is missing),
        # pack_rank_stats | sharded: merge of the gathered lists + reduce_shard_stats
        self.launches = 3 + 2 + 2 + 2 + 3 + 3 + (2 if self.world > 1 else 0)
        stats = self.ranking.host_statistics(st, self.n_total, want_auc=False)          # D2H of NB-length arrays
        return self.ranking.aggregate(stats, self.n_total, want_auc=False), st


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
    """Route fd 1 to stderr for the rest of the process (NCCL prints its version banner on stdout) and
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
        os.environ["NCCL_DEBUG"] = os.environ.get("FRX_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        torch.distributed.init_process_group("nccl", device_id=dev)
    cfg = dict(CFG)
    nb, n_local = args.brands, args.posts_per_gpu
    d = cfg["dv"] + cfg["dt"]
    w, e, labels, visual, text = make_workload(dev, rank, nb, n_local, cfg)
    ev = Evaluator(dev, world, rank, nb, n_local, cfg)
    pk = peaks()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                  # nvidia-smi needs ~100 ms to start reporting: start before warm-up
    for _ in range(max(args.warmup, 3)):
        result, st = ev.step(w, e, labels, visual, text)
    barrier_sync(world)

    # ---- timed region: K steps, device resident inputs (12.3 GB of fp32 inputs per step >> 126 MB L2)
    ev.lib.frx_probe_enable(1)
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    profile_range = bool(os.environ.get("FRX_PROFILE_RANGE"))   # ncu --profile-from-start off
    if profile_range:
        torch.cuda.profiler.start()
    t_start = time.perf_counter()
    beg.record()
    for _ in range(args.steps):
        result, st = ev.step(w, e, labels, visual, text)
    end.record()
    barrier_sync(world)
    t_end = time.perf_counter()
    if profile_range:
        torch.cuda.profiler.stop()
    ms_total = max_over_ranks(beg.elapsed_time(end), dev, world)
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    buf = (torch.zeros(4096, dtype=torch.float32)).numpy()
    n_probe = ev.lib.frx_probe_read(buf.ctypes.data, 4096)
    ev.lib.frx_probe_enable(0)
    # one probed launch per step: the main fused score + top-k kernel
    topk_ms = float(np.mean(buf[:n_probe])) if n_probe else float("nan")
    ms_step = ms_total / args.steps
    pairs = float(nb) * float(n_local) * world
    value = pairs / (ms_step * 1e-3)
    flops = 2.0 * nb * n_local * d
    achieved = flops / (topk_ms * 1e-3) / 1e12

    # ---- e2e: pinned host inputs -> H2D (chunked, overlapped with finalisation) -> metrics on host
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(ev, w, e, labels, visual, text, args, dev, world)
        except RuntimeError as ex:       # e.g. the host cannot pin 12.3 GB per rank
            e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.brands, args.cpu_sample_posts, cfg, steps=1)

    if rank == 0:
        traffic = None
        prof = os.path.join(ROOT, "profiles", "score_topk_traffic.json")
        if os.path.exists(prof):
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: %d brands x %d posts per GPU, 2048-d visual + 1024-d text fp32 "
                                   "embeddings, A=2000 aspects, score + top-100 + NDCG@10/50 (+recall@k, MedR)"
                                   % (nb, n_local),
                       "brands": nb, "posts_per_gpu": n_local, "posts_total": n_local * world, "dim": d, "k": cfg["k"],
                       "sharding": "posts x %d, one all-gather of top-k lists" % world,
                       "cache": "inputs larger than L2 (12.3 GB fp32 + 6.1 GB bf16 per step vs 126 MB)",
                       "seed": SEED0, "host_cores_bound_to_gpu_numa_node": numa_cores},
            "clocks": clocks,
            "gpu_launches": ev.launches * args.steps,
            "roofline": {"bound": "tensor", "kernel": "frx::score_kernel<MODE_TOPK> (tcgen05 bf16, fused top-k)",
                         "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tf_sustained"], "frac_of_burst_peak": achieved / pk["tf_burst"],
                         "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                         "kernel_ms": topk_ms, "kernel_share_of_step": topk_ms / ms_step,
                         "algorithmic_flops_per_launch": flops, "traffic": traffic},
            "metrics_sample": {"MedR": float(result[0]), "MeanR": float(result[1]), "NDCG@10": float(result[3]),
                               "NDCG@50": float(result[4]), "r1": result[5], "r5": result[6], "r10": result[7]},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_e2e(ev, w, e, labels, visual, text, args, dev, world):
    """Same step, but the post embeddings and labels start in pinned HOST memory every step."""
    ops = ev.ops
    n_local, dv, dt = ev.n_local, ev.cfg["dv"], ev.cfg["dt"]
    h_visual = torch.empty((n_local, dv), dtype=torch.float32, pin_memory=True)
    h_text = torch.empty((n_local, dt), dtype=torch.float32, pin_memory=True)
    h_labels = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    h_visual.copy_(visual); h_text.copy_(text); h_labels.copy_(labels)
    torch.cuda.synchronize()
    from fancyrec_b200 import ingest
    ld = ops.round_up(ev.d, 64)
    post_op = torch.empty((n_local, ld), dtype=torch.bfloat16, device=dev)
    d_labels = torch.empty(n_local, dtype=torch.int32, device=dev)

    def step():
        brand = ops.brand_embed(w, e, nb=ev.nb)
        brand_op = ops.finalize_posts(brand, final_norm=True)[1]
        d_labels.copy_(h_labels, non_blocking=True)
        # pinned host rows -> chunked H2D on a copy stream, overlapped with the finalisation of the previous chunk
        ingest.finalize_from_host(h_visual, h_text, visual_norm=True, text_norm=True, final_norm=True,
                                  out_bf16=post_op, device=dev, chunk_posts=131072)
        st = ev.sharded.sharded_rank_statistics(brand_op, post_op, d_labels, ev.d, ev.cfg["k"], ev.n_total,
                                                workspace=ev.workspace)
        stats = ev.ranking.host_statistics(st, ev.n_total, want_auc=False)
        return ev.ranking.aggregate(stats, ev.n_total, want_auc=False)

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
    pairs = float(ev.nb) * float(n_local) * world
    h2d = n_local * (dv + dt) * 4 + n_local * 4
    d2h = ev.nb * 5 * 8                      # the packed int64 [5, NB] statistics block
    return {"value": pairs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "api": "ops.brand_embed + ingest.finalize_from_host (pinned host -> device, chunked, overlapped) + "
                   "sharded.sharded_rank_statistics + ranking.aggregate; same call chain as "
                   "evaluator.test_post_ranking"}


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference algorithm on the host cores
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
            "sample": "%d brands x %d posts (%.1f%% of one GPU's posts), same dims / k / metrics; oracle/ NumPy "
                      "restatement of evaluator.py + torch.mm on %d host threads"
                      % (nb, sample_posts, 100.0 * sample_posts / CFG["np_per_gpu"], n_threads)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(CFG)
    t0 = time.perf_counter()
    cpu = cpu_baseline(args.brands, args.cpu_sample_posts, cfg, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world,
        "steps": max(1, args.steps), "warmup": min(args.warmup, 1), "ms_per_step": cpu["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1] sample: %s" % cpu["sample"], "brands": args.brands,
                   "posts_sample": args.cpu_sample_posts, "dim": cfg["dv"] + cfg["dt"], "k": cfg["k"]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Python and cannot travel to the GPU box; its algorithm is timed through the "
                "oracle/ port (validated bit-exact against the reference in tests/). The reference's own per-brand "
                "Python loops are ~100x slower than this vectorised port (SURVEY.md 6).",
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
