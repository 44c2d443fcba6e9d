#!/usr/bin/env python
"""bench.py -- GAN train step-pairs/s (one D step + one G step, mr_gan.py:204-213) on 1..8 B200.

Workload (BASELINE.json configs[2] slice, compact MREO shape): a group of `--folds`
independent table-1 fold-trainings per GPU (force+temperature modality, D=1200, N_train=6000,
N_test=1200, B=50), synthetic data, random-init weights.  One bench "step" = ONE EPOCH of the
group = 120 D+G step-pairs per fold x folds, launched as one CUDA graph (plus the batch-wise
test pass of mr_gan.py:219-223).  Folds shard over GPUs with no collective (weak scaling).

  value  : step-pairs/s, device time (CUDA events on the launching stream), fold data and the
           epoch's index arrays already resident in HBM.
  e2e    : the same metric through the public host API (load_fold + train_epoch with HOST
           numpy buffers + stats read-back + final eval), wall clock, H2D/D2H inside.
  --impl reference : the restated CPU baseline (oracle/torch_twin.py, torch CPU fp32, all host
           threads) -- Keras 2.0.9/Theano 0.9 cannot be installed here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gan_train_step_pairs_per_sec"
UNIT = "step-pairs/s"


def algo_work(D, B=50):
    """SURVEY.md 8(d) / BASELINE.md 3: algorithmic FLOPs and bytes per D+G step-pair."""
    P_D = 1000 * D + 751500
    P_G = 500 * D + 300000
    N_D = 1000 * D + 753756
    N_G = 501 * D + 302000
    flops = 2 * B * (9 * P_D + P_G - 3000 * D) + 2 * B * (3 * P_G + 3 * (P_D - 1500) - 50000)
    nbytes = 28 * N_D + 28 * N_G + 4 * (3 * B * D + 2 * B * 100)
    return flops, nbytes, N_D, N_G


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_jobs(D_modality, n_folds, seed):
    """Index-only folds of one synthetic MREO-shape dataset (table 1 reuses X for every labeled % and fold)."""
    from sklearn.model_selection import StratifiedKFold
    from mr_gan_b200 import foldprep, synthetic
    X, y = synthetic.synthetic_dataset(D_modality, seed=seed, dtype=np.float32)
    percents = [100, 50, 16, 8, 4, 2, 1]
    folds = []
    k = 0
    while len(folds) < n_folds:
        skf = StratifiedKFold(n_splits=6, shuffle=True, random_state=seed + k)
        for tr, te in skf.split(X, y):
            if len(folds) < n_folds:
                rng = np.random.default_rng([seed, len(folds)])
                folds.append((foldprep.prepare_fold_indices(y, tr, te, percents[k % len(percents)], None, rng), rng))
        k += 1
    return X, y.astype(np.int32), folds


def cpu_pairs_per_sec(D, B, n_pairs, warm=3, threads=None):
    """Restated CPU baseline: torch-CPU fp32 twin, Python loop, two calls per iteration, host noise."""
    import torch
    from oracle import gan_oracle as O, torch_twin as T
    # torchrun exports OMP_NUM_THREADS=1: use every host thread unless a thread count is asked for
    torch.set_num_threads(threads or max(1, os.cpu_count() or 1))
    rng = np.random.default_rng(0)
    m = T.TorchGan(O.init_disc_params(D, rng), O.init_gen_params(D, rng), dtype=torch.float32)
    X = rng.standard_normal((6000, D)).astype(np.float32)
    y = rng.integers(0, 6, 6000)
    def pair(t):
        sl = slice((t % 100) * B, (t % 100 + 1) * B)
        noise = np.random.normal(0, 1, size=[B, 100]).astype(np.float32)
        m.disc_step(X[sl], y[sl], X[sl], noise)
        noise = np.random.normal(0, 1, size=[B, 100]).astype(np.float32)
        m.gen_step(X[sl], noise)
    for t in range(warm):
        pair(t)
    t0 = time.perf_counter()
    for t in range(n_pairs):
        pair(t)
    dt = time.perf_counter() - t0
    return n_pairs / dt, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return 0
    D, B = args.width, 50
    pairs_per_step = args.ref_pairs
    cpu_pairs_per_sec(D, B, max(1, args.warmup) * pairs_per_step, warm=0)       # warm-up steps
    t0 = time.perf_counter()
    import torch
    from oracle import gan_oracle as O, torch_twin as T  # noqa: F401
    v, threads = cpu_pairs_per_sec(D, B, args.steps * pairs_per_step, warm=0)
    dt = time.perf_counter() - t0
    sample = "%d step-pairs per step of one fold (D=%d, B=%d), torch-CPU fp32 twin of the oracle" % (pairs_per_step, D, B)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * pairs_per_step / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "wall_ms": 1e3 * dt,
            "config": {"workload": "mr_gan table-1 fold, force+temperature D=%d, B=50 (CPU sample)" % D},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def run_dp(args, rank, world, local, barrier):
    """BASELINE.json config 5: one fold at a large global batch, batch rows split over the ranks, BN / feature-matching
    statistics and the flat gradient all-reduced with NCCL over NVLink (strong scaling: the global batch is fixed)."""
    import torch
    import torch.distributed as dist
    from mr_gan_b200.engine import FoldGroup
    from mr_gan_b200.model import init_disc, init_gen
    D, Bg = args.dp_width, args.dp_batch
    Bl, nb = Bg // world, 4
    W, K = max(args.warmup, 3), args.steps
    rng = np.random.default_rng(0)
    pD, pG = init_disc(D, rng), init_gen(D, rng)
    ntr = nb * Bl
    X = np.random.default_rng(100 + rank).standard_normal((ntr, D)).astype(np.float32)
    y = (np.arange(ntr) % 6).astype(np.int32)
    fg = FoldGroup([(D, ntr, 600, 4242)], precision=args.precision, batch=Bl, device=local, eval_each_epoch=False)
    if world > 1:
        uid = [FoldGroup.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        fg.dp_init(rank, world, uid[0])
    fg.set_params(0, 1, pG)
    fg.set_params(0, 0, pD)
    fg.load_fold(0, X, y, X[:600], y[:600])
    idx = np.arange(ntr, dtype=np.int32)[None, :]
    sampler = ClockSampler(local)
    for w in range(W):
        fg.train_epoch(idx, idx, idx)
    sampler.start()
    barrier()
    l0 = fg.kernel_launches
    t0 = time.perf_counter()
    dev_ms = 0.0
    for k in range(K):
        st = fg.train_epoch(idx, idx, idx)
        dev_ms += fg.last_device_ms
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = (float(x) for x in t.tolist())
    pairs = K * nb
    flops, nbytes, N_D, N_G = algo_work(D, Bg)
    hbm, tf, how = peaks()
    value = pairs / (dev_ms * 1e-3)
    ach = flops * value / 1e12
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "tf32",
            "data": "synthetic",
            "config": {"workload": "mr_gan.py table-5 fold at large batch, data-parallel: D=%d (1 s contact mic), global batch %d "
                                   "(%d per GPU); 1 step = %d D+G step-pairs" % (D, Bg, Bl, nb),
                       "parallelism": "dp%d, NCCL all-reduce of BN/FM statistics and flat gradients (%.0f MB D + %.0f MB G per pair)"
                                      % (world, 4e-6 * N_D, 4e-6 * N_G), "precision": args.precision},
            "e2e": {"value": pairs / (wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(3 * 4 * ntr),
                    "d2h_bytes_per_step": 32, "wall_ms": wall_ms},
            "gpu_launches": int(fg.kernel_launches - l0),
            "roofline": {"kernel": "whole step pair (tcgen05 GEMMs dominate at this batch)", "bound": "tensor", "achieved": ach,
                         "peak": tf * world, "unit": "TFLOP/s", "frac": ach / (tf * world), "traffic": None, "peak_source": how,
                         "note": "tf32 operands: the tf32 tensor peak is half the bf16 figure used as denominator"},
            "clocks": clocks, "sanity": {"last_loss_lab": float(st[0, 0]), "last_loss_gen": float(st[0, 3])}}
    fg.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--folds", type=int, default=74, help="fold-trainings grouped per GPU (table 1 has 294 = 4 x 73.5; 74 = 148/2 keeps every kernel at whole waves)")
    ap.add_argument("--modality", type=int, default=2, help="2 = force+temperature (D=1200)")
    ap.add_argument("--precision", default=os.environ.get("MRGAN_PRECISION", "tf32"), choices=["fp32", "tf32", "f16"])
    ap.add_argument("--ref-pairs", type=int, default=12)
    ap.add_argument("--cpu-pairs", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "dp"],
                    help="sweep: fold-sharded table-1 group (headline); dp: ONE fold at large batch, data-parallel (config 5)")
    ap.add_argument("--dp-batch", type=int, default=8192, help="global batch of the dp workload")
    ap.add_argument("--dp-width", type=int, default=12032, help="input width of the dp workload (1 s contact mic, table 5)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from mr_gan_b200 import synthetic
    args.width = synthetic.feature_width(args.modality)
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from mr_gan_b200 import foldprep
    from mr_gan_b200.engine import FoldGroup
    from mr_gan_b200.model import fold_key, init_disc, init_gen

    if args.workload == "dp":
        return run_dp(args, rank, world, local, barrier)

    D, B, G = args.width, 50, args.folds
    W, K = max(args.warmup, 3), args.steps
    X, y, folds = make_jobs(args.modality, G, seed=1000 * rank)
    ntr, nte = len(folds[0][0].train_rows), len(folds[0][0].test_rows)
    nb = ntr // B
    fg = FoldGroup([(D, ntr, nte, fold_key(rank, i)) for i in range(G)], precision=args.precision, device=local)

    def load_all():          # dataset upload (once) + device-side fold preparation of every fold (scaler, gather)
        fg.load_dataset(0, X, y)
        for i, (f, rng) in enumerate(folds):
            fg.prepare_fold(i, 0, f.train_rows, f.test_rows)

    def draw():
        per = [foldprep.epoch_indices(rng, ntr, f.lab_rows, f.unl_rows) for f, rng in folds]
        return [np.stack([p[s] for p in per]) for s in range(3)]

    for i, (f, rng) in enumerate(folds):
        fg.set_params(i, 1, init_gen(D, rng))
        fg.set_params(i, 0, init_disc(D, rng))
    load_all()
    pre = [draw() for _ in range(W + K)]

    sampler = ClockSampler(local)
    for w in range(W):
        fg.train_epoch(*pre[w])
    sampler.start()
    # ---- region 1: device-timed, inputs resident -------------------------------------------
    barrier()
    l0 = fg.kernel_launches
    t0 = time.perf_counter()
    dev_ms = 0.0
    for k in range(K):
        st = fg.train_epoch(*pre[W + k])
        dev_ms += fg.last_device_ms
    barrier()
    wall1 = time.perf_counter() - t0
    launches = fg.kernel_launches - l0
    # ---- region 2: end to end through the host API -----------------------------------------
    barrier()
    t0 = time.perf_counter()
    load_all()                                            # H2D of the dataset + per-fold index arrays, fold prep on the device
    t_load = time.perf_counter() - t0
    nxt = draw()
    for k in range(K):
        fg.train_epoch(*nxt, wait=False)
        if k + 1 < K:
            nxt = draw()                                  # host permutations overlap the GPU epoch
        st = fg.epoch_result()                            # D2H of the epoch statistics
    t_train = time.perf_counter() - t0 - t_load
    errs = [fg.eval(i) for i in range(G)]
    barrier()
    wall2 = time.perf_counter() - t0
    t_eval = wall2 - t_train - t_load
    clocks = sampler.stop()

    t = torch.tensor([dev_ms, wall1 * 1e3, wall2 * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall1_ms, wall2_ms = (float(x) for x in t.tolist())
    pairs = K * nb * G * world
    value = pairs / (dev_ms * 1e-3)
    e2e = pairs / (wall2_ms * 1e-3)

    # ---- live roofline probe of the dominant kernel (CUDA events on the launching stream) ----
    flops, nbytes, N_D, N_G = algo_work(D, B)
    hbm, tf, how = peaks()
    probe = {k: fg.time_op(k, reps=10) for k in ("adam_d", "dw1", "fwd1", "adam_g")}
    step_ms = {k: fg.time_op(k, reps=3) for k in ("disc_step", "gen_step")}
    dom = max(("adam_d", "dw1", "fwd1"), key=lambda k: probe[k])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")       # dram__bytes_read+write per launch from `ncu --set full`
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get("%s/%s" % (dom, args.precision), {})
        if tj.get("folds") == G and tj.get("D") == D:
            traffic = tj["bytes"]
    if dom == "adam_d" or (dom == "dw1" and args.precision == "tf32"):
        # HBM-bound: layer-1 dW with the Adam update fused in its epilogue (tf32) / the flat Adam kernel (fp32).
        # algorithmic bytes = W, m, v read + written once (24 B per parameter); gradients stay on chip
        n_par = (D + 1) * 1000 if dom == "dw1" else N_D
        ach = 24.0 * n_par * G / (probe[dom] * 1e-3) / 1e9
        name = "k_gemm_tc<dW + fused Adam> of D layer 1" if dom == "dw1" else "k_adam (flat, D net)"
        roof = {"kernel": name + ", all folds", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                "frac": ach / hbm, "traffic": traffic, "algorithmic_bytes_per_launch": 24.0 * n_par * G}
    else:
        fl = 2.0 * 3 * B * (D + 1) * 1000 * G
        ach = fl / (probe[dom] * 1e-3) / 1e12
        roof = {"kernel": ("dW" if dom == "dw1" else "forward") + " GEMM of D layer 1, all folds", "bound": "tensor",
                "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf, "traffic": traffic}
    roof["peak_source"] = how
    roof["launch_ms"] = probe[dom]
    step_gbs = nbytes * value / world / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "tf32", "data": "synthetic",
            "config": {"workload": "mr_gan.py table-1 fold group: %d folds/GPU, force+temperature D=%d, N_train=%d, "
                                   "N_test=%d, B=%d; 1 step = 1 epoch = %d D+G step-pairs per fold + test pass"
                                   % (G, D, ntr, nte, B, nb),
                       "folds_per_gpu": G, "D": D, "batch": B, "precision": args.precision,
                       "l2": "state of the group (%.0f MB) exceeds L2; no flush needed" % (12e-6 * (N_D + N_G) * G),
                       "parallelism": "fold-sharded x%d, no collective" % world},
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(3 * 4 * ntr * G + (X.nbytes + 4 * (ntr + nte) * G) / K),
                    "d2h_bytes_per_step": int(G * 8 * 4), "wall_ms": wall2_ms,
                    "breakdown_ms": {"load_and_prepare_folds": 1e3 * t_load, "epochs": 1e3 * t_train, "final_eval": 1e3 * t_eval},
                    "note": "includes the dataset upload and device-side fold preparation once, host permutations, final eval"},
            "gpu_launches": int(launches), "wall_ms_region1": wall1_ms,
            "fold_trainings_per_hour": value / (100 * nb) * 3600.0,
            "roofline": roof,
            "step_roofline": {"bound": "hbm", "achieved": step_gbs, "peak": hbm, "unit": "GB/s", "frac": step_gbs / hbm,
                              "algorithmic_bytes_per_pair": nbytes, "algorithmic_flops_per_pair": flops,
                              "achieved_tflops": flops * value / world / 1e12},
            "kernel_ms": probe, "step_ms": step_ms, "clocks": clocks,
            "sanity": {"final_test_err_mean": float(np.mean(errs)), "last_loss_lab": float(st[:, 0].mean())}}
    if rank == 0 and world == 1 and not args.no_cpu:
        v, threads = cpu_pairs_per_sec(D, B, args.cpu_pairs)
        v1, _ = cpu_pairs_per_sec(D, B, max(2, args.cpu_pairs // 6), warm=1, threads=1)     # 1 thread and all threads
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "host_cpus": os.cpu_count(),
                                "value_1thread": v1,
                                "sample": "%d D+G step-pairs of one fold (D=%d, B=%d), torch-CPU fp32 twin of the oracle "
                                          "(restated baseline, not Keras 2.0.9/Theano 0.9)" % (args.cpu_pairs, D, B)}
    fg.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
