#!/usr/bin/env python
"""Headline benchmark: encounters/s of (interpolation fwd+bwd + DEC assign) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1] ("c2"): 1,000,000 synthetic encounters x 6 vitals x <=256
irregular observations in 24 h, 96 reference points, per GPU (weak scaling: every rank owns
its own 1M-encounter shard), plus the DEC q/p assignment of one 256-d latent per encounter
(K = 4).  One step = one pass of the whole hot path over the shard:

    SCI fwd -> CCI fwd -> CCI+SCI bwd (one kernel) -> RBF fwd -> RBF bwd -> DEC q (+labels, column sum)
    -> all-reduce(column sum) -> DEC p -> all-reduce(parameter gradients)

`value`  : whole-job encounters/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : the same metric through the public nn.Module API with HOST (pinned) input buffers:
           H2D of every chunk of x, autograd fwd+bwd, D2H of loss/grads/labels in the timed region.
`roofline`: dominant kernel vs the measured HBM peak (schema of the task); `roofline_sfu` adds
           the MUFU.EX2 view, which is the resource that actually binds the interpolation kernels.
`cpu_baseline`: the reference's algorithm (oracle/ref_port.py, torch-CPU autograd) on a bounded
           sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

C, T, R, HOURS, D_LAT, K_CLUST = 6, 256, 96, 24.0, 256, 4
CFG_NAME = "c2: interp fwd+bwd + DEC assign, 1M enc x 6 vitals x <=256 obs, 96 ref points"
# BASELINE.json configs; the headline metric is quoted on c2 (the default).  c1 is the reference's
# CPU-runnable case, c5 the stress shape (its 10M encounters are processed in HBM-sized shards: the
# default shard here is 131,072 encounters = 12.9 GB of x per GPU per step).
WORKLOADS = {
    "c1": dict(T=64, R=48, K=4, B=1000, name="c1: interp fwd+bwd + DEC assign, 1,000 enc x 6 vitals x <=64 obs, 48 ref points"),
    "c2": dict(T=256, R=96, K=4, B=1_000_000, name=CFG_NAME),
    # c3 = c2 with the whole DEC step of the joint-clustering loop: q (+labels, column sum), then ONE fused kernel
    # for p, the KL sum and the closed-form gradients dz, dmu (clustering_interp.py:186,205-207)
    "c3": dict(T=256, R=96, K=4, B=1_000_000, dec_kl=True,
               name="c3: interp fwd+bwd + DEC q/p/KL fwd+bwd, K=4, 1M enc x 6 vitals x <=256 obs, 96 ref points"),
    # c4 = the K-selection sweep of p2 AS WRITTEN (p2_clustering_optK.py:33,36,37,353-410): K = 2..10, 20 reference draws,
    # n_init = 10 on a 1M x 64 latent matrix; one step = one whole sweep (c4_arm below; not an interp workload)
    "c4": dict(T=256, R=96, K=10, B=1_000_000, sweep=True,
               name="c4: gap-statistic k-means sweep K=2..10, 20 reference draws, n_init=10, 1M x 64-d latents (p2 path)"),
    # p1 = one training step of the WHOLE pretrain network: the unmodified pretrain_interp.Net of the staged
    # reference on top of the B200 mirrors - SCI -> CCI -> BiLSTM encoder -> BiLSTM decoder -> compress_fc -> RBF read-out,
    # masked reconstruction MSE, backward through everything, Adam step (p1_arm below)
    "p1": dict(T=256, R=96, K=4, B=16_384, net=True,
               name="p1: full pretrain Net step (interp + BiLSTM encoder/decoder + compress_fc + RBF read-out, fwd+bwd+Adam), "
                    "6 vitals x <=256 obs, 96 ref points, 16,384 encounters per GPU per step"),
    "c5": dict(T=1024, R=192, K=16, B=131_072,
               name="c5 (stress shape): interp fwd+bwd + DEC assign, 6 vitals x <=1024 obs, 192 ref points, K=16, "
                    "one 131,072-encounter shard of the 10M per step"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--encounters", type=int, default=0, help="encounters per GPU per step (0 = the workload's)")
    ap.add_argument("--e2e-encounters", type=int, default=262_144)
    ap.add_argument("--e2e-chunk", type=int, default=32_768)
    ap.add_argument("--e2e-upload", default="packed", choices=["packed", "dense"],
                    help="host format of x in the e2e arm: ragged/packed rows (default) or the dense planes")
    ap.add_argument("--cpu-sample", type=int, default=256, help="encounters in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c4-dim", type=int, default=64)
    ap.add_argument("--c4-kmax", type=int, default=10)
    ap.add_argument("--c4-refs", type=int, default=20)
    ap.add_argument("--c4-ninit", type=int, default=10)
    ap.add_argument("--c4-draw", default="device", choices=["device", "device32", "host"],
                    help="reference sets: float64 uniform draws generated on the GPU (default; the reference draws float64 "
                         "on the host, :370), float32 device draws, or numpy host draws uploaded per set")
    ap.add_argument("--c4-cpu-n", type=int, default=1500, help="rows of the CPU-baseline sweep (n x n matrices)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short c3 / c4 / c5 runs printed as `other_configs` next to the c2 line")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU result checks printed as `parity` (N > 1)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU arm (reference algorithm on host cores)
# ----------------------------------------------------------------------------------------------
def cpu_arm(sample, steps, warmup):
    """Times the reference's CPU implementation of the step on `sample` encounters of the workload's shape + the DEC
    step on the same number of latents.  With the reference's own modules staged under baseline/_ref
    (oracle/make_ref.py; built by __graft_entry__.build() where /root/reference exists) these ARE the unmodified
    upstream classes (kind "reference"); otherwise oracle/ref_port (the same ATen-level algorithm restated, kind
    "port").  Returns (encounters/s, seconds per step, threads, kind)."""
    import numpy as np
    import torch
    from deep_interpolation_clustering_b200 import synth
    from oracle import make_ref, ref_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = torch.from_numpy(synth.make_encounters(sample, C, T, HOURS, seed=0))
    p = synth.make_interp_params(C, seed=1)
    rng = np.random.RandomState(2)
    v = torch.from_numpy(rng.normal(size=(sample, C, R)).astype(np.float32)).requires_grad_(True)
    g = torch.from_numpy(rng.normal(size=(sample, R, 3 * C)).astype(np.float32))
    zn, mun = synth.make_latents(sample, D_LAT, K_CLUST, seed=3)
    z = torch.from_numpy(zn).requires_grad_(True)
    ref = make_ref.import_reference()
    if ref is not None:
        kind, cpu = "reference", torch.device("cpu")
        sci = ref.interpolation_layer.SingleChannelInterp(R, HOURS, C, T, cpu)          # interpolation_layer.py:14
        cci = ref.interpolation_layer.CrossChannelInterp(C, T, cpu)                     # :90
        rbf = ref.rbf.RBF(HOURS, R, C, C, 0.0, ref.rbf.basis_func_dict()["gaussian"], cpu)   # rbf.py:38
        rbf.compress_fc = torch.nn.Identity()      # the metric's step has no encoder: v is the read-out's input
        ca = ref.dec.ClusterAssignment(K_CLUST, D_LAT, 1.0, torch.from_numpy(mun))      # dec.py:14
        mu = ca.cluster_centers

        def step():
            for t in (sci.kernel, cci.kernel, rbf.kernel, v, z, mu):
                t.grad = None
            out = cci(sci(x))
            rec = rbf(v, x)
            (((out * g).sum() / x.shape[0]) + ref_port.masked_mse(x, rec, C)).backward()
            q = ca(z)
            pt = ref.dec.target_distribution(q).detach()
            torch.nn.functional.kl_div(q.log(), pt, reduction="batchmean").backward()   # clustering_interp.py:205-207
    else:
        kind = "port"
        mu = torch.from_numpy(mun).requires_grad_(True)
        sci, cci, rbf = ref_port.SingleChannelInterp(R, HOURS, C), ref_port.CrossChannelInterp(C), \
            ref_port.RBFReadout(R, HOURS, C)

        def step():
            for t in (sci.kernel, cci.kernel, rbf.kernel, v, z, mu):
                t.grad = None
            ref_port.interp_step(sci, cci, rbf, x, v, g)
            ref_port.dec_step(z, mu, 1.0)
    sci.kernel.data = torch.from_numpy(p["sci_kernel"])
    cci.kernel.data = torch.from_numpy(p["cci_kernel"])
    rbf.kernel.data = torch.from_numpy(p["rbf_kernel"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, dt, cores, kind


# ----------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# device arm
# ----------------------------------------------------------------------------------------------
class HotPath:
    """Preallocated buffers + direct C-ABI calls for one shard (no allocation in the timed loop)."""

    # launches of OUR kernels behind each C-ABI call (kernel + reductions)
    LAUNCHES = dict(sci_fwd=1, cci_fwd=1, cci_sci_bwd=6, rbf_fwd=1, rbf_bwd=4, dec_q=2, dec_p=1, dec_kl=2)

    def __init__(self, B, dev, seed, dec_kl=False):
        self.dec_kl = dec_kl
        import torch
        from deep_interpolation_clustering_b200 import _lib, synth
        self.t, self.L, self._lib, self.B, self.dev = torch, _lib.lib(), _lib, B, dev
        f32 = dict(dtype=torch.float32, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(seed + 100)
        self.x = synth.make_encounters_device(B, C, T, HOURS, 5.0, seed, dev)
        p = synth.make_interp_params(C, seed=1)
        self.k_sci = torch.tensor(p["sci_kernel"], **f32)
        self.k_cci = torch.tensor(p["cci_kernel"], **f32)
        self.k_rbf = torch.tensor(p["rbf_kernel"], **f32)
        self.ref_t = torch.linspace(0, HOURS, R).to(dev)
        self.u = torch.empty((B, 3 * C, R), **f32)
        self.stats = torch.empty((B, 3 * C, R), **f32)          # gradient-moment rows [U1 | U0 | U1']
        self.out = torch.empty((B, 3 * C, R), **f32)
        self.g_out = torch.randn((B, 3 * C, R), generator=g, **f32)
        self.v = torch.randn((B, C, R), generator=g, **f32)
        self.rec = torch.empty((B, C, T), **f32)
        self.inv = torch.empty((B, C, T), **f32)
        self.g_rec = torch.randn((B, C, T), generator=g, **f32)
        self.g_v = torch.empty((B, C, R), **f32)
        self.z, self.mu = synth.make_latents_device(B, D_LAT, K_CLUST, seed + 7, dev)
        self.q = torch.empty((B, K_CLUST), **f32)
        self.p = torch.empty((B, K_CLUST), **f32)
        self.labels = torch.empty(B, dtype=torch.int32, device=dev)
        self.colsum = torch.empty(K_CLUST, dtype=torch.float64, device=dev)
        self.grads = torch.empty(C + C * C + C, **f32)          # packed [d sci.kernel | d cci.kernel | d rbf.kernel]
        self.ws_i = torch.empty(int(self.L.dic_interp_bwd_workspace_bytes(B, C)), dtype=torch.uint8, device=dev)
        self.ws_c = torch.empty(int(self.L.dic_cci_sci_bwd_workspace_bytes(B, C)), dtype=torch.uint8, device=dev)
        self.ws_d = torch.empty(int(self.L.dic_dec_workspace_bytes(K_CLUST, D_LAT)), dtype=torch.uint8, device=dev)
        if dec_kl:
            self.g_z = torch.empty_like(self.z)
            self.g_mu = torch.empty_like(self.mu)
            self.kl = torch.empty(1, dtype=torch.float64, device=dev)
        self.n_valid = float(self.x[:, C:2 * C].sum())

    def kernels(self, st):
        """[(name, callable)] in step order; each callable issues one C-ABI call on stream st."""
        L, P, B = self.L, self._lib.ptr, self.B
        chk = self._lib.check
        gs, gc, gr = self.grads[:C], self.grads[C:C + C * C], self.grads[C + C * C:]
        tail = [("dec_p", lambda: chk(L.dic_dec_p(P(self.q), P(self.colsum), P(self.p), B, K_CLUST, st), "dec_p"))]
        if self.dec_kl:      # p, KL sum and the gradients dz, dmu in one kernel; scale = weight 10 / B (p3_clustering_main.py:85)
            tail = [("dec_kl", lambda: chk(L.dic_dec_kl_fwd_bwd(P(self.z), P(self.mu), P(self.colsum), P(self.p), P(self.kl),
                                                                 P(self.g_z), P(self.g_mu), P(self.ws_d), B, D_LAT, K_CLUST,
                                                                 1.0, 10.0 / B, st), "dec_kl"))]
        return [
            ("sci_fwd", lambda: chk(L.dic_sci_fwd(P(self.x), P(self.k_sci), P(self.ref_t), P(self.u), P(self.stats),
                                                  B, C, T, R, 0, st), "sci_fwd")),
            ("cci_fwd", lambda: chk(L.dic_cci_fwd(P(self.u), P(self.k_cci), P(self.out), B, C, R, st), "cci_fwd")),
            # CCI backward with the SCI backward folded in: the gradient of the SCI output never reaches HBM
            ("cci_sci_bwd", lambda: chk(L.dic_cci_sci_bwd(P(self.u), P(self.k_cci), P(self.k_sci), P(self.stats),
                                                          P(self.g_out), P(gc), P(gs), P(self.ws_c), B, C, R, st),
                                        "cci_sci_bwd")),
            ("rbf_fwd", lambda: chk(L.dic_rbf_fwd(P(self.v), P(self.x), P(self.k_rbf), P(self.ref_t), P(self.rec),
                                                  P(self.inv), B, C, T, R, 0, st), "rbf_fwd")),
            ("rbf_bwd", lambda: chk(L.dic_rbf_bwd(P(self.v), P(self.x), P(self.k_rbf), P(self.ref_t), P(self.rec),
                                                  P(self.inv), P(self.g_rec), P(self.g_v), P(gr), P(self.ws_i),
                                                  B, C, T, R, 0, st), "rbf_bwd")),
            ("dec_q", lambda: chk(L.dic_dec_q_fwd(P(self.z), P(self.mu), P(self.q), P(self.labels), P(self.colsum),
                                                  P(self.ws_d), B, D_LAT, K_CLUST, 1.0, st), "dec_q")),
        ] + tail

    def algorithmic(self):
        """Algorithmic HBM bytes and MUFU exps per launch (split per kernel)."""
        B, nv = self.B, self.n_valid
        ct, cr = C * T * 4.0, C * R * 4.0
        byt = dict(sci_fwd=3 * ct + 3 * cr + 3 * cr, cci_fwd=6 * cr, cci_sci_bwd=9 * cr,      # u, grad_out, stats in
                   rbf_fwd=2 * ct + cr + 2 * ct,
                   rbf_bwd=2 * ct + 3 * ct + cr + cr, dec_q=D_LAT * 4.0 + K_CLUST * 4.0 + 4.0, dec_p=2 * K_CLUST * 4.0,
                   dec_kl=2 * D_LAT * 4.0 + K_CLUST * 4.0)
        ex2 = dict(sci_fwd=2 * nv * R, rbf_fwd=nv * R, rbf_bwd=nv * R)     # sci_bwd no longer sweeps the observations
        return {k: v * B for k, v in byt.items()}, ex2


def make_step(hp, kernels, world):
    """One pass of the hot path over the shard `hp` (the order of the module docstring), with the path's own
    exchange steps when the encounters are sharded over `world` ranks."""
    import torch.distributed as dist

    def step():
        for name, fn in kernels:
            if name in ("dec_p", "dec_kl") and world > 1:
                dist.all_reduce(hp.colsum)               # f_j over the global batch (dec.py:73)
            fn()
        if world > 1:
            dist.all_reduce(hp.grads)                    # parameter gradients of the sharded batch
            if hp.dec_kl:
                dist.all_reduce(hp.g_mu)                 # centre gradients (K x D floats)
    return step


def other_configs(args, rank, world, dev):
    """The other BASELINE.json configurations next to the headline line, each as a SHORT device-timed run at this world
    size so that every configuration has a driver-visible number (the full-size runs are `--workload c3|c4|c5`):
      c3  the joint-clustering step (c2 + the fused DEC p / KL / gradient kernel) on 262,144 encounters per GPU;
      c5  the stress shape (T = 1024, R = 192, K = 16) on one 32,768-encounter shard per GPU (10M encounters are
          processed shard by shard: generation on the device, nothing larger ever resident);
      c4  the gap-statistic sweep on the full 1M x 64 matrix, K = 2..10, REDUCED to 2 reference draws and n_init = 2
          (the sweep as written - 20 draws, n_init = 10 - is `--workload c4`: profiles/r02_c4_*.json)."""
    global T, R, K_CLUST
    import torch
    import torch.distributed as dist
    saved = (T, R, K_CLUST)
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_hot_path(B, dec_kl, steps=3):
        hp = HotPath(B, dev, seed=1000 * rank + 17, dec_kl=dec_kl)
        stream = torch.cuda.current_stream(dev)
        step = make_step(hp, hp.kernels(stream.cuda_stream), world)
        for _ in range(3):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nv = hp.n_valid / (B * C)
        del hp, step
        torch.cuda.empty_cache()
        return float(t), nv

    try:
        for tag in ("c3", "c5"):
            w = WORKLOADS[tag]
            T, R, K_CLUST = w["T"], w["R"], w["K"]
            B = 262_144 if tag == "c3" else 32_768
            ms, nv = timed_hot_path(B, bool(w.get("dec_kl")))
            out[tag] = {"value": round(world * B / (ms * 1e-3), 1), "unit": "encounters/s", "ms_per_step": round(ms, 3),
                        "encounters_per_gpu_per_step": B, "n_gpus": world, "scaling": "weak", "steps": 3, "warmup": 3,
                        "config": {"workload": w["name"], "max_obs": T, "ref_points": R, "clusters": K_CLUST,
                                   "mean_valid_obs": round(nv, 2)}}
    finally:
        T, R, K_CLUST = saved
    # c4 (reduced draws / restarts, full matrix), task-parallel over the ranks
    from deep_interpolation_clustering_b200.gap import KM
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    N, D = 1_000_000, 64
    g = torch.Generator(device=dev).manual_seed(4)
    centres = 4.0 * torch.randn((5, D), generator=g, device=dev)
    X = (centres[torch.randint(0, 5, (N,), generator=g, device=dev)] + torch.randn((N, D), generator=g, device=dev)).contiguous()
    group = dist.group.WORLD if world > 1 else None

    def sweep(k_max, refs, n_init):
        df = KM(k_max, None, [], n_init, refs).compute_gap_internal_metric(
            KMeansB200(n_init=n_init, random_state=0), X, k_max=k_max, n_references=refs, version=1, draw="device",
            group=group, task_parallel=True, seed=0)
        torch.cuda.synchronize(dev)
        return df

    for _ in range(2):
        sweep(3, 1, 1)
    barrier()
    t0 = time.perf_counter()
    df = sweep(10, 2, 2)
    barrier()
    wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    out["c4_reduced"] = {"value": round(N / float(wall), 1), "unit": "rows/s", "seconds_per_sweep": round(float(wall), 3),
                         "n_gpus": world, "scaling": "strong", "steps": 1, "warmup": 2,
                         "config": {"workload": "c4 REDUCED: gap-statistic k-means sweep K=2..10 on 1M x 64-d latents with 2 "
                                                "reference draws and n_init=2 (as written: 20 draws, n_init=10 = "
                                                "`bench.py --workload c4`)",
                                    "rows": N, "dim": D, "k": "2..10", "n_references": 2, "n_init": 2,
                                    "kmeans_fits_per_sweep": 9 * 3 * 2, "pairwise_evaluations_per_sweep": 27,
                                    "parallelism": f"task-parallel x{world}"},
                         "best_k": int(df["gap"].astype(float).idxmax())}
    # the kernel the sweep spends its time in, live: one Lloyd pass (K = 10, loop form) over the data matrix (float32) and
    # over a reference-set-sized float64 matrix, against the measured HBM figure (matrices larger than L2)
    try:
        from deep_interpolation_clustering_b200.kmeans import _Device
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        stream = torch.cuda.current_stream(dev)
        passes = {}
        for name, Xk in (("data_f32", X), ("reference_set_f64", X.to(torch.float64))):
            st = _Device(Xk, 10)
            cen = Xk[:10].clone().contiguous()
            for _ in range(3):
                st.assign(cen, 1 | 4)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(20):
                st.assign(cen, 1 | 4)
            b.record(stream)
            b.synchronize()
            ms = a.elapsed_time(b) / 20
            nbytes = N * D * Xk.element_size() + N * 4
            passes[name] = {"ms": round(ms, 4), "gbps": round(nbytes / 1e9 / (ms * 1e-3), 1),
                            "hbm_frac": round(nbytes / 1e9 / (ms * 1e-3) / hbm_peak, 4)}
            del st, Xk
        out["c4_reduced"]["lloyd_pass_K10"] = passes
        out["c4_reduced"]["lloyd_pass_note"] = ("dic_kmeans_assign (COUNT_CHANGES | NO_INERTIA) incl. its partial-sum pass, 20 "
                                                "launches, CUDA events on the launching stream; algorithmic bytes = one read "
                                                "of the matrix + the labels; peak = " +
                                                ("MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s"))
    except Exception as e:                      # a reporting extra: never the reason a bench line is lost
        out["c4_reduced"]["lloyd_pass_K10"] = {"error": str(e)[:200]}
    del X
    torch.cuda.empty_cache()
    return out


def device_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from deep_interpolation_clustering_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.encounters
    hp = HotPath(B, dev, seed=1000 * rank, dec_kl=bool(WORKLOADS[args.workload].get("dec_kl")))
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream
    kernels = hp.kernels(st)
    step = make_step(hp, kernels, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms) / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # launches of OUR kernels in one step, COUNTED (CUPTI through torch.profiler, names in namespace dic::) on one
    # extra step outside the timed region; the static per-call table is only the fallback when CUPTI is unavailable
    launches_per_step, launches_how = sum(HotPath.LAUNCHES[n] for n, _ in kernels), "table (CUPTI unavailable)"
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize(dev)
        counted = sum(e.count for e in prof.key_averages() if "dic::" in e.key)
        if counted > 0:
            launches_per_step, launches_how = int(counted), "counted with CUPTI (kernels named dic::*) on one step"
    except Exception:       # noqa: BLE001
        pass

    # per-kernel durations (CUDA events on the launching stream), for the roofline objects
    per = {n: [] for n, _ in kernels}
    for _ in range(3):
        for name, fn in kernels:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            per[name].append(a.elapsed_time(b))
    kms = {n: statistics.mean(v) for n, v in per.items()}

    e2e = None
    if not args.no_e2e:
        prev_affinity = bind_to_gpu_numa(local_rank)
        e2e = e2e_arm(args, hp, dev, world)
        if e2e is not None:
            e2e["host_numa_binding"] = prev_affinity is not None
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)        # the CPU baseline below uses all host cores again

    # result checks at this world size (sharded gradients, DEC p/KL, k-means, gap sweeps vs the single-GPU truth), so
    # that a scaling record carries parity and not only speed (deep_interpolation_clustering_b200/dist_check.py)
    parity = None
    if world > 1 and not args.no_parity:
        from deep_interpolation_clustering_b200 import dist_check
        try:
            rep = dist_check.run(rank, world, dev)
            parity = {"world": world, "ok": True, **{k: float(f"{v:.3e}") for k, v in rep.items()}}
        except AssertionError as e:          # noqa: PERF203
            parity = {"world": world, "ok": False, "error": str(e)[:300]}
        ok = torch.tensor([1.0 if parity["ok"] else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity["all_ranks_ok"] = bool(ok.item() == 1.0)

    byt, ex2 = hp.algorithmic()
    n_valid_mean = hp.n_valid / (B * C)
    others = None
    if args.workload == "c2" and not args.no_other_configs:
        del step, kernels, hp
        torch.cuda.empty_cache()
        try:
            others = other_configs(args, rank, world, dev)
        except Exception as e:       # noqa: BLE001 - the headline line must not be lost to a side run
            others = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    mufu = ctypes.c_double()
    _lib.check(_lib.lib().dic_probe_mufu(ctypes.byref(mufu), st), "probe_mufu")
    ffma = ctypes.c_double()
    _lib.check(_lib.lib().dic_probe_ffma(ctypes.byref(ffma), st), "probe_ffma")
    ktab = {}
    for n, m in kms.items():
        ktab[n] = {"ms": round(m, 4), "share": round(m / sum(kms.values()), 4),
                   "algorithmic_gb": round(byt[n] / 1e9, 3), "gbps": round(byt[n] / 1e9 / (m * 1e-3), 1),
                   "hbm_frac": round(byt[n] / 1e9 / (m * 1e-3) / hbm_peak, 4)}
        if n in ex2:
            ktab[n]["gex2_per_s"] = round(ex2[n] / 1e9 / (m * 1e-3), 1)
            ktab[n]["mufu_frac"] = round(ex2[n] / (m * 1e-3) / mufu.value, 4)
    dom = max(kms, key=kms.get)
    # DRAM traffic, executed MUFU count and pipe utilisation come from ONE ncu --set full capture of this step
    # (benchmarks/capture_traffic.py -> profiles/dram_traffic.json), tied to the SHA-256 of the kernel sources it was
    # measured on: if the tree has changed since, the numbers are refused (null + a note), never printed stale.
    traffic, ncu_k, traffic_note = None, None, "no capture (profiles/dram_traffic.json missing)"
    try:
        from deep_interpolation_clustering_b200.build import sources_sha256
        tr = json.load(open(os.path.join(REPO, "profiles", "dram_traffic.json")))
        if tr.get("sources_sha256") != sources_sha256():
            traffic_note = "refused: profiles/dram_traffic.json was captured on different kernel sources (stale)"
        elif args.workload not in ("c2", "c3"):
            traffic_note = "capture is of the c2 shape"
        elif dom in tr["kernels"]:
            ncu_k = tr["kernels"][dom]
            traffic = round(ncu_k["bytes_per_launch"] * (B / tr["encounters"]) / 1e9, 3)
            traffic_note = f"ncu dram__bytes_read+write of one launch at {tr['encounters']} encounters, scaled to {B}"
    except Exception as e:       # noqa: BLE001
        traffic_note = f"no capture ({type(e).__name__})"
    roofline = {"kernel": dom, "bound": "hbm", "achieved": ktab[dom]["gbps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": ktab[dom]["hbm_frac"], "traffic": traffic, "traffic_unit": "GB per launch", "traffic_note": traffic_note,
                "peak_source": peak_src,
                "note": "the interpolation sweeps are MUFU / issue bound, not HBM bound: see roofline_sfu and roofline_issue"}
    # SFU view on EXECUTED exponentials: the kernels skip pairs whose weight is below 2^-27..2^-30 of the largest and
    # evaluate the high-pass exponential only inside its narrow window, so the algorithmic count (2 per pair for SCI,
    # 1 for RBF) is not what runs; executed = ncu sm__inst_executed_pipe_xu (warp instructions) x 32 lanes.
    exec_ex2 = None
    if ncu_k and ncu_k.get("xu_warp_inst"):
        exec_ex2 = ncu_k["xu_warp_inst"] * 32.0 * (B / tr["encounters"])
    elif ncu_k and ncu_k.get("xu_pipe_pct") and ncu_k.get("ncu_duration_us"):
        # no absolute count in the export: XU-pipe utilisation of the capture (executed MUFU instructions per cycle over
        # the pipe's peak) x the capture's duration = executed lane-operations at the probe's peak rate
        ncu_ms = ncu_k["ncu_duration_us"] * (1.0 if ncu_k["ncu_duration_us"] < 1e3 else 1e-3)     # export unit: ms
        exec_ex2 = ncu_k["xu_pipe_pct"] / 100.0 * mufu.value * (ncu_ms * 1e-3) * (B / tr["encounters"])
    roofline_sfu = {"kernel": dom, "bound": "sfu", "unit": "Gex2/s",
                    "achieved": round(exec_ex2 / 1e9 / (kms[dom] * 1e-3), 1) if exec_ex2 else None,
                    "peak": round(mufu.value / 1e9, 1),
                    "frac": round(exec_ex2 / (kms[dom] * 1e-3) / mufu.value, 4) if exec_ex2 else None,
                    "algorithmic_gex2_per_s": ktab[dom].get("gex2_per_s"),
                    "algorithmic_over_peak": ktab[dom].get("mufu_frac"),
                    "ffma_peak_gops": round(ffma.value / 1e9, 1),
                    "peak_source": "dic_probe_mufu: ex2.approx-only kernel timed in this run",
                    "note": "frac = EXECUTED MUFU lane-operations per second (ncu XU-pipe instruction count of the capture x 32 / "
                            "this run's kernel time) over the probe's rate; algorithmic_over_peak counts every (observation, "
                            "grid point) exponential of the reference and may exceed 1 because skipped pairs are not executed"}
    roofline_issue = {"kernel": dom, "bound": "issue", "unit": "% of issue slots",
                      "achieved": ncu_k.get("issue_active_pct") if ncu_k else None, "peak": 100.0,
                      "frac": round(ncu_k["issue_active_pct"] / 100.0, 4) if ncu_k and ncu_k.get("issue_active_pct") else None,
                      "xu_pipe_pct": ncu_k.get("xu_pipe_pct") if ncu_k else None,
                      "fma_pipe_pct": ncu_k.get("fma_pipe_pct") if ncu_k else None,
                      "shared_bank_conflict_share": round(ncu_k["shared_bank_conflict_share"], 4) if ncu_k else None,
                      "source": traffic_note}

    line = {
        "metric": "encounters/s (interp fwd+bwd + DEC assign)", "value": round(value, 1), "unit": "encounters/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CFG_NAME, "encounters_per_gpu": B, "vitals": C, "max_obs": T, "ref_points": R,
                   "hours": HOURS, "latent_dim": D_LAT, "clusters": K_CLUST, "mean_valid_obs": round(n_valid_mean, 2),
                   "parallelism": f"encounter-sharded x{world}",
                   "l2": f"inputs ({B * 4 * C * T * 4 / 1e9:.1f} GB/GPU) " + ("exceed L2" if B * 4 * C * T * 4 > 2.5e8 else
                                                                           "fit L2: flushed by the other kernels' buffers")},
        "clocks": clocks, "gpu_launches": args.steps * launches_per_step, "gpu_launches_per_step": launches_per_step,
        "gpu_launches_how": launches_how,
        "kernels": ktab, "roofline": roofline, "roofline_sfu": roofline_sfu, "roofline_issue": roofline_issue, "e2e": e2e,
    }
    if parity is not None:
        line["parity"] = parity
    if others is not None:
        line["other_configs"] = others
    if not args.no_cpu_baseline:
        v, dt, cores, kind = cpu_arm(args.cpu_sample, 2, 1)
        how = "the reference's own modules staged in baseline/_ref (oracle/make_ref.py)" if kind == "reference" \
            else "oracle/ref_port.py (the reference's algorithm restated; no staged reference on this box)"
        line["cpu_baseline"] = {"value": round(v, 1), "unit": "encounters/s", "cores": cores, "kind": kind,
                                "sample": f"{args.cpu_sample} encounters of this workload's shape + {args.cpu_sample} "
                                          f"latents, torch CPU autograd through {how}, {dt:.2f} s/step"}
    if world > 1:
        dist.destroy_process_group()
    return line


# ----------------------------------------------------------------------------------------------
# c4: the K-selection sweep of p2 (gap statistic), task-parallel over the ranks
# ----------------------------------------------------------------------------------------------
def c4_cpu_arm(n_rows, dim, k_max, refs, n_init):
    """The reference's sweep on a matrix small enough for its n_c x n_c distance matrices: the staged reference's own
    KM.compute_gap_internal_metric (p2_clustering_optK.py:353-410, kind "reference") over scikit-learn's KMeans exactly
    like :284, or - without a staged reference - oracle/kmeans_oracle.gap_statistic, the same routine restated (kind
    "port").  Returns (rows/s, seconds, cores, kind)."""
    import numpy as np
    from sklearn.cluster import KMeans
    from deep_interpolation_clustering_b200 import synth
    from oracle import kmeans_oracle, make_ref
    X = synth.make_blobs(n_rows, dim, 5, seed=4)
    np.random.seed(123)
    ref = make_ref.import_reference(p2=True)
    t0 = time.perf_counter()
    if ref is not None:
        kind = "reference"
        km = ref.p2.KM(k_max, "/tmp/dic_bench_p2", [], n_init, refs)
        km.compute_gap_internal_metric(KMeans(n_init=n_init), X, k_max=k_max, n_references=refs, version=1)
    else:
        kind = "port"
        kmeans_oracle.gap_statistic(lambda k, data: KMeans(n_clusters=k, n_init=n_init).fit_predict(data), X,
                                    k_max=k_max, n_references=refs, version=1)
    dt = time.perf_counter() - t0
    return n_rows / dt, dt, os.cpu_count() or 1, kind


def c4_arm(args, rank, world, local_rank):
    """One step = one gap-statistic sweep as p2 runs it (K = 2..k_max, `refs` reference draws, n_init restarts, version
    1 inertia) through KM.compute_gap_internal_metric + KMeansB200.  Every rank holds the whole matrix (256 MB) and takes
    the (k, reference set) fits round-robin; ONE all-reduce of the inertia table ends the sweep."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from deep_interpolation_clustering_b200 import _lib
    from deep_interpolation_clustering_b200.gap import KM, pairwise_dist_sum
    from deep_interpolation_clustering_b200.kmeans import KMeansB200, _Device
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, D = args.encounters, args.c4_dim
    g = torch.Generator(device=dev).manual_seed(4)                   # the same matrix on every rank
    centres = 4.0 * torch.randn((5, D), generator=g, device=dev)
    lab = torch.randint(0, 5, (N,), generator=g, device=dev)
    X = (centres[lab] + torch.randn((N, D), generator=g, device=dev)).contiguous()     # blobs, sigma 1
    draw = None if args.c4_draw == "host" else args.c4_draw
    group = dist.group.WORLD if world > 1 else None

    def sweep(data, k_max, refs, n_init, timers=None):
        km = KM(k_max, None, [], n_init, refs)
        km.timers = timers
        df = km.compute_gap_internal_metric(KMeansB200(n_init=n_init, random_state=0), data, k_max=k_max,
                                            n_references=refs, version=1, draw=draw, group=group, task_parallel=True,
                                            seed=0)
        torch.cuda.synchronize(dev)
        return df

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warm = max(args.warmup, 3)
    for _ in range(warm):                                            # every kernel and shape class once, cheaply
        sweep(X, min(args.c4_kmax, 3), 1, 1)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timers = {}
    e0.record(stream)
    for i in range(args.steps):
        df = sweep(X, args.c4_kmax, args.c4_refs, args.c4_ninit)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms) / args.steps

    # e2e: the call p2 makes - a HOST float32 ndarray in, a DataFrame out (H2D of the matrix and of nothing else;
    # D2H of the labels for the internal metrics' interface and of the K sums per evaluation)
    e2e = None
    if not args.no_e2e:
        Xh = X.cpu().numpy()
        barrier()
        t0 = time.perf_counter()
        df_h = sweep(Xh, args.c4_kmax, args.c4_refs, args.c4_ninit)
        barrier()
        wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        same = bool(np.allclose(df_h.astype(float).values, df.astype(float).values, rtol=1e-9, atol=0))
        e2e = {"value": round(N / float(wall), 1), "unit": "rows/s", "seconds_per_sweep": round(float(wall), 3),
               "h2d_bytes_per_step": int(N * D * 4 * world), "d2h_bytes_per_step": int(8 * 9 * (args.c4_refs + 1) * 10),
               "same_table_as_device_input": same,
               "path": "KM.compute_gap_internal_metric(KMeansB200(n_init), X_host float32 ndarray, k_max, n_references, "
                       "version=1) -> pandas DataFrame: the signature of p2_clustering_optK.py:353; H2D of the matrix inside "
                       "the timed region, reference sets drawn on the device"}

    # phase breakdown (synchronised timers; a separate, reduced sweep so that the timed sweeps run asynchronously)
    sweep(X, args.c4_kmax, 2, 2, timers)
    # kernel rooflines, live: one Lloyd pass (K = 10, the dtype the reference sets have) and the pairwise inertia
    L = _lib.lib()
    ref_dtype = torch.float32 if args.c4_draw == "device32" else torch.float64
    kern = {}
    for name, Xk in (("lloyd_pass_data_f32", X), ("lloyd_pass_refs", X.to(ref_dtype))):
        st = _Device(Xk, 10)
        cen = Xk[:10].clone().contiguous()
        st.assign(cen, 0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(20):
            st.assign(cen, 1 | 4)
        b.record(stream)
        b.synchronize()
        ms = a.elapsed_time(b) / 20
        nbytes = N * D * Xk.element_size() + N * 4
        kern[name] = {"ms": round(ms, 4), "dtype": str(Xk.dtype).replace("torch.", ""), "algorithmic_gb": round(nbytes / 1e9, 4),
                      "gbps": round(nbytes / 1e9 / (ms * 1e-3), 1)}
        del st
    n_pw = min(N, 262_144)
    Xc = X[:n_pw].contiguous()
    pairwise_dist_sum(Xc)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(3):
        s_pw = pairwise_dist_sum(Xc)
    b.record(stream)
    b.synchronize()
    ms = a.elapsed_time(b) / 3
    pairs = n_pw * (n_pw - 1) / 2
    kern["pairwise_inertia"] = {"ms": round(ms, 3), "rows": n_pw, "unordered_pairs_per_s": round(pairs / (ms * 1e-3), 1),
                                # 3 split products x (2 flop per multiply-add) x D per UNORDERED pair: the tile grid is the upper triangle
                                "tflops_dense_equiv": round(3 * 2 * pairs * D / (ms * 1e-3) / 1e12, 2),
                                "note": "3 split-float16 MMAs per product (hi.hi + hi.lo + lo.hi), upper triangle only; "
                                        "one sqrt per pair in the epilogue"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    for k in ("lloyd_pass_data_f32", "lloyd_pass_refs"):
        kern[k]["hbm_frac"] = round(kern[k]["gbps"] / hbm_peak, 4)
    fits = 9 * (args.c4_refs + 1) * args.c4_ninit if args.c4_kmax == 10 else (args.c4_kmax - 1) * (args.c4_refs + 1) * args.c4_ninit
    phase = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in timers.items()}
    fit_s = timers.get("fit_ref", 0) + timers.get("fit_data", 0)
    pw_s = timers.get("inertia_ref", 0) + timers.get("inertia_data", 0)
    dom = "lloyd_pass_refs" if fit_s >= pw_s else "pairwise_inertia"
    if dom == "lloyd_pass_refs":
        roofline = {"kernel": "kmeans_assign (one Lloyd pass over a reference set, K=10)", "bound": "hbm",
                    "achieved": kern[dom]["gbps"], "peak": hbm_peak, "unit": "GB/s", "frac": kern[dom]["hbm_frac"], "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
    else:
        tf_peak = float(peaks.get("bf16_tflops", 1650.0))
        roofline = {"kernel": "pairwise_tc64 (tcgen05 kind::f16, 3 split products per pair)", "bound": "tensor",
                    "achieved": kern[dom]["tflops_dense_equiv"], "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": round(kern[dom]["tflops_dense_equiv"] / tf_peak, 4), "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json (cuBLAS bf16)" if "bf16_tflops" in peaks else "fallback"}
    line = {
        "metric": "latent rows/s through one gap-statistic sweep (K=2..10, 20 reference draws, n_init=10)",
        "value": round(N / (ms_per_step * 1e-3), 1), "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": round(ms_per_step, 1), "seconds_per_sweep": round(ms_per_step / 1e3, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 data / " + ("f32" if args.c4_draw == "device32" else "f64") + " reference sets",
        "data": "synthetic",
        "config": {"workload": CFG_NAME, "rows": N, "dim": D, "k": f"2..{args.c4_kmax}", "n_references": args.c4_refs,
                   "n_init": args.c4_ninit, "kmeans_fits_per_sweep": fits, "pairwise_evaluations_per_sweep": (args.c4_kmax - 1) * (args.c4_refs + 1),
                   "draws": args.c4_draw, "parallelism": f"task-parallel x{world} ((k, reference set) fits dealt round-robin, data replicated)",
                   "warmup_steps": f"{warm} reduced sweeps (K=2..3, 1 reference set, n_init=1): every kernel class warm",
                   "l2": "each matrix (256-512 MB) exceeds L2 (126 MB)"},
        "clocks": clocks, "gpu_launches": "not counted: data-dependent (Lloyd iterations until convergence; ~2-3 launches each)",
        "kernels": kern, "phases_reduced_sweep_2refs_ninit2": phase, "roofline": roofline, "e2e": e2e,
        "gap_table": {str(int(k)): round(float(v), 6) for k, v in zip(df["k"], df["gap"])},
        "best_k": int(df["gap"].astype(float).idxmax()),
    }
    if not args.no_cpu_baseline:
        v, dt, cores, kind = c4_cpu_arm(args.c4_cpu_n, D, args.c4_kmax, args.c4_refs, args.c4_ninit)
        how = "the staged reference's own KM.compute_gap_internal_metric (baseline/_ref)" if kind == "reference" \
            else "oracle/kmeans_oracle.gap_statistic (the reference's routine restated)"
        line["cpu_baseline"] = {"value": round(v, 2), "unit": "rows/s", "cores": cores, "kind": kind,
                                "sample": f"the same sweep (K=2..{args.c4_kmax}, {args.c4_refs} draws, n_init={args.c4_ninit}) on "
                                          f"{args.c4_cpu_n} rows x {D}: {how} over scikit-learn KMeans, {dt:.1f} s; the "
                                          "reference's n_c x n_c distance matrices make N = 1M infeasible on any host (terabytes)"}
    if world > 1:
        dist.destroy_process_group()
    return line


# ----------------------------------------------------------------------------------------------
# p1: one training step of the whole pretrain network
# ----------------------------------------------------------------------------------------------
def _net_args():
    import types
    return types.SimpleNamespace(num_variables=C, num_timestamps=T, ref_points=R, hours_from_admission=HOURS, dropout=0.2,
                                 aux_tasks={}, fake_detection=False, triple_margin=0., cluster_number=K_CLUST)


def p1_cpu_arm(sample, steps=2):
    """The staged reference's own pretrain_interp.Net (unmodified, its own operators, torch CPU) on `sample` encounters:
    forward, rec_loss (:169-175), backward, Adam step.  Returns (encounters/s, seconds per step, cores, kind)."""
    import torch
    from deep_interpolation_clustering_b200 import synth
    from oracle import make_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mod = make_ref.import_net_module("pretrain_interp", b200=False)
    if mod is None:
        return None
    net = mod.Net(_net_args(), torch.device("cpu")).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    x = torch.from_numpy(synth.make_encounters(sample, C, T, HOURS, seed=0))

    def step():
        opt.zero_grad(set_to_none=True)
        hidden, rec, _ = net(x)
        loss = net.rec_loss(x[:, :C], rec, x[:, C:2 * C])["loss"]
        loss.backward()
        opt.step()

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, dt, cores, "reference"


def p1_arm(args, rank, world, local_rank):
    """One step = forward + rec_loss + backward + Adam of the unmodified pretrain_interp.Net (staged reference file) built
    on the B200 mirrors, on B encounters per GPU; gradients all-reduced over the ranks (data parallel)."""
    import torch
    import torch.distributed as dist
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.packed import PackedEncounters, PackedStaging
    from oracle import make_ref
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mod = make_ref.import_net_module("pretrain_interp", b200=True)
    if mod is None:
        raise SystemExit("--workload p1 needs the staged reference (python oracle/make_ref.py where /root/reference exists)")
    B = args.encounters
    torch.manual_seed(0)
    net = mod.Net(_net_args(), dev).to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    x = synth.make_encounters_device(B, C, T, HOURS, 5.0, 1000 * rank, dev)
    params = [p for p in net.parameters()]

    def step(xb):
        opt.zero_grad(set_to_none=True)
        hidden, rec, _ = net(xb)
        loss = net.rec_loss(xb[:, :C], rec, xb[:, C:2 * C])["loss"]
        loss.backward()
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat /= world
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step(x)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        loss = step(x)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms) / args.steps

    # kernel shares of one step (CUPTI): our kernels (dic::*) vs library kernels
    shares, launches = {}, 0
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(x)
            torch.cuda.synchronize(dev)
        tot = sum(e.device_time_total for e in prof.key_averages()) or 1.0
        for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]:
            shares[e.key[:70]] = {"ms": round(e.device_time_total / 1e3, 3), "share": round(e.device_time_total / tot, 4),
                                  "launches": e.count}
        launches = int(sum(e.count for e in prof.key_averages() if "dic::" in e.key))
    except Exception:       # noqa: BLE001
        pass

    # e2e: pinned host batch (packed ragged rows) -> upload -> the same step -> loss back on the host
    e2e = None
    if not args.no_e2e:
        host = x.cpu()
        pk = PackedEncounters.from_dense(host)
        staging = PackedStaging.for_chunks(pk, B, dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            xb = staging.upload(pk, 0, B)
            lv = float(step(xb))
        barrier()
        wall = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        e2e = {"value": round(world * B / float(wall), 1), "unit": "encounters/s", "ms_per_step": round(float(wall) * 1e3, 3),
               "h2d_bytes_per_step": int(pk.nbytes(0, B)) * world, "d2h_bytes_per_step": 4 * world, "loss": lv,
               "path": "pinned host PackedEncounters -> PackedStaging.upload -> pretrain_interp.Net (unmodified staged "
                       "reference file on the B200 mirrors incl. the BiLSTMs) fwd + rec_loss + bwd + Adam -> loss to the host"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    line = {
        "metric": "encounters/s (full pretrain Net step: interp + BiLSTM enc/dec + read-out, fwd+bwd+Adam)",
        "value": round(world * B / (ms_per_step * 1e-3), 1), "unit": "encounters/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CFG_NAME, "encounters_per_gpu": B, "vitals": C, "max_obs": T, "ref_points": R,
                   "parameters": int(sum(p.numel() for p in params)), "parallelism": f"data-parallel x{world}",
                   "l2": f"activations ({B * R * 1024 * 4 / 1e9:.1f} GB of gate pre-activations alone) exceed L2"},
        "clocks": clocks, "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
        "kernels": shares, "loss": float(loss), "e2e": e2e,
    }
    k = next((v for n, v in shares.items() if "lstm_fwd_kernel" in n), None)
    if k:
        # the recurrence: 2 B R 2 x 128 x 512 algorithmic flops per forward launch, two launches (encoder, decoder) per step
        flops = 2.0 * B * R * 2 * 128 * 512 * k["launches"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:       # noqa: BLE001
            pass
        peak = float(peaks.get("bf16_tflops", 1653.3))
        line["roofline"] = {"kernel": "lstm_fwd_kernel (persistent BiLSTM recurrence)", "bound": "tensor",
                            "achieved": round(flops / (k["ms"] * 1e-3) / 1e12, 2), "peak": peak, "unit": "TFLOP/s",
                            "frac": round(flops / (k["ms"] * 1e-3) / 1e12 / peak, 4), "traffic": None,
                            "note": "algorithmic float32 flops of h W_hh^T; the kernel issues 3 fp16 MMAs per product (split "
                                    "operands) and its step time is set by the MUFU-bound gate epilogue and the cluster "
                                    "barrier, not by the tensor pipe (profiles/r02_ncu_lstm_fwd_summary.txt)"}
    if not args.no_cpu_baseline:
        r = p1_cpu_arm(min(args.cpu_sample, 64))
        if r is not None:
            v, dt, cores, kind = r
            line["cpu_baseline"] = {"value": round(v, 1), "unit": "encounters/s", "cores": cores, "kind": kind,
                                    "sample": f"{min(args.cpu_sample, 64)} encounters through the staged reference's own "
                                              f"pretrain_interp.Net on its own operators (torch CPU), fwd + rec_loss + bwd + "
                                              f"Adam, {dt:.2f} s/step"}
    if world > 1:
        dist.destroy_process_group()
    return line


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE the pinned host buffers are
    allocated and first touched, so that their pages live on the GPU's NUMA node (with 8 ranks streaming from
    host memory at once, remote-node pages halve the aggregate H2D rate).  Returns the previous affinity."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if cpus:
            os.sched_setaffinity(0, cpus)
            return prev
    except Exception:
        pass
    return None


def e2e_arm(args, hp, dev, world):
    """Public-API path with host buffers: pinned x chunks -> H2D -> nn.Module fwd/bwd (autograd) + DEC step
    -> D2H of loss, parameter grads and labels.  Double-buffered on a copy stream."""
    import torch
    import torch.distributed as dist
    import deep_interpolation_clustering_b200 as dic
    from deep_interpolation_clustering_b200 import functional as F_
    Bc = min(args.e2e_chunk, hp.B)
    n_chunks = max(1, min(args.e2e_encounters, hp.B) // Bc)
    from deep_interpolation_clustering_b200 import PackedEncounters, PackedStaging
    packed_mode = args.e2e_upload == "packed"
    host = [hp.x[i * Bc:(i + 1) * Bc].cpu() for i in range(min(2, n_chunks))]
    if packed_mode:      # packed once per data set, outside the timed region (what p0_data_process.py's rows already are)
        host = [PackedEncounters.from_dense(h) for h in host]
        staging = [PackedStaging(C, Bc, max(h.floats() for h in host), dev) for _ in range(2)]
        chunk_bytes = [h.nbytes() for h in host]
    else:
        host = [h.pin_memory() for h in host]
        chunk_bytes = [Bc * 3 * C * T * 4 for _ in host]
    dbuf = [torch.empty((Bc, 3 * C, T), dtype=torch.float32, device=dev) for _ in range(2)]   # live planes only
    sci = dic.SingleChannelInterp(R, HOURS, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    rbf = dic.RBF(HOURS, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()
    sci.kernel.data, cci.kernel.data, rbf.kernel.data = hp.k_sci.clone(), hp.k_cci.clone(), hp.k_rbf.clone()
    params = [sci.kernel, cci.kernel, rbf.kernel]
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    host_out = torch.empty(C + C * C + C + 1, dtype=torch.float32).pin_memory()
    host_lab = torch.empty(Bc, dtype=torch.int32).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def upload(k, slot):
        if packed_mode:
            staging[slot].upload(host[k], 0, Bc, out=dbuf[slot], stream=copy_stream)
        else:
            F_.upload_encounters(host[k], out=dbuf[slot], stream=copy_stream)

    # The uploads run as ONE continuous prefetch stream, like a data loader with one batch of look-ahead: while chunk
    # g is computed, chunk g + 1 is on the wire - also across a step boundary (the first chunk of step s + 1 travels
    # while the last chunk of step s is computed and its results are read back).
    state = {"next": 0}        # global index of the next chunk to upload; chunk g lives in slot g & 1

    def prefetch():
        g = state["next"]
        slot = g & 1
        with torch.cuda.stream(copy_stream):
            if g >= 2:
                copy_stream.wait_event(free[slot])        # the compute that last read this slot is done
            upload(g % len(host), slot)
            ready[slot].record(copy_stream)
        state["next"] = g + 1

    def step():
        for p_ in params:
            p_.grad = None
        loss_acc = torch.zeros((), device=dev)
        for i in range(n_chunks):
            if state["next"] == 0:
                prefetch()                                 # very first chunk of the run
            g = state["next"] - 1
            cur = g & 1
            prefetch()                                     # chunk g + 1 (possibly the next step's first)
            main.wait_event(ready[cur])
            x = dbuf[cur]
            sl = slice(i * Bc, (i + 1) * Bc)
            v = hp.v[sl].detach().requires_grad_(True)
            out = cci(sci(x))
            rec = rbf(v, x)
            m = x[:, C:2 * C]
            loss = (out.permute(0, 2, 1) * hp.g_out[sl]).sum() / Bc + ((rec * m - x[:, :C] * m) ** 2).sum() / m.sum()
            loss.backward()
            res = F_.dec_kl_step(hp.z[sl], hp.mu, 1.0, weight=10.0, want_grad_z=False)
            loss_acc = loss_acc + loss.detach()
            host_lab.copy_(res["labels"], non_blocking=True)
            free[cur].record(main)
        packed = torch.cat([sci.kernel.grad, cci.kernel.grad.flatten(), rbf.kernel.grad, loss_acc[None]])
        if world > 1:
            dist.all_reduce(packed)
        host_out.copy_(packed, non_blocking=True)
        main.synchronize()

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    n = max(2, args.steps // 2)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(n):
        step()
    e1.record(main)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    per_step = float(t_ms) / n
    enc = n_chunks * Bc
    return {"value": round(world * enc / (per_step * 1e-3), 1), "unit": "encounters/s",
            "h2d_bytes_per_step": int(sum(chunk_bytes[i % len(host)] for i in range(n_chunks))), "d2h_bytes_per_step": int(enc * 4 + host_out.numel() * 4),
            "encounters_per_step_per_gpu": enc, "chunk": Bc, "ms_per_step": round(per_step, 3),
            "upload": args.e2e_upload,
            "path": ("pinned host PackedEncounters (ragged rows: valid prefix of value/time + one count per vital, packed "
                     "once per data set outside the timed region) -> PackedStaging.upload (3 contiguous H2D copies + "
                     "device-side expansion to dense planes per chunk, copy stream, double buffered, one chunk of look-ahead "
                     "also across step boundaries)" if packed_mode else
                     "pinned host x (B,4C,T) -> upload_encounters (one strided DMA of the 3 live planes per chunk, copy "
                     "stream, double buffered)") +
                    " -> SingleChannelInterp/CrossChannelInterp/RBF modules (autograd fwd+bwd; RBF.compress_fc = Identity, "
                    "v and the latents z stay device-resident: the metric's step has no encoder) + dec_kl_step -> D2H labels, "
                    "loss, parameter grads"}


def main():
    global T, R, K_CLUST, CFG_NAME
    args = parse()
    w = WORKLOADS[args.workload]
    T, R, K_CLUST, CFG_NAME = w["T"], w["R"], w["K"], w["name"]
    if args.encounters <= 0:
        args.encounters = w["B"]
    if args.workload == "c5":                  # 98 KB of x per encounter: keep the e2e chunks and CPU sample small
        args.e2e_encounters = min(args.e2e_encounters, 32_768)
        args.e2e_chunk = min(args.e2e_chunk, 8_192)
        args.cpu_sample = min(args.cpu_sample, 16)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and w.get("sweep"):
        if rank != 0:
            return
        v, dt, cores, kind = c4_cpu_arm(args.c4_cpu_n, args.c4_dim, args.c4_kmax, args.c4_refs, args.c4_ninit)
        print(json.dumps({"impl": "reference", "metric": "latent rows/s through one gap-statistic sweep (K=2..10, 20 reference "
                          "draws, n_init=10)", "value": round(v, 2), "unit": "rows/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                          "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f32 data / f64 reference sets", "data": "synthetic",
                          "config": {"workload": CFG_NAME, "rows": args.c4_cpu_n, "dim": args.c4_dim},
                          "cpu_baseline": {"value": round(v, 2), "unit": "rows/s", "cores": cores, "kind": kind,
                                           "sample": f"{args.c4_cpu_n} rows (the n_c x n_c matrices cap the reference)"},
                          "e2e": {"value": round(v, 2), "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
        v, dt, cores, kind = cpu_arm(args.cpu_sample, steps, warmup)
        line = {"impl": "reference", "metric": "encounters/s (interp fwd+bwd + DEC assign)", "value": round(v, 1),
                "unit": "encounters/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": CFG_NAME, "vitals": C, "max_obs": T, "ref_points": R, "hours": HOURS,
                           "latent_dim": D_LAT, "clusters": K_CLUST},
                "cpu_baseline": {"value": round(v, 1), "unit": "encounters/s", "cores": cores, "kind": kind,
                                 "sample": f"each step = {args.cpu_sample} encounters of this workload's shape (bounded sample; "
                                           "the reference's (B,C,T,R) temporaries cap its batch) through " +
                                           ("the reference's own unmodified modules (interpolation_layer.py, rbf.py, dec.py) "
                                            "staged in baseline/_ref by oracle/make_ref.py" if kind == "reference" else
                                            "oracle/ref_port.py, the torch-CPU restatement of the reference (no staged "
                                            "reference on this box)")},
                "e2e": {"value": round(v, 1), "unit": "encounters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        sys.stderr.write(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks; running 1 rank\n")
    if w.get("net"):
        line = p1_arm(args, rank, world, local_rank)
    elif w.get("sweep"):
        if args.steps == 5:            # the default step count is for the interp workloads; one sweep is ~1 minute
            args.steps = 1
        line = c4_arm(args, rank, world, local_rank)
    else:
        line = device_arm(args, rank, world, local_rank)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
