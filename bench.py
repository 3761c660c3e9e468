#!/usr/bin/env python
"""bench.py — PHD filter updates/s of the B200 PHD measurement-update path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C3|C2|C3mf|C5|...] [--impl b200|reference]

A "step" is one RBPHDFilter::update() (map update + particle weighting + merge + prune + weight
sums (+ all-reduce) + normalisation) over one batch of synthetic particles; an "update" is one
(particle x GM component x measurement) cell, so  value = N_total * nM_in * nZ / t_step.
Weak scaling: every GPU owns 8 000 particles (C3); 8 GPUs = BASELINE config C4 (64 000).

One JSON line is printed by rank 0 (see DESIGN.md "Measurement" for every key).
Only the cpu_baseline leg and --impl reference execute anything under oracle/ (the CPU checker /
the reference's own sources compiled into oracle/_ref); the timed GPU path never does.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PHD filter updates/sec (particles x GM x meas)"
UNIT = "updates/s"
L2_FLUSH_BYTES = 512 << 20

WORKLOADS = {
    # name: (synth config, particles per GPU, description)
    "C3": ("C3", 8000, "C3/C4: 8000 particles per GPU x 200 GM x 30 meas, RngBrg, SC-PHD weighting (8 GPUs = C4, 64000 particles)"),
    "C2": ("C2", 1000, "C2: 1000 particles per GPU x 100 GM x 20 meas, RngBrg, multi-feature weighting"),
    "C3mf": ("C3", 8000, "C3 shape with multi-feature weighting: 8000 particles per GPU x 200 GM x 30 meas"),
    "C3sparse": ("C3", 8000, "C3 shape, sparse world (about 11 % of the components in range), SC-PHD"),
    "N1k": ("C3", 1000, "north_star sweep: 1000 particles per GPU x 200 GM x 30 meas, SC-PHD"),
    "N64k": ("C3", 64000, "north_star sweep: 64000 particles per GPU x 200 GM x 30 meas, SC-PHD"),
    "C5": ("C5", 4000, "C5 shape: 4000 particles per GPU x 150 GM (3-D: x, y, diameter) x 12 meas, MeasurementModel_VictoriaPark "
                       "(P_D from a 720-beam lidar scan), multi-feature weighting"),
}


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(name, rank, n_override=0):
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import synth
    cfgname, n_gpu, desc = WORKLOADS[name]
    kw = dict(N=n_override or n_gpu, shard_id=rank)
    if name == "C3mf":
        kw["use_cluster_process"] = 0
    if name == "C3sparse":
        kw["world"] = "sparse"
    return synth.make_config(cfgname, **kw), desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, n_particles, threads, repeats):
    """Time the reference's own OpenMP RBPHDFilter::update() (oracle/_ref, compiled from the
    reference sources) — or the oracle port if that library is absent — on the first
    n_particles particles of the workload.  Returns (kind, best_seconds, list of seconds)."""
    from oracle import binding as ob
    sub = wl.shard(0, max(1, wl.N // n_particles)) if n_particles < wl.N else wl
    kind = "reference" if ob.have_ref() else "port"
    which = "ref" if kind == "reference" else "oracle"
    stage = 5 if kind == "reference" else ob.STAGE_FULL   # 5: the public update() incl. normalizeWeights
    ts = []
    for _ in range(repeats):
        r = ob.run(sub, which=which, stage=stage, n_threads=threads)
        ts.append(r.elapsed_s)
    return kind, sub, ts


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl, desc = make_workload(a.config, 0)
    nM = int(round(wl.count.mean()))
    threads = len(os.sched_getaffinity(0))
    n_s = min(wl.N, a.ref_particles)   # the whole C3 shard (8000 particles, about 0.6 s per step on 16 threads)
    kind, sub, _ = cpu_reference_run(wl, n_s, threads, max(0, a.warmup))
    _, _, ts = cpu_reference_run(wl, n_s, threads, a.steps)
    t = sum(ts) / len(ts)
    units = int(sub.count.sum()) * sub.nZ
    v = units / t
    sample = f"{sub.N} of {wl.N} particles of the workload per step ({'oracle/_ref: reference sources, OpenMP' if kind == 'reference' else 'oracle port'}, {threads} threads)"
    single = None
    try:   # the same path on ONE host thread, on a quarter of the sample
        _, sub1, t1 = cpu_reference_run(wl, max(1, n_s // 4), 1, 1)
        single = dict(value=int(sub1.count.sum()) * sub1.nZ / t1[0], unit=UNIT, cores=1,
                      sample=f"{sub1.N} particles, one run ({t1[0]:.3f} s)")
    except Exception:
        single = None
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=1e3 * t, higher_is_better=True, scaling=a.scaling, vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=desc, particles_per_step=sub.N, gm_per_particle=nM, meas=sub.nZ),
                cpu_baseline=dict(value=v, unit=UNIT, cores=threads, kind=kind, sample=sample, single_thread=single),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def _kernel_fingerprint():
    """sha256 of the kernel sources: an ncu figure recorded for another build of the kernel is never reported."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "rfs-slam_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def _ncu_record(config):
    """(dram bytes per launch, warp instructions per launch) of the last committed ncu capture of this workload — only if
    it was taken from the kernel sources that are running now (profiles/traffic.json carries their fingerprint)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(tp))
    except Exception:
        return None, None, "no capture"
    rec = d.get(config)
    if not isinstance(rec, dict):
        return None, None, "no capture of this workload"
    if rec.get("kernel_fingerprint") != _kernel_fingerprint():
        return None, None, "the committed ncu capture is of another build of the kernel"
    return rec.get("dram_bytes"), rec.get("warp_instructions"), rec.get("capture")


def run_workload(a, name, ctx, steps, warmup, headline):
    """One workload on this rank's GPU: device-resident steps, the host-facing step, the stage times, the live roofline.
    Returns the JSON line (rank 0) or None."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from rfs_slam_b200 import capi
    from rfs_slam_b200.dist import ShardedUpdater
    from rfs_slam_b200.phd import PHDUpdater, pinned_array

    rank, world, local, dev, flush = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"], ctx["flush"]
    n_override = a.particles
    if a.scaling == "strong":   # the total is fixed, every GPU owns 1 / world of it
        total = a.particles_total or WORKLOADS[name][1]
        n_override = max(1, total // world)
    wl, desc = make_workload(name, rank if world > 1 else int(os.environ.get("RFSB200_BENCH_SHARD", "0")), n_override)
    N, nZ = wl.N, wl.nZ
    units_local = int(wl.count.sum()) * nZ
    D = wl.dim                                  # 2: RngBrg, 3: VictoriaPark
    gbytes = 4.0 * (D + D * (D + 1) // 2 + 1)   # fp32 bytes per Gaussian: 24 (2-D) / 40 (3-D)
    # capacities per particle: 256 Gaussians (2-D workloads: up to 200 in + the corrector's); C5 holds 150 in, 192 is enough
    # and buys the Victoria Park kernel three more resident warps per SM (overflows would show in n_overflow_per_step)
    cap = 192 if name == "C5" else 256
    up = PHDUpdater(N, gm_capacity=cap, z_capacity=32, device=local, precision=32, lmk_dim=D)
    up.load_workload(wl)
    fused = not a.no_fused
    sh = ShardedUpdater(up, device=dev, fused=fused)
    fused = sh.fused
    up.synchronize()
    FLAGS = capi.UPDATE_NO_COMMIT   # every step starts from the same state, so nM_in is constant

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also warms NCCL) -------------------------------------------------------------
    for _ in range(warmup):
        flush.zero_()
        sh.step(wl.Z, flags=FLAGS)
    torch.cuda.synchronize()
    so = up.update(wl.Z, flags=FLAGS | capi.UPDATE_NO_NORMALIZE)   # statistics of one step
    nM_out_mean = so.gm_total_out / N
    n_murty = int(so.n_murty)
    n_overflow = int(so.n_overflow)

    # ---- cross-GPU check (untimed): the in-kernel sum over peer memory against the NCCL all-reduce ----------------
    collective_check = None
    w_f = None
    if world > 1 and fused:
        # the sums the kernels exchanged through the peer mailboxes (added in rank order, the same bits on every rank) and
        # the weights normalised with them, against: every rank's local sums all-gathered by NCCL, added in rank order on
        # the host, and the locally normalised weights.  (NCCL's own all-reduce adds in ring / tree order, which may round
        # differently from the rank order for more than two ranks; the NCCL path of the library is timed with --no-fused.)
        sh.step(wl.Z, flags=FLAGS)
        w_f = up.get_weights(1).copy()
        s_f = sh.sums.cpu().numpy().copy()
        up.update(wl.Z, flags=FLAGS | capi.UPDATE_NO_NORMALIZE, want_stats=False)
        w_u = up.get_weights(1).copy()
        loc = sh.sums.clone()
        parts = [torch.zeros_like(loc) for _ in range(world)]
        dist.all_gather(parts, loc)
        tot = np.zeros(2)
        for q in parts:
            tot += q.cpu().numpy()
        same = bool(np.array_equal(s_f.view(np.uint64), tot.view(np.uint64)) and
                    np.array_equal(w_f.view(np.uint64), (w_u / tot[0]).view(np.uint64)))
        t = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        collective_check = "bit-identical" if int(t.item()) == 1 else "DIFFERS"
        if up.comm_error():
            collective_check = "comm_error"

    # ---- timed region: K steps, device-timed, L2 flushed between steps -----------------------------
    K = steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    up.profile_begin(K)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    t_wall0 = time.perf_counter()

    def flush_and_align():
        flush.zero_()                 # > L2 (126 MB): the step reads its inputs from HBM
        if world > 1:
            # the 512 MiB memset is a measurement artefact whose duration differs from GPU to GPU; back-to-back
            # steps of a real run start aligned by the previous step's exchange, so the ranks are re-aligned here
            # (untimed, on the launch stream) before the step's first event: the library's own mailbox barrier when the
            # peer mailboxes are connected (ranks leave it within an NVLink round trip), else a 4-byte all-reduce
            if fused:
                up.comm_barrier()
            else:
                dist.all_reduce(ctx["align"])

    # More than one GPU, peer mailboxes connected: the steps run with the cross-GPU sums DEFERRED
    # (RFSB200_UPDATE_DEFER_NORMALIZE): a step sends its [sum w, sum w^2] and ends, the next one picks the pairs up during
    # set-up (they arrived a step ago; with the committed state it would divide the weights by their total on the way in
    # — here every step restarts from the same state, so the open weights are those of the back buffer, which the next
    # step overwrites), and the LAST step's normalisation is closed inside the timed region (rfsb200_comm_resolve).
    # --eager-exchange times the variant in which every step waits for all peers and normalises before it ends.
    defer = world > 1 and fused and not a.eager_exchange
    ev_close = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for k in range(K):
        flush_and_align()
        ev[k][0].record()
        sh.step(wl.Z, flags=FLAGS, defer=defer)    # fused update kernel (+ cross-GPU sum, + normalise unless deferred)
        ev[k][1].record()
    ev_close[0].record()
    if defer:
        sh.resolve()
    ev_close[1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if rank == 0 else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    close_ms = ev_close[0].elapsed_time(ev_close[1]) if defer else 0.0
    t_local = (sum(step_ms) + close_ms) / 1e3
    deferred_check = None
    if defer and w_f is not None:   # the last deferred step, closed by rfsb200_comm_resolve, against the eager step of the check above
        same = bool(np.array_equal(up.get_weights(1).view(np.uint64), w_f.view(np.uint64)))
        t = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        deferred_check = "bit-identical" if int(t.item()) == 1 else "DIFFERS"
    eager_ms = None
    if defer:   # the eager exchange beside it (untimed for `value`): same steps, every one waits for the slowest rank
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        for k in range(K):
            flush_and_align()
            ev2[k][0].record()
            sh.step(wl.Z, flags=FLAGS)
            ev2[k][1].record()
        barrier()
        te = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in ev2) / K], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        eager_ms = float(te.item())
    kern_us = up.profile_read()
    tt = torch.tensor([t_local], dtype=torch.float64, device=dev)
    uu = torch.tensor([float(units_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(uu, op=dist.ReduceOp.SUM)
    t_max = float(tt.item())
    units_total = float(uu.item())
    value = units_total * K / t_max
    per_rank = None
    if world > 1:   # what every rank saw: its own step time and its kernel time (the wait for the slowest peer included)
        mine = torch.tensor([1e3 * t_local / K, float(np.mean(kern_us)) if len(kern_us) else 0.0], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = dict(ms_per_step=[float(x[0]) for x in allr], kernel_us=[float(x[1]) for x in allr])
    comm_err = int(up.comm_error()) if world > 1 and fused else 0

    # ---- e2e: the same step through the C ABI with HOST buffers, timed on the HOST clock ------------------------
    e2e = None
    if not a.no_e2e:
        h_pose = pinned_array((N, 3)); h_pose[:] = wl.pose
        h_w = pinned_array((N,)); h_w[:] = wl.weight
        h_wout = pinned_array((N,))
        h_mask = pinned_array((N,), np.uint64)
        h_nfov = pinned_array((N,), np.int32)
        pc = wl.pose_cov
        md_desc = capi.model_desc(wl.model)
        Zc = np.ascontiguousarray(wl.Z, dtype=np.float64)

        def e2e_step():
            if D == 3:                                        # the lidar scan changes before every update
                up.set_model_desc(md_desc)                    # H2D: scan (staged through pinned memory)
            # H2D: poses + particle weights (pinned) + Z ; kernels (+ cross-GPU sum) ; D2H: normalised particle
            # weights, unused-measurement masks, nLandmarksInFOV — one ABI call, one synchronisation (fused path)
            return sh.step_host(h_pose, pc, h_w, Zc, flags=FLAGS, w_out=h_wout, unused_out=h_mask, nfov_out=h_nfov)

        for _ in range(3):
            e2e_step()
        barrier()
        Ke = K
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        te_host, te_dev = 0.0, 0.0
        for _ in range(Ke):
            flush_and_align()
            torch.cuda.synchronize()          # the flush (and the alignment) are over: nothing of the step can hide under them
            e0.record()
            t0 = time.perf_counter()
            e2e_step()                        # returns after its own synchronisation: the results are in the host buffers
            te_host += time.perf_counter() - t0
            e1.record()
            e1.synchronize()
            te_dev += e0.elapsed_time(e1) / 1e3
        barrier()
        tt = torch.tensor([te_host, te_dev], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te_max, td_max = float(tt[0].item()), float(tt[1].item())
        h2d = N * 3 * 8 + N * 8 + (6 * 8 if pc is not None else 0) + nZ * D * 8 + (len(wl.model["scan"]) * 8 if D == 3 else 0)
        d2h = N * 8 + N * 8 + N * 4
        e2e = dict(value=units_total * Ke / te_max, unit=UNIT, h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=d2h * world,
                   ms_per_step=1e3 * te_max / Ke, timer="host", device_ms=1e3 * td_max / Ke,
                   what=("host perf_counter around ONE rfsb200_update_host call per step (L2 flush finished and synchronised "
                         "before the clock starts): pinned host poses + particle weights + Z in, update (+ cross-GPU sum + "
                         "normalisation), normalised weights + unused-measurement masks + in-FOV counts out in pinned host "
                         "buffers when the call returns; maps stay resident in HBM; max over ranks. device_ms = the same "
                         "span between CUDA events"
                         if fused else
                         "host perf_counter around rfsb200_set_poses(pinned host poses+weights) + rfsb200_update(host Z) + NCCL "
                         "all-reduce + rfsb200_normalize + rfsb200_get_weights + rfsb200_get_unused into pinned host buffers; maps stay resident"))

    # ---- stage times (untimed extra steps with the stage-timing build of the kernel) ------------------------------
    stages = None
    if D == 2 and not a.no_stages:
        try:
            for _ in range(2):
                flush.zero_()
                up.update(wl.Z, flags=FLAGS | capi.UPDATE_NO_NORMALIZE | capi.UPDATE_STAGE_TIMES)
            stages = up.stage_times()
            stages["note"] = ("one fused kernel: wall time split into set-up / particle loop / epilogue by the device global "
                              "timer, the particle loop attributed to the phases by warp cycles (TimingInfo of the reference)")
        except Exception as e:   # noqa: BLE001
            stages = dict(error=str(e))

    # ---- roofline of the dominant kernel (phd_update_kernel), live ------------------------------------
    peak, peak_src = _peaks()
    nM_in = int(wl.count.sum()) / N
    per_particle = 60.0 if D == 2 else 36.0   # pose 12 + pose cov 24 (2-D model only) + weight r/w 16 + count r/w 8
    b_alg = N * (gbytes * nM_in + gbytes * nM_out_mean + per_particle) + 4.0 * D * nZ      # SURVEY §8(d), DESIGN.md
    if D == 3:
        b_alg += 8.0 * len(wl.model["scan"])
    k_us = float(np.mean(kern_us)) if len(kern_us) else float("nan")
    achieved = b_alg / (k_us * 1e-6) / 1e9
    traffic, n_inst, cap_src = (None, None, None)
    if N == WORKLOADS[name][1] and a.scaling == "weak":
        traffic, n_inst, cap_src = _ncu_record(name)
    issue = None
    if n_inst and clk is not None:
        sm_hz = 1e6 * float(clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0)
        slots = 148 * 4 * sm_hz * k_us * 1e-6
        issue = dict(warp_instructions=int(n_inst), issue_slots=slots, frac=n_inst / slots,
                     note="instruction-issue roofline: the kernel is issue / latency bound, not HBM bound")
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, issue=issue,
                    ncu_capture=cap_src,
                    kernel=("phd_update_kernel<float>" if D == 2 else "phd_update_vp_kernel<float>"), kernel_us=k_us, algorithmic_bytes=b_alg, peak_source=peak_src,
                    kernel_share_of_step=k_us * 1e-3 / (1e3 * t_local / K))

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and headline and not a.no_cpu_baseline:
            threads = len(os.sched_getaffinity(0))
            n_s = min(N, 8000)
            kind, sub, ts = cpu_reference_run(wl, n_s, threads, 2)
            t = min(ts)
            cpu = dict(value=int(sub.count.sum()) * nZ / t, unit=UNIT, cores=threads, kind=kind,
                       sample=f"{sub.N} particles of the same workload, best of {len(ts)} runs of the reference's OpenMP "
                              f"RBPHDFilter::update() ({t:.3f} s per step, {threads} threads; built -O3 -march=x86-64-v3 -fopenmp: the "
                              f"library travels to the GPU box, so not -march=native)")
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=warmup,
                    ms_per_step=1e3 * t_max / K, higher_is_better=True, scaling=a.scaling, vs_baseline=None,
                    dtype="f32", data="synthetic",
                    config=dict(workload=desc, particles_total=N * world, particles_per_gpu=N, gm_per_particle=nM_in,
                                gm_out_per_particle=nM_out_mean, meas=nZ,
                                l2="flushed between steps (512 MiB memset, untimed"
                                   + (", followed by an untimed barrier that re-aligns the ranks)" if world > 1 else ")"),
                                timing="CUDA events per step on the launch stream, summed; max over ranks",
                                parallelism=f"particles block-partitioned over {world} GPU(s); one all-reduce of [sum w, sum w^2] per step "
                                            + ("inside the update kernel over NVLink peer memory" if fused else "by NCCL")
                                            + ("; consumed DEFERRED: a step sends its pair and ends, the next step picks the peers' pairs up during "
                                               "set-up, the last step's normalisation is closed inside the timed region (see `exchange`)" if defer else "")),
                    e2e=e2e, gpu_launches=(1 if fused else 2) * K + (1 if defer else 0), roofline=roofline, cpu_baseline=cpu, clocks=clk,
                    stages=stages, n_murty_per_step=n_murty, n_overflow_per_step=n_overflow,
                    wall_s_timed_region_incl_flush=t_wall)
        if world > 1:
            line["collective_check"] = collective_check
            line["comm_error"] = comm_err
            line["per_rank"] = per_rank
            if fused:
                line["exchange"] = dict(
                    mode="deferred" if defer else "eager",
                    close_ms=close_ms if defer else None,
                    deferred_check=deferred_check,
                    eager_ms_per_step=eager_ms,
                    note="deferred: RFSB200_UPDATE_DEFER_NORMALIZE — no rank waits for the slowest one inside a step; results bit-identical to the "
                         "eager exchange (tests/test_gpu_multi.py); `close_ms` = rfsb200_comm_resolve after the last step, inside the timed "
                         "region and inside `value`; `eager_ms_per_step` = the same steps with the wait + normalisation in every launch "
                         "(--eager-exchange), max over ranks")
    up.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU owns the workload's particle count; strong: the total is fixed and split over the GPUs")
    ap.add_argument("--particles-total", type=int, default=0, help="--scaling strong: total particles (default: the workload's count)")
    ap.add_argument("--ref-particles", type=int, default=8000, help="--impl reference: particles per step (capped at the workload's)")
    ap.add_argument("--extras", default="auto",
                    help="other workloads appended to the headline line as `extra` sub-lines: 'auto' = C2,C3mf,C5,N1k,N64k on the "
                         "default single-GPU C3 run, 'none', or a comma-separated list")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stages", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="NCCL all-reduce + normalise kernel instead of the in-kernel sum over peer memory")
    ap.add_argument("--eager-exchange", action="store_true", help="N > 1: every step waits for the peers' sums and normalises before it ends "
                    "(default: deferred — the next step picks them up, the last one is closed inside the timed region)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 0)
    # stdout carries exactly ONE JSON line: everything libraries print (NCCL's version banner, ...)
    # goes to stderr, the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        return run_reference_arm(a)
    if a.warmup < 3:
        a.warmup = 3   # timing rule: at least 3 warm-up steps

    import torch
    import torch.distributed as dist
    import rfs_slam_b200  # noqa: F401

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the PHD update path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # one non-default stream carries the library's kernels, torch's events / memsets and NCCL
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = dict(rank=rank, world=world, local=local, dev=dev,
               flush=torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev),
               align=torch.zeros(1, dtype=torch.float32, device=dev))

    line = run_workload(a, a.config, ctx, a.steps, a.warmup, headline=True)

    # ---- the other workloads of the north_star sweep, as sub-lines of the same record -----------------------------
    extras = []
    if a.extras == "auto":
        if world == 1 and a.config == "C3" and a.scaling == "weak" and not a.particles:
            extras = ["C2", "C3mf", "C5", "N1k", "N64k"]
    elif a.extras != "none":
        extras = [x for x in a.extras.split(",") if x in WORKLOADS]
    if extras:
        sub = {}
        for name in extras:
            try:
                l = run_workload(a, name, ctx, max(5, min(a.steps, 20)), max(3, min(a.warmup, 5)), headline=False)
            except Exception as e:   # noqa: BLE001  (a failing extra must not take the headline with it)
                l = dict(error=f"{type(e).__name__}: {e}") if rank == 0 else None
            if rank == 0 and l is not None:
                keep = ("value", "unit", "ms_per_step", "steps", "warmup", "e2e", "roofline", "stages", "config", "n_murty_per_step",
                        "n_overflow_per_step", "error")
                sub[name] = {k: l[k] for k in keep if k in l}
        if rank == 0 and line is not None:
            line["extra"] = sub
    if rank == 0 and line is not None:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
