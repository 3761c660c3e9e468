"""The CUDA kernel SOURCES, interpreted on the host (tests/simt/: every thread a fiber, warp / CTA primitives as
rendezvous points), against the reference's golden vectors and the pinned oracle — so that the CPU suite checks the
kernel logic itself and not only the oracle and the host code.  This is test infrastructure: the package never loads
the interpreter build, nothing here is a measurement, and the parity tests proper remain the `-m gpu` ones, which run
the sm_100a build through the same C ABI.  What the interpreter adds to them: the most adversarial lane interleaving
(a lane runs alone until its next rendezvous), NaN-filled dynamic shared memory and garbage-filled device memory,
every warps-per-CTA shape, and randomised sweeps that cost no GPU time.
"""
import inspect
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers
from helpers import TOL32, TOL64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("simt_host", os.path.join(ROOT, "tests", "simt", "host.py"))
host = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(host)


@pytest.fixture(scope="module")
def simt_lib():
    return host.load()


def _compare(wl, prec, ref, cnt, mean, cov, w, pw):
    tol = TOL32 if prec == 32 else TOL64
    r = helpers.compare_maps(cnt, mean, cov, w, ref["count"], ref["mean"], ref["cov"], ref["w"], tol)
    rw = helpers.compare_weights(pw, ref["weight"], tol)
    bad = set(r["bad"]) | set(int(i) for i in rw["idx_bad"])
    if prec == 32 and bad:
        robust = helpers.robust_mask(wl)
        bad = {i for i in bad if robust[i]}
    assert not bad, f"particles differ from the reference: {sorted(bad)[:10]}"


def test_interpreter_primitives(tmp_path):
    """the interpreter itself: shuffles, votes, partial masks, early exits, __syncthreads, atomics, mbarrier"""
    exe = str(tmp_path / "selftest")
    simt = os.path.join(ROOT, "tests", "simt")
    subprocess.run([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-I", simt, os.path.join(simt, "selftest.cpp"), "-o", exe],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "selftest OK" in r.stdout
    # writes behind the dynamic shared memory of a launch / behind a device allocation are caught
    for mode, msg in (("oob-smem", "dynamic shared memory"), ("oob-global", "outside a 400-byte allocation")):
        r = subprocess.run([exe, mode], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and msg in r.stderr, (mode, r.returncode, r.stderr)


@pytest.mark.parametrize("nw", [1, 3, None], ids=["1warp", "3warps", "auto"])
@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("case", helpers.GOLDEN_CASES + helpers.GOLDEN_CASES_VP)
def test_interpreted_kernels_against_reference_golden(simt_lib, case, prec, nw):
    """golden vectors written by the compiled reference (tests/golden/make_golden.py); CTAs of 1 and 3 warps and of
    the automatic size (the largest that fits: 16 warps for the fp32 single-cluster kernel), two "SMs": several particles per warp, several CTAs per launch, the last-CTA epilogue"""
    wl, g = helpers.load_golden(case)
    ref = helpers.golden_stage(g, 4)
    with host.interpreted(sm_count=2, warps_per_cta=nw):
        so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=prec)
        _compare(wl, prec, ref, cnt, mean, cov, w, pw)
        mask, nfov = up.get_unused()
        ok = helpers.robust_mask(wl) if prec == 32 else np.ones(wl.N, bool)
        assert np.array_equal(mask[ok], ref["unused"][ok]) and np.array_equal(nfov[ok], ref["nfov"][ok])
        assert so.n_launches >= 1 and so.n_overflow == 0
        assert so.gm_total_in == int(wl.count.sum()) and so.gm_total_out == int(cnt.sum())
        up.normalize()
        wn = up.get_weights()
        assert wn.sum() == pytest.approx(1.0, abs=1e-12)
        assert np.allclose(wn, g["s5_weight"], rtol=1e-3 if prec == 32 else 1e-9, atol=1e-300)
        up.close()


def _param_sets(fn):
    """the cartesian product of a test function's pytest.mark.parametrize marks, as keyword dictionaries"""
    combos = [({}, "")]
    for m in [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]:
        names = [n.strip() for n in m.args[0].split(",")] if isinstance(m.args[0], str) else list(m.args[0])
        ids = m.kwargs.get("ids")
        nxt = []
        for c, cid in combos:
            for v in m.args[1]:
                vals = tuple(v.values) if isinstance(v, type(pytest.param(0))) else (tuple(v) if len(names) > 1 else (v,))
                d = dict(c)
                d.update(zip(names, vals))
                label = ids(vals[0]) if callable(ids) else "-".join(str(x) for x in vals)
                nxt.append((d, (cid + "-" if cid else "") + str(label)))
        combos = nxt
    return combos


def _gpu_tests(module_name, skip):
    """every `-m gpu` test of a module whose only fixture is cuda_required, one pytest.param per parametrisation"""
    mod = __import__(module_name)
    out = []
    for name, fn in sorted(vars(mod).items()):
        if not (name.startswith("test_") and callable(fn)) or any(s in name for s in skip):
            continue
        params = list(inspect.signature(fn).parameters)
        if not params or params[0] != "cuda_required":
            continue
        for kw, cid in _param_sets(fn):
            if set(params[1:]) == set(kw):
                out.append(pytest.param(fn, kw, id=f"{module_name}.{name}" + (f"[{cid}]" if cid else "")))
    return out


@pytest.mark.parametrize("fn,kw", _gpu_tests("test_gpu_parity", skip=("randomised",)) + _gpu_tests("test_gpu_vp", skip=()) +
                         _gpu_tests("test_gpu_fullsize", skip=("c4_every_shard",)) + _gpu_tests("test_motion", skip=()) +
                         _gpu_tests("test_gpu_births", skip=()) + _gpu_tests("test_gpu_multi", skip=("equals_nccl", "independent_of_the_number", "particle_exchange")))
def test_gpu_test_bodies_on_the_interpreted_kernels(simt_lib, fn, kw):
    """The `-m gpu` parity tests, bodies and sizes unchanged, with the binding pointed at the interpreter build: the
    reference's golden vectors and the oracle for both plugin sets and precisions, culled against exhaustive merge,
    multi-step sequences, NO_COMMIT / empty Z, pose-covariance modes, capacity overflow, API surface, matrix
    permanents, large partitions, the DP workspace, fused normalisation, update_host with and without copies, Victoria
    Park predict / births, the candidate-list births, the full-size C2 / C3 / C5 property tests (8 "SMs": every warp works through hundreds of
    particles), particle propagation."""
    with host.interpreted(sm_count=8 if "full_size" in fn.__name__ else 2):
        fn(simt_lib, **kw)


def test_smoke_body_on_the_interpreted_kernels(simt_lib, capsys):
    """__graft_entry__.smoke() — what the driver runs on cuda:0 before the bench — with the binding pointed at the
    interpreter build: both weightings of the 2-D model and the Victoria Park model against the oracle"""
    import __graft_entry__ as g
    with host.interpreted(sm_count=2):
        g.smoke()
    assert "smoke OK" in capsys.readouterr().out


def test_pageable_buffers_take_the_staged_path_with_the_same_results(simt_lib, monkeypatch):
    """rfsb200_update_host decides per call: buffers of rfsb200_host_alloc (known ranges) and other page-locked memory
    are read / written by the kernels, a pageable buffer anywhere sends the whole step through the staged copies"""
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater, pinned_array
    wl = synth.make_workload(N=40, nM=50, nZ=10, use_cluster_process=0, config_id=81)
    res = []
    with host.interpreted(sm_count=2):
        for pinned in (True, False):
            monkeypatch.setenv("SIMT_PAGEABLE", "0" if pinned else "1")   # what the interpreter's cudaPointerGetAttributes reports
            alloc = pinned_array if pinned else (lambda shape, dtype=np.float64: np.zeros(shape, dtype))
            up = PHDUpdater(wl.N, gm_capacity=128, z_capacity=16)
            up.set_model(wl.model); up.set_filter_cfg(wl.cfg); up.upload_maps(wl.count, wl.mean, wl.cov, wl.w)
            pose = alloc((wl.N, 3)); pose[:] = wl.pose
            w_in = alloc((wl.N,)); w_in[:] = wl.weight
            w_out, mask, nfov = alloc((wl.N,)), alloc((wl.N,), np.uint64), alloc((wl.N,), np.int32)
            so = up.update_host(pose, wl.pose_cov, w_in, np.ascontiguousarray(wl.Z), flags=capi.UPDATE_FUSED_ALLREDUCE,
                                w_out=w_out, unused_out=mask, nfov_out=nfov, want_stats=True)
            res.append((w_out.copy(), mask.copy(), nfov.copy(), so.sum_w, so.n_launches, up.download_maps()))
            up.close()
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert a[4] == 1 and b[4] == 2            # pinned: the update kernel alone; pageable: pose_convert + update
    for x, y in zip(a[5], b[5]):
        assert np.array_equal(x, y)


def test_interpreted_randomised_sweep():
    """tools/fuzz_parity.py --simt: random sizes / thresholds / models for both plugin sets and random CTA shapes,
    fp64 kernels against the oracle, exact structure"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "40", "777", "fp64", "--simt"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


_DROPIN_SCRIPT = """
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import test_gpu_dropin as t
for births in (None, t.CANDIDATES):
    for sc in (1, 0):
        for resample in (False, True):
            t.test_dropin_header_matches_reference_class(None, sc, resample, births)
print("dropin sequences OK")
"""


def test_dropin_header_against_the_reference_class_with_interpreted_kernels(simt_lib):
    """the body of tests/test_gpu_dropin.py — the drop-in C++ header against the reference class it replaces on scripted
    sequences (births in the direct and in the candidate-list form, landmark process noise, resampling with the same
    drand48 stream, an empty measurement set), both weightings, both precisions — in a fresh process whose rfsb200_* symbols are bound to the interpreter build"""
    from oracle import binding as ob
    if not ob.have_seq():
        pytest.skip("oracle/_ref/libseq_{ref,b200}.so not built (needs /root/reference at build time)")
    env = dict(os.environ, LD_PRELOAD=simt_lib._name, SIMT_SM_COUNT="2")
    r = subprocess.run([sys.executable, "-c", _DROPIN_SCRIPT.format(root=ROOT, tests=os.path.join(ROOT, "tests"))],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "dropin sequences OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_randomised_sequences_dropin_against_reference_class(simt_lib):
    """tools/fuzz_sequences.py: random scenarios (particle counts, lengths, landmark densities, both weightings, with and
    without resampling) through predict / births / update / resample of the drop-in header and of the reference class"""
    from oracle import binding as ob
    if not ob.have_seq():
        pytest.skip("oracle/_ref/libseq_{ref,b200}.so not built (needs /root/reference at build time)")
    env = dict(os.environ, LD_PRELOAD=simt_lib._name, SIMT_SM_COUNT="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_sequences.py"), "25", "5"], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_unchanged_simulator_on_the_dropin_header_with_interpreted_kernels(simt_lib, tmp_path, monkeypatch):
    """BASELINE config C1 end to end on the CPU: the body of tests/test_gpu_sim_c1.py (the reference's
    src/rbphdslam2dSim.cpp UNCHANGED, once on the reference's filter header and once on the drop-in header over the
    C ABI, 50 particles, the first 160 steps; identical particle poses, weights and best-particle maps with the fp64 kernels, same
    final error with the fp32 kernels) with the ABI symbols of the drop-in binary bound to the interpreter build."""
    import test_gpu_sim_c1 as c1
    if not os.path.exists(os.path.join(c1.REFDIR, "rbphdslam2dSim_b200")):
        pytest.skip("oracle/_ref/rbphdslam2dSim_{ref,b200} not built (needs /root/reference at build time)")
    # the first 160 of the 600 steps keep the CPU suite short (the -m gpu test runs all of them; poses follow the ground
    # truth for the first 100, src/rbphdslam2dSim.cpp:590-593)
    short = tmp_path / "refdir"
    short.mkdir()
    for f in ("rbphdslam2dSim_ref", "rbphdslam2dSim_b200"):
        os.symlink(os.path.join(c1.REFDIR, f), short / f)
    xml = open(os.path.join(c1.REFDIR, "rbphdslam2dSim.xml")).read()
    assert "<timesteps>600</timesteps>" in xml
    (short / "rbphdslam2dSim.xml").write_text(xml.replace("<timesteps>600</timesteps>", "<timesteps>160</timesteps>"))
    monkeypatch.setattr(c1, "REFDIR", str(short))
    monkeypatch.setenv("LD_PRELOAD", simt_lib._name)
    monkeypatch.setenv("SIMT_SM_COUNT", "2")
    c1.test_unchanged_simulator_runs_on_the_dropin(simt_lib, tmp_path)


def test_unchanged_victoria_park_driver_on_the_dropin_header_with_interpreted_kernels(simt_lib, tmp_path, monkeypatch):
    """BASELINE config C5 end to end on the CPU: the body of tests/test_gpu_sim_c5.py (the reference's
    src/rbphdslam_VictoriaPark.cpp UNCHANGED on both headers, head of the Victoria Park dataset, 100 particles:
    identical poses / weights / best-particle maps for the first 60 updates with the fp64 kernels, the same
    trajectory estimate afterwards and with the fp32 kernels)."""
    import test_gpu_sim_c5 as c5
    monkeypatch.setenv("LD_PRELOAD", simt_lib._name)
    monkeypatch.setenv("SIMT_SM_COUNT", "2")
    c5.test_unchanged_victoria_park_driver_runs_on_the_dropin(simt_lib, tmp_path)


@pytest.mark.parametrize("world", [2, 3])
def test_fused_cross_gpu_sum_between_interpreted_ranks(simt_lib, world):
    """RFSB200_UPDATE_FUSED_ALLREDUCE with more than one rank, on the CPU: every rank is a context driven by its own
    host thread (the interpreter's state is per thread), the "IPC handles" of the mailboxes are plain pointers, and the
    last CTA of each rank's update kernel stores its [sum w, sum w^2] into every rank's mailbox, waits for the others
    and normalises its shard — the exchange that otherwise only runs over NVLink peer memory.  Several steps, so both
    epoch parities of the double-buffered slots are used; the host-facing zero-copy step on top of it."""
    import threading
    from oracle import binding as ob
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater, pinned_array
    wl = synth.make_workload(N=41, nM=50, nZ=10, use_cluster_process=1, config_id=79)
    ref = ob.run(wl, sort_mode=ob.SORT_STABLE)
    wn_ref = ref.weight / ref.weight.sum()
    with host.interpreted(sm_count=2, warps_per_cta=2):
        shards = [wl.shard(r, world) for r in range(world)]
        ups = []
        for sh in shards:
            u = PHDUpdater(sh.N, gm_capacity=128, precision=64)
            u.load_workload(sh)
            ups.append(u)
        handles = [u.comm_export() for u in ups]
        for r, u in enumerate(ups):
            u.comm_connect(r, world, handles)
        w_host = [pinned_array((sh.N,)) for sh in shards]
        pose = [pinned_array((sh.N, 3)) for sh in shards]
        for r, sh in enumerate(shards):
            pose[r][:] = sh.pose
        for step in range(4):
            outs, errs = [None] * world, []

            def run(r):
                try:
                    f = capi.UPDATE_FUSED_ALLREDUCE | capi.UPDATE_NO_COMMIT
                    if step < 3:
                        outs[r] = ups[r].update(shards[r].Z, flags=f, want_stats=True)
                    else:   # the host-facing step: results stored into the pinned buffer by the same launch
                        outs[r] = ups[r].update_host(pose[r], shards[r].pose_cov, None, np.ascontiguousarray(shards[r].Z), flags=f,
                                                     w_out=w_host[r], want_stats=True)
                except Exception as e:   # noqa: BLE001
                    errs.append((r, e))
            ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
            for t in ts:
                t.start()
            for t in ts:
                t.join(timeout=120)
            assert not errs, errs
            assert not any(u.comm_error() for u in ups)
            # every rank holds the same global sums, bit for bit (added in rank order on every rank)
            assert len({(o.sum_w, o.sum_w2) for o in outs}) == 1
            assert outs[0].sum_w == pytest.approx(float(ref.weight.sum()), rel=1e-12)
            got = np.concatenate([u.get_weights(1) for u in ups])
            assert np.allclose(got, wn_ref, rtol=1e-10) and got.sum() == pytest.approx(1.0, abs=1e-12)
            if step == 3:
                assert np.array_equal(np.concatenate(w_host), got)
        for u in ups:
            u.close()


def _sharded_worker(rank, world, port, out_dir):
    """one rank of the N > 1 path on the CPU: the interpreted update kernel on this rank's block of particles
    (no normalisation), the ONE data-path collective — a SUM all-reduce of [sum w, sum w^2], here over gloo, in place on
    the buffer the kernel wrote — then the normalisation kernel; exactly the sequence of dist.ShardedUpdater.step()"""
    import ctypes as C
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import capi, dist as rd, synth
    from rfs_slam_b200.phd import PHDUpdater
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = synth.make_workload(N=45, nM=60, nZ=12, use_cluster_process=1, config_id=78)
    sh = wl.shard(rank, world)
    with host.interpreted(sm_count=1, warps_per_cta=3):
        up = PHDUpdater(sh.N, gm_capacity=128, precision=64)
        up.load_workload(sh)
        up.update(sh.Z, flags=capi.UPDATE_NO_NORMALIZE)
        buf = (C.c_double * 2).from_address(up.weight_sums_device_ptr())   # "device" memory is host memory here
        sums = torch.from_numpy(np.frombuffer(buf, dtype=np.float64))
        local = sums.clone()
        rd.allreduce_sums(sums)
        up.normalize()
        wn = up.get_weights()
        cnt, mean, cov, w = up.download_maps()
        up.close()
    lo, hi = rd.block_range(wl.N, rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, wn=wn, local=local.numpy(), sums=sums.numpy(), count=cnt, w=w)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_interpreted_update_with_gloo_allreduce(simt_lib, tmp_path, world):
    import socket
    import torch.multiprocessing as mp
    from oracle import binding as ob
    from rfs_slam_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_sharded_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    wl = synth.make_workload(N=45, nM=60, nZ=12, use_cluster_process=1, config_id=78)
    ref = ob.run(wl, sort_mode=ob.SORT_STABLE)
    got, cnt, covered, partial = np.zeros(wl.N), np.zeros(wl.N, dtype=np.int64), 0, np.zeros(2)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        assert lo == covered
        covered = hi
        got[lo:hi], cnt[lo:hi] = z["wn"], z["count"]
        partial += z["local"]
        assert np.allclose(z["sums"], [ref.weight.sum(), (ref.weight ** 2).sum()], rtol=1e-12)   # every rank: the global sums
    assert covered == wl.N and np.allclose(partial, [ref.weight.sum(), (ref.weight ** 2).sum()], rtol=1e-12)
    assert np.allclose(got, ref.weight / ref.weight.sum(), rtol=1e-10)
    assert got.sum() == pytest.approx(1.0, abs=1e-12)
    assert np.array_equal(cnt, ref.count)   # the maps do not depend on the sharding


def test_the_package_never_loads_the_interpreter_build():
    """the product path binds csrc/librfsb200.so only: no reference to tests/simt anywhere in the package or bench.py"""
    pkg = os.path.join(ROOT, "rfs-slam_b200")
    files = [os.path.join(pkg, f) for f in os.listdir(pkg) if f.endswith(".py")] + [os.path.join(ROOT, "bench.py")]
    for f in files:
        text = open(f).read()
        assert "simt" not in text.lower(), f
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import capi
    assert capi.LIB_PATH.endswith(os.path.join("csrc", "librfsb200.so"))
