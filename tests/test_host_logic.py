"""Host-side logic that needs no GPU: synthetic workloads, sharding, descriptors."""
import os

import numpy as np

import rfs_slam_b200  # noqa: F401
from rfs_slam_b200 import capi, synth


def test_workload_is_deterministic_and_shaped():
    a = synth.make_workload(N=32, nM=50, nZ=10, config_id=1)
    b = synth.make_workload(N=32, nM=50, nZ=10, config_id=1)
    assert np.array_equal(a.mean, b.mean) and np.array_equal(a.Z, b.Z)
    assert a.count.sum() == a.mean.shape[0] == a.cov.shape[0] == a.w.shape[0] == 32 * 50
    assert a.Z.shape == (10, 2)
    # covariances are positive definite
    det = a.cov[:, 0] * a.cov[:, 2] - a.cov[:, 1] ** 2
    assert (det > 0).all() and (a.cov[:, 0] > 0).all()
    # clutter integral follows MeasurementModel_RngBrg::clutterIntensityIntegral
    assert a.model["clutter_integral"] == a.model["clutter_intensity"] * 2 * np.arccos(-1) * (10.0 - 0.5)


def test_parity_extras_cover_the_quirk_regions():
    wl = synth.make_workload(N=200, nM=100, nZ=20, config_id=2, parity_extras=True)
    assert (wl.count == 0).any()                      # Q10: empty maps
    r = np.hypot(wl.landmarks[:, 0], wl.landmarks[:, 1])
    md = wl.model
    assert ((r > md["range_max"]) & (r < md["range_max"] + md["range_buffer"])).any()   # Q2 outside band
    assert ((r < md["range_min"]) & (r > md["range_min"] - md["range_buffer"])).any()
    b = np.arctan2(wl.landmarks[:, 1], wl.landmarks[:, 0])
    assert (np.abs(np.abs(b) - np.pi) < 0.02).sum() >= 2                                  # Q3 bearing wrap


def test_shards_partition_the_particles():
    wl = synth.make_workload(N=37, nM=20, nZ=6, config_id=3, ragged=0.3)
    parts = [wl.shard(r, 4) for r in range(4)]
    assert sum(p.N for p in parts) == wl.N
    assert np.array_equal(np.concatenate([p.count for p in parts]), wl.count)
    assert np.array_equal(np.concatenate([p.mean for p in parts]), wl.mean)
    assert np.array_equal(np.concatenate([p.pose for p in parts]), wl.pose)
    for p in parts:
        assert np.array_equal(p.Z, wl.Z)


def test_descriptor_marshalling():
    wl = synth.make_workload(N=4, nM=5, nZ=3, config_id=4)
    d = capi.model_desc(wl.model)
    assert d.model_id == 1 and d.R[0] == 5e-3 and d.R[3] == 5e-4 and d.R[1] == 0.0
    assert d.range_max == 10.0 and d.innov_thr_bearing == 0.2
    c = capi.filter_cfg(wl.cfg)
    assert c.eval_point_count == 15 and c.use_cluster_process == 1 and c.pruning_threshold == 0.01


def test_replay_dump_layout_roundtrip(tmp_path):
    """tools/replay_dump.py reads what the drop-in header writes with RFSB200_DUMP_UPDATE=k (layout in its docstring)."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import replay_dump
    from rfs_slam_b200 import capi, synth
    for wl in (synth.make_workload(N=5, nM=7, nZ=4, config_id=401), synth.make_vp_workload(N=5, nM=14, nZ=4, config_id=402)):
        D = wl.dim
        md, fc = capi.model_desc(wl.model), capi.filter_cfg(wl.cfg)
        pc = np.zeros((wl.N, 6)) if wl.pose_cov is None else np.tile(wl.pose_cov, (wl.N, 1))
        scan = np.asarray(wl.model.get("scan", []), dtype=np.float64)
        path = tmp_path / f"dump{D}.bin"
        with open(path, "wb") as f:
            f.write(np.array([0x52465342, wl.N, D, wl.nZ, C.sizeof(md), C.sizeof(fc)], np.int32).tobytes())
            f.write(wl.count.astype(np.int32).tobytes())
            f.write(np.array([int(wl.count.sum())], np.int64).tobytes())
            for a in (wl.mean, wl.cov, wl.w, wl.pose, pc, wl.weight, wl.Z):
                f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
            f.write(bytes(md))
            f.write(np.array([len(scan)], np.int32).tobytes())
            f.write(scan.tobytes())
            f.write(bytes(fc))
        got = replay_dump.load(str(path))
        assert got.dim == D and got.N == wl.N and got.nZ == wl.nZ
        for a, b in ((got.count, wl.count), (got.mean, wl.mean), (got.cov, wl.cov), (got.w, wl.w), (got.pose, wl.pose), (got.Z, wl.Z)):
            assert np.array_equal(a, b)
        assert got.model["model_id"] == wl.model["model_id"] and got.cfg["eval_point_count"] == wl.cfg["eval_point_count"]
        if D == 3:
            assert got.model["scan"] == wl.model["scan"] and got.model["pd_table"] == list(wl.model["pd_table"])
            # and the oracle gives the same answer on the reloaded workload
            from oracle import binding as ob
            a, b = ob.run(wl), ob.run(got)
            assert np.array_equal(a.count, b.count) and np.allclose(a.weight, b.weight, rtol=1e-12)
