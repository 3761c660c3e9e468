import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import rfs_slam_b200
from rfs_slam_b200 import capi, synth
from rfs_slam_b200.phd import PHDUpdater, pinned_array
import bench, ctypes as C
wl, _ = bench.make_workload("C3", 0)
N = wl.N
up = PHDUpdater(N, gm_capacity=256, z_capacity=32)
up.load_workload(wl)
h_pose = pinned_array((N, 3)); h_pose[:] = wl.pose
h_w = pinned_array((N,)); h_w[:] = wl.weight
h_wout = pinned_array((N,)); h_mask = pinned_array((N,), np.uint64); h_nfov = pinned_array((N,), np.int32)
Zc = np.ascontiguousarray(wl.Z, dtype=np.float64)
FL = capi.UPDATE_NO_COMMIT | capi.UPDATE_FUSED_ALLREDUCE
for _ in range(5): up.update_host(h_pose, wl.pose_cov, h_w, Zc, flags=FL, w_out=h_wout, unused_out=h_mask, nfov_out=h_nfov)
K = 200
t0 = time.perf_counter()
for _ in range(K): up.update_host(h_pose, wl.pose_cov, h_w, Zc, flags=FL, w_out=h_wout, unused_out=h_mask, nfov_out=h_nfov)
t1 = time.perf_counter()
print("python update_host, L2 warm, back to back: %.1f us" % ((t1 - t0) / K * 1e6))
# raw ctypes call with prebuilt args
lib = up.lib
p = up._host_ptr
args = (up.ctx, p(h_pose), p(wl.pose_cov), 1, p(h_w), p(Zc), Zc.shape[0], FL, p(h_wout), p(h_mask), p(h_nfov), None)
t0 = time.perf_counter()
for _ in range(K): lib.rfsb200_update_host(*args)
t1 = time.perf_counter()
print("raw ctypes call: %.1f us" % ((t1 - t0) / K * 1e6))
for _ in range(3): so = up.update(wl.Z, flags=FL)
so = up.update(wl.Z, flags=FL, want_stats=True)
print("device-resident update elapsed_us (events): %.1f" % so.elapsed_us)
t0 = time.perf_counter()
for _ in range(K): up.update(wl.Z, flags=FL)
up.synchronize()
t1 = time.perf_counter()
print("async update back to back: %.1f us" % ((t1 - t0) / K * 1e6))
up.profile_begin(50)
for _ in range(50): up.update_host(h_pose, wl.pose_cov, h_w, Zc, flags=FL, w_out=h_wout, unused_out=h_mask, nfov_out=h_nfov)
k = up.profile_read()
print("kernel alone inside update_host (events): mean %.1f min %.1f us" % (k.mean(), k.min()))
up.profile_begin(50)
for _ in range(50): up.update(wl.Z, flags=FL, want_stats=False)
up.synchronize()
k = up.profile_read()
print("kernel alone, device-resident (events): mean %.1f min %.1f us" % (k.mean(), k.min()))
t0 = time.perf_counter()
for _ in range(K): up.update(wl.Z, flags=FL, want_stats=False)
up.synchronize()
t1 = time.perf_counter()
print("async update back to back: %.1f us" % ((t1 - t0) / K * 1e6))
# host step without the optional outputs
t0 = time.perf_counter()
for _ in range(K): up.update_host(h_pose, wl.pose_cov, h_w, Zc, flags=FL, w_out=h_wout)
t1 = time.perf_counter()
print("update_host, weights out only: %.1f us" % ((t1 - t0) / K * 1e6))
def timeit(label, fn):
    for _ in range(5): fn()
    up.profile_begin(50)
    t0 = time.perf_counter()
    for _ in range(50): fn()
    t1 = time.perf_counter()
    k = up.profile_read()
    print("%-50s host %.1f us  kernel %.1f us" % (label, (t1 - t0) / 50 * 1e6, k.mean()))
timeit("update_host no outputs (fused)", lambda: up.update_host(h_pose, wl.pose_cov, h_w, Zc, flags=FL))
timeit("update_host no outputs, no weights in", lambda: up.update_host(h_pose, wl.pose_cov, None, Zc, flags=FL))
timeit("update_host all outputs (fused)", lambda: up.update_host(h_pose, wl.pose_cov, h_w, Zc, flags=FL, w_out=h_wout, unused_out=h_mask, nfov_out=h_nfov))
timeit("update_host all outputs, NO_NORMALIZE", lambda: up.update_host(h_pose, wl.pose_cov, h_w, Zc, flags=capi.UPDATE_NO_COMMIT | capi.UPDATE_NO_NORMALIZE, w_out=h_wout, unused_out=h_mask, nfov_out=h_nfov))
timeit("update_host all outputs, normalize kernel (sync)", lambda: up.update_host(h_pose, wl.pose_cov, h_w, Zc, flags=capi.UPDATE_NO_COMMIT, w_out=h_wout, unused_out=h_mask, nfov_out=h_nfov))
timeit("update (resident) + sync", lambda: (up.update(wl.Z, flags=FL, want_stats=False), up.synchronize()))
