"""Summaries for profiles/ from the round's ncu captures (gpurun_out/prof_<tag>.ncu-rep): key metrics, stage shares, and
profiles/traffic.json (DRAM bytes / warp instructions per launch, tagged with the fingerprint of the kernel sources so
that bench.py never reports them for another build).   usage: python tools/ncu_summary.py r2"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
traffic = {"_note": "per launch, from `ncu --set full --clock-control none` captures of tools/r2_profiles.sh (cold-cache, under the profiler); "
                    "bench.py reports them only while the kernel sources still have this fingerprint",
           }
fp = bench._kernel_fingerprint()
for tag, cfg, npart in (("C3", "C3", 8000), ("C3mf", "C3mf", 8000), ("C5", "C5", 4000), ("C2", "C2", 1000)):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{rnd}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, v = rows[0], rows[1], rows[-1]
    out = [f"ncu --set full --clock-control none --import-source on -k regex:phd_update -s 2 -c 1  python tools/profile_step.py ... ({cfg})",
           "kernel: " + v[h.index("Kernel Name")]]
    vals = {}
    for w in WANT:
        if w in h:
            i = h.index(w)
            out.append("%-90s %-16s %s" % (w, units[i], v[i]))
            try:
                vals[w] = float(v[i]) * UNIT_SCALE.get(units[i], 1.0)
            except ValueError:
                pass
    open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_{tag}_metrics.txt"), "w").write("\n".join(out) + "\n")
    dump = os.path.join("/tmp", f"src_{rnd}_{tag}.csv")
    with open(dump, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=f, stderr=subprocess.DEVNULL)
    if tag != "C5":
        st = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stages.py"), dump, str(npart)], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_{tag}_stages.txt"), "w").write(
            f"stage shares of the update kernel on {cfg} (ncu source page: warp instructions executed per particle / stall samples)\n" + st)
    traffic[cfg] = dict(dram_bytes=int(vals.get("dram__bytes_read.sum", 0) + vals.get("dram__bytes_write.sum", 0)),
                        warp_instructions=int(vals.get("smsp__inst_executed.sum", 0)), kernel_fingerprint=fp,
                        capture=f"profiles/{rnd}_ncu_{tag}_metrics.txt")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
