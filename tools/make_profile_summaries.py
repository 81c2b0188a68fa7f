"""Turn what a GPU run left in gpurun_out/ into the committed summaries under profiles/.

usage: python tools/make_profile_summaries.py [--tag r01]
reads (whatever exists):
  gpurun_out/bench_<tag>.json, bench_<tag>_ref.json, bench_2gpu.json     -> profiles/<tag>_bench*.json
  gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum)   -> profiles/<tag>_launches_bench_step.csv,
                                                                            profiles/<tag>_launch_shares.csv
  gpurun_out/prof_{jac,sqr,lq}_<tag>.ncu-rep (ncu --set full)            -> profiles/<tag>_k_*_ncu_summary.csv
  gpurun_out/free_run.log, parity_report.txt                             -> profiles/<tag>_free_run_..., <tag>_teacher_forced_parity.txt
"""
import argparse
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

NCU_KEEP = [
    "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def copy(src, dst):
    if os.path.isfile(os.path.join(OUT, src)):
        shutil.copyfile(os.path.join(OUT, src), os.path.join(PROF, dst))
        print("wrote", dst)


def launches(tag):
    src = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.isfile(src):
        return
    shutil.copyfile(src, os.path.join(PROF, f"{tag}_launches_bench_step.csv"))
    lines = [ln for ln in open(src) if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = {}
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) * 1e-6
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launch_shares.csv"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline` "
                f"(one dQA step of 888 realisations, N=21, chi=32, skipping the warm-up launches)\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's "
                "roofline.families[*].share_of_step\n")
        f.write("kernel,launches,total_ms,share\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name},{n},{ms:.3f},{ms / tot:.3f}\n")
    print("wrote launch shares;", len(rows), "launches,", f"{tot:.1f} ms")


def ncu_summary(rep, kernel, tag):
    path = os.path.join(OUT, rep)
    if not os.path.isfile(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(os.path.join(PROF, f"{tag}_{kernel}_ncu_summary.csv"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none, {vals[hdr.index('Kernel Name')]}, 1 mid-chain launch, N=21, chi=32, plan of 888 "
                f"realisations (tools/profile_step.py --chi 32 --batch 888; a grid of 296 instances = one of the three sub-batches the "
                f"step graph launches, 888 = a whole-batch launch of the profiled step)\n")
        f.write("metric,unit,value\n")
        for key in NCU_KEEP:
            if key in hdr:
                i = hdr.index(key)
                f.write(f"{key},{units[i]},{vals[i].replace(',', '') if key not in ('Block Size', 'Grid Size') else vals[i]}\n")
    print("wrote", f"{tag}_{kernel}_ncu_summary.csv")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    a = ap.parse_args()
    t = a.tag
    copy(f"bench_{t}.json", f"{t}_bench.json")
    copy(f"bench_{t}_ref.json", f"{t}_bench_reference_arm.json")
    copy("bench_2gpu.json", f"{t}_bench_2gpu.json")
    for n in (2, 4, 8):  # round 2 names
        copy(f"{t}_bench_{n}gpu.json", f"{t}_bench_{n}gpu.json")
        copy(f"{t}_c4_{n}gpu.json", f"{t}_c4_{n}gpu.json")
    copy(f"{t}_c4_1gpu.json", f"{t}_c4_1gpu.json")
    copy(f"{t}_c4_eps_mean_8gpu.npz", f"{t}_c4_eps_mean.npz")
    copy(f"{t}_c5.json", f"{t}_c5.json")
    copy(f"{t}_c5_batch74.json", f"{t}_c5_batch74.json")
    copy(f"{t}_c2_full_runs.jsonl", f"{t}_c2_full_runs.jsonl")
    copy(f"{t}_fp64_peaks.json", f"{t}_fp64_peaks.json")
    copy(f"chi_sweep_{t}.jsonl", f"{t}_chi_sweep.jsonl")
    copy(f"pytest_gpu_{t}.log", f"{t}_pytest_gpu.log")
    copy("free_run.log", f"{t}_free_run_vs_reference_envelope.txt")
    copy("parity_report.txt", f"{t}_teacher_forced_parity.txt")
    launches(t)
    ncu_summary(f"prof_jac_{t}.ncu-rep", "k_jacobi", t)
    ncu_summary(f"prof_sqr_{t}.ncu-rep", "k_sector_qr", t)
    ncu_summary(f"prof_lq_{t}.ncu-rep", "k_lq", t)
    ncu_summary(f"prof_energy_{t}.ncu-rep", "k_energy", t)
    ncu_summary(f"prof_theta_{t}.ncu-rep", "k_theta", t)


if __name__ == "__main__":
    sys.exit(main())
