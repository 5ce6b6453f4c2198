"""Summarises ncu outputs brought back from the GPU box into small, committed files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python tools/ncu_summary.py full gpurun_out/prof_panel_r1.ncu-rep profiles/r1_panel_full.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_dmma.sum",
    "smsp__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
]


def short(name: str) -> str:
    m = re.search(r"(\w+_kernel)", name)
    base = m.group(1) if m else name[:40]
    t = re.search(r"<([^>]*)>", name)
    return base + (f"<{t.group(1)}>" if t else "")


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            val = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            rows.append((short(r["Kernel Name"]), val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): {len(rows)} launches, {tot / 1e3:.2f} ms of kernel time\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {us / 1e3:.3f} | {us / c:.1f} | {100 * us / tot:.1f}% |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines(True) if not l.startswith("==")]
    rd = list(csv.reader(io.StringIO("".join(lines))))
    if len(rd) < 3:
        print("no data in", src)
        return
    header, units = rd[0], rd[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n")
        for row in rd[2:]:
            d = dict(zip(header, row))
            u = dict(zip(header, units))
            f.write(f"## `{short(d.get('Kernel Name', '?'))}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {u.get(k, '')} |\n")
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
