"""Turns the ncu reports of one GPU visit (gpurun_out/*.ncu-rep, launches.csv, bench.json) into the tracked summaries under profiles/.
usage: python tools/make_profiles.py <tag>      e.g. s7 -> profiles/r01_*_s7.*"""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "s7"

METRICS = [
    ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid (CTAs)"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"), ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed_op_texture.sum", "warp TEX instructions"), ("l1tex__t_requests_pipe_tex_mem_texture.sum", "TEX quad requests"),
    ("l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum", "TEX wavefronts"), ("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "TEX wavefronts % of peak"),
    ("l1tex__t_sectors_pipe_tex_mem_texture.sum", "L1TEX texture sectors"), ("l1tex__t_sector_hit_rate.pct", "L1TEX hit rate %"), ("lts__t_sectors.sum", "L2 sectors"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall LG throttle / issue"),
]


def table(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return "(empty report)\n"
    h, units, data = rows[0], rows[1], rows[2:]
    ki = h.index("Kernel Name")
    names = [r[ki].split("(")[0].replace("void ", "")[:40] for r in data]
    s = "| metric | " + " | ".join(names) + " |\n|---|" + "---|" * len(names) + "\n"
    for m, label in METRICS:
        if m not in h:
            continue
        i = h.index(m)
        vals = []
        for r in data:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.3g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[i]}".strip())
        s += f"| {label} (`{m}`) | " + " | ".join(vals) + " |\n"
    return s


def main():
    os.makedirs(P, exist_ok=True)
    parts = [f"# Round 1, {tag} -- ncu `--set full --clock-control none --import-source on`, bench.py config 2 (CornellBox-Glossy 256^3, 1920x1080), one launch per kernel\n"]
    for rep, title, cmd in (("cone_full.ncu-rep", "Cone tracer", "-k regex:cone_kernel -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu"),
                            ("mip_full.ncu-rep", "Mip stage (running frame loop: sparse tile list)", "-k regex:mip_ -s 6 -c 2 python bench.py --steps 1 --warmup 3 --no-cpu"),
                            ("mip_dense_full.ncu-rep", "Mip stage, dense build (VCT_MIP_DENSE=1: every tile read and written)", "VCT_MIP_DENSE=1 ... -k regex:mip_ -s 6 -c 2 ..."),
                            ("small_full.ncu-rep", "Voxelizer, G-buffer, tile list, shade", "-k regex:'vox_|cam_|sparse_|fill_|tile_list|shade' ...")):
        path = os.path.join(O, rep)
        if os.path.exists(path):
            parts.append(f"\n## {title}\n\n`ncu --set full --clock-control none --import-source on {cmd}`\n\n" + table(path))
    open(os.path.join(P, f"r01_ncu_{tag}.md"), "w").write("".join(parts))
    for src, dst in (("launches.csv", f"r01_launches_{tag}.csv"), ("launch_summary.txt", f"r01_launch_summary_{tag}.txt"), ("bench.json", f"r01_bench_{tag}.json"),
                     ("quick_time.txt", f"r01_stage_times_{tag}.txt")):
        if os.path.exists(os.path.join(O, src)):
            shutil.copy(os.path.join(O, src), os.path.join(P, dst))
    print(open(os.path.join(P, f"r01_ncu_{tag}.md")).read()[:3000])


if __name__ == "__main__":
    main()
