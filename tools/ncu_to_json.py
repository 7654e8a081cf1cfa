"""Extracts the counters bench.py's `roofline` object needs from an `ncu --set full` capture of the cone kernel and stores them under
profiles/ (tracked): TEX wavefronts, DRAM bytes, warp instructions and the duration of ONE launch on the named workload.

    python tools/ncu_to_json.py gpurun_out/cone_full.ncu-rep config2_sampler1 [profiles/r02_cone_kernel_ncu.json]

bench.py reads the file, prints its path + hash in the JSON line and divides the wavefront count by the live CUDA-event time."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = {
    "tex_wavefronts": "l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum",
    "tex_wavefront_pct": "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "tex_requests": "l1tex__t_requests_pipe_tex_mem_texture.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "warp_instructions": "smsp__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex_sectors": "l1tex__t_sectors_pipe_tex_mem_texture.sum",
    "time_us": "gpu__time_duration.sum",
    "sm_cycles_elapsed_max": "sm__cycles_elapsed.max",
    "registers": "launch__registers_per_thread",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r02_cone_kernel_ncu.json")
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cone = [r for r in rows[2:] if "cone_kernel" in r[ki]]
    if not cone:
        sys.exit("no cone_kernel launch in " + rep)
    r = cone[0]
    rec = {"kernel": r[ki].split("(")[0].replace("void ", ""), "source": os.path.basename(rep), "ncu": "--set full --clock-control none (cold caches, serialised)"}
    for name, metric in FIELDS.items():
        if metric not in hdr:
            continue
        i = hdr.index(metric)
        v = float(r[i].replace(",", ""))
        rec[name] = v * SCALE.get(units[i], 1.0)
    data = {}
    if os.path.exists(out_path):
        data = json.load(open(out_path))
    data[key] = rec
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps({key: rec}, indent=1))


if __name__ == "__main__":
    main()
