"""ncu raw pages -> profiles/<round>_kernel_counters.json.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > profiles/X_ncu_full.csv
    python tools/ncu_counters.py profiles/r02_kernel_counters.json k_step=profiles/A.csv k_net2=profiles/B.csv [note=...]

Per kernel: one list entry per captured launch.  bench.py reads `dram_bytes_per_launch` (k_step) for
roofline.traffic and `tensor_pipe_active_pct_of_elapsed` (k_net2) for nn_roofline."""
import csv
import json
import sys

METRICS = {
    "duration_us": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "tensor_pipe_active_pct_of_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_active_pct_of_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l2_to_sm_read": "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
    "warp_instructions": "smsp__inst_executed.sum",
    "stall_barrier_samples": "smsp__pcsamp_warps_issue_stalled_barrier",
    "stall_long_scoreboard_samples": "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "pc_samples": "smsp__pcsamp_sample_count",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = {}
    for key, name in METRICS.items():
        if name not in hdr:
            continue
        i = hdr.index(name)
        u = units[i]
        vals = []
        for r in data:
            try:
                vals.append(float(r[i].replace(",", "")) * SCALE.get(u, 1.0))
            except ValueError:
                pass
        out[key] = vals
        out[key + "_unit"] = "byte" if u.endswith("byte") else ("us" if u in ("us", "ms", "ns", "s") else u)
    if "dram_read" in out and "dram_write" in out:
        n = len(out["dram_read"])
        out["dram_bytes_per_launch"] = sum(a + b for a, b in zip(out["dram_read"], out["dram_write"])) / max(1, n)
    names = sorted({r[hdr.index("Kernel Name")] for r in data})
    out["kernel_names"] = names
    return out


def main():
    dst = sys.argv[1]
    doc = {"files": []}
    for arg in sys.argv[2:]:
        k, v = arg.split("=", 1)
        if k in ("note", "source"):
            doc[k] = v
        else:
            doc[k] = load(v)
            doc["files"].append(v)
    with open(dst, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps({k: (v.get("duration_us"), v.get("dram_bytes_per_launch")) for k, v in doc.items() if isinstance(v, dict)}))


if __name__ == "__main__":
    main()
