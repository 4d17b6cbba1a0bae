"""Condense an Nsight Compute report into the few counters the design argues from.

  ncu -i REPORT.ncu-rep --page raw --csv > raw.csv     (works without a GPU)
  python tools/summarize_ncu.py raw.csv out.json

Per captured launch: duration, DRAM bytes / throughput, FP64 and DMMA pipe
utilisation, achieved occupancy, registers, shared-memory wavefronts."""
import csv, json, sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_cycles_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
}


def main(raw, out):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "")}
        for k, name in KEYS.items():
            if k in col and r[col[k]] not in ("", "n/a"):
                try:
                    d[name] = float(r[col[k]].replace(",", ""))
                    d[name + "_unit"] = units[col[k]]
                except ValueError:
                    pass
        if "dram_read" in d and "duration" in d:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
            rd = d["dram_read"] * scale.get(d["dram_read_unit"], 1)
            wr = d.get("dram_write", 0.0) * scale.get(d.get("dram_write_unit", "byte"), 1)
            sec = d["duration"] * tscale.get(d["duration_unit"], 1)
            d["dram_bytes"] = rd + wr
            d["dram_gbs"] = (rd + wr) / sec / 1e9
            d["duration_us"] = sec * 1e6
        launches.append({k: v for k, v in d.items() if not k.endswith("_unit")})
    json.dump(launches, open(out, "w"), indent=1)
    for d in launches:
        print(d["kernel"][:60].ljust(60), {k: (round(v, 2) if isinstance(v, float) else v)
                                           for k, v in d.items() if k not in ("kernel",)})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
