"""Turns an .ncu-rep (ncu --set full) into the small JSON summary kept under profiles/ (the .ncu-rep itself stays in
gpurun_out/, which is scratch).    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02/x_summary.json [note]"""
import csv, io, json, subprocess, sys

KEEP = {
    "gpu__time_duration.sum": "gpu_time", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct_of_peak",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct_of_peak",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_of_peak",
    "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct", "launch__registers_per_thread": "registers_per_thread",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct", "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "launch__grid_size": "grid_size", "launch__block_size": "block_size", "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks", "sm__cycles_active.avg": "sm_cycles_active_avg",
    "smsp__cycles_active.avg": "smsp_cycles_active_avg", "gpc__cycles_elapsed.max": "gpc_cycles_elapsed_max",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEEP:
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    v = vals[i]
                d[KEEP[h]] = v
                if units[i]:
                    d[KEEP[h] + "_unit"] = units[i]
        launches.append(d)
    res = {"source": rep, "how": "ncu --set full --clock-control none --import-source on, one launch, read here with ncu -i ... --page raw --csv", "note": note,
           "launches": launches}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(launches[0])[:400])


if __name__ == "__main__":
    main()
