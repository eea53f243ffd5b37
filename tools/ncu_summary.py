"""Summarise an .ncu-rep (raw page) into JSON: python tools/ncu_summary.py rep.ncu-rep [ray_steps_per_launch] > out.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_active_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_per_sm",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__t_sector_hit_rate.pct": "l1_sector_hit_pct", "lts__t_sector_hit_rate.pct": "l2_sector_hit_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_requests",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors_from_l1",
    "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum": "l2_read_sector_misses",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread", "launch__grid_size": "grid", "launch__block_size": "block",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "l1tex__data_pipe_lsu_wavefronts.sum": "l1_data_pipe_lsu_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_lsu_wavefronts_pct_of_peak",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed": "l1_lsu_writeback_active_pct",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum": "l1_tag_stage_load_wavefronts",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warp_latency_per_inst_issued.ratio": "warp_latency_per_instruction_cycles",
}


def to_base(value, unit):
    v = float(value.replace(",", ""))
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
    return v * mult.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS and vals[i] not in ("", "n/a"):
                try:
                    d[KEYS[h]] = to_base(vals[i], units[i])
                except ValueError:
                    pass
        if "dram_read" in d:
            d["dram_bytes_per_launch"] = d["dram_read"] + d.get("dram_write", 0.0)
        if d.get("l1_global_load_requests"):
            d["sectors_per_load_request"] = d["l1_global_load_sectors"] / d["l1_global_load_requests"]
        if steps:
            d["ray_steps"] = steps
            d["warp_instructions_per_warp_step"] = d.get("warp_instructions", 0) / (steps / 32.0)
            d["dram_bytes_per_ray_step"] = d.get("dram_bytes_per_launch", 0) / steps
            d["l2_to_l1_bytes_per_ray_step"] = d.get("l2_read_sectors_from_l1", 0) * 32.0 / steps
            d["l1_sector_bytes_per_ray_step"] = d.get("l1_global_load_sectors", 0) * 32.0 / steps
            d["g_ray_steps_per_s_under_ncu"] = steps / d["duration"] / 1e9
            if "l1_data_pipe_lsu_wavefronts" in d:
                d["l1_data_pipe_wavefronts_per_ray_step"] = d["l1_data_pipe_lsu_wavefronts"] / steps
        out.append(d)
    print(json.dumps(out if len(out) > 1 else out[0], indent=1))


if __name__ == "__main__":
    main()
