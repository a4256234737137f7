"""Turn an .ncu-rep (ncu --set full) into a small committed summary.

    python bench_tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_x_name.json [note]

Keeps the metrics the roofline discussion uses: duration, clocks, tensor-pipe
activity, DRAM bytes, L2/L1 throughput, shared-memory wavefronts, registers,
issue utilisation and the warp-stall histogram.
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")][:120]}
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if h.startswith("smsp__pcsamp"):
                    if v > 0:
                        d.setdefault("stall_samples", {})[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = v
                else:
                    d[h] = {"value": v, "unit": units[i]}
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            d["dram_traffic_bytes"] = rd["value"] * scale[rd["unit"]] + wr["value"] * scale[wr["unit"]]
        launches.append(d)
    json.dump({"source": rep, "note": note, "launches": launches}, open(out, "w"), indent=1)
    for d in launches:
        print(d["kernel"][:60], d.get("gpu__time_duration.sum"), d.get("dram_traffic_bytes"),
              d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    main()
