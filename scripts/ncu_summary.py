"""Summarises an `ncu --set full` report into the per-kernel JSON kept under profiles/ (units normalised).
usage: python scripts/ncu_summary.py gpurun_out/<name>.ncu-rep profiles/<out>.json"""
import csv, json, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
WANT = {
    "gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_slots_busy_pct",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_memory_path_active_pct",
    "launch__registers_per_thread": "registers_per_thread", "launch__block_size": "block", "launch__grid_size": "grid",
    "launch__cluster_size": "cluster", "launch__shared_mem_per_block_dynamic": "dynamic_smem_bytes",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "")}
        for h, k in WANT.items():
            if h not in idx:
                continue
            try:
                v = float(r[idx[h]].replace(",", ""))
            except ValueError:
                continue
            d[k] = v * UNIT.get(units[idx[h]], 1.0)
        st = {}
        for h in hdr:
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[idx[h]].replace(",", "") or 0)
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        d["warp_stall_samples_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6]}
        if "dram_read_bytes" in d and "duration_us" in d:
            d["dram_GBps"] = (d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)) / d["duration_us"] / 1e3
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print("%-34s %8.1f us  dram %.3f+%.3f GB (%.0f GB/s, %.0f %%)  tensor %.0f %%  l2 %.0f %%  issue %.0f %%  %s" % (
            d["kernel"][:34], d.get("duration_us", 0), d.get("dram_read_bytes", 0) / 1e9, d.get("dram_write_bytes", 0) / 1e9,
            d.get("dram_GBps", 0), d.get("dram_throughput_pct", 0), d.get("tensor_pipe_active_pct", 0), d.get("l2_throughput_pct", 0),
            d.get("issue_slots_busy_pct", 0), d["warp_stall_samples_pct"]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
