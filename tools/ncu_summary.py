"""Summarise an .ncu-rep (read on the CPU box): headline metrics + hottest SASS lines."""
import csv, subprocess, sys, io

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        res.append({h: (u, v) for h, u, v in zip(hdr, units, vals)})
    return res

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def main(rep, top=0.008):
    for d in raw(rep):
        print("=====", d.get("Kernel Name", ("", ""))[1][:110])
        for k in KEYS:
            if k in d:
                print(f"  {k:66s} {d[k][1]:>22s} {d[k][0]}")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
    smp = sum(int(r[ix["# Samples"]]) for r in data) or 1
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
    print("  total warp-inst", tot, " samples", smp)
    print("  stalls:", {k: round(100 * v / smp, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > smp * 0.01})
    for r in data:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
        if n / tot > top or s / smp > 0.012:
            print(f'  {r[ix["Address"]][-5:]} inst {n/tot*100:5.2f}% smp {s/smp*100:5.2f}% thr {r[ix["Avg. Threads Executed"]]:>4s}  {r[ix["Source"]][:84]}')

if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.008)
