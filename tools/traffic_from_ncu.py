"""Turn an `ncu --metrics ... --csv` log of ONE kernel launch into the small JSON bench.py reads for roofline.traffic.
usage: python tools/traffic_from_ncu.py <metrics.csv> <out.json> <blocks> <hash_log> <kernel label>"""
import csv, io, json, sys

def main(path, out, blocks, hash_log, label):
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    m = {}
    for r in csv.DictReader(io.StringIO(txt)):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "ms": 1e6, "us": 1e3, "ns": 1, "s": 1e9}.get(u, 1)
        m[r["Metric Name"]] = v * scale
        kern = r["Kernel Name"]
    rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    j = {"blocks": int(blocks), "kernel": label, "kernel_name": kern[:80],
         "dram_bytes_per_launch": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
         "duration_ns_under_ncu": int(m["gpu__time_duration.sum"]),
         "smsp_inst_executed": int(m["smsp__inst_executed.sum"]),
         "issue_active_pct": round(m["smsp__issue_active.avg.pct_of_peak_sustained_active"], 2),
         "l1_hit_pct": round(m["l1tex__t_sector_hit_rate.pct"], 2), "l2_hit_pct": round(m["lts__t_sector_hit_rate.pct"], 2),
         "warps_active_pct": round(m["sm__warps_active.avg.pct_of_peak_sustained_active"], 2),
         "source": f"ncu --metrics ... --clock-control none -k <kernel> -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu ({path})"}
    json.dump(j, open(out, "w"), indent=1)
    print(json.dumps(j))

if __name__ == "__main__":
    main(*sys.argv[1:6])
