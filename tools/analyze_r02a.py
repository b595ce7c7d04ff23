"""Read what tools/r02_first_call.sh left under gpurun_out/r02a/ and print the round's first numbers side by side:
GPU test tail, both bench arms, the A/B of the experimental compressor variants (speed lines from tools/probe.py and the ncu
instruction counts per 64 KiB block), the second HC design, the e2e chunk sweep.
usage: python tools/analyze_r02a.py [gpurun_out/r02a]"""
import csv
import io
import json
import os
import sys


def ncu_metrics(path):
    try:
        txt = open(path).read()
        txt = txt[txt.index('"ID"'):]
    except (OSError, ValueError):
        return None
    m = {}
    for r in csv.DictReader(io.StringIO(txt)):
        try:
            m[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"ms": 1e6, "us": 1e3, "ns": 1, "s": 1e9}.get(r["Metric Unit"], 1)
        except ValueError:
            pass
    return m


def tail(path, n=3):
    try:
        return [l.rstrip()[:300] for l in open(path).read().splitlines()[-n:]]
    except OSError:
        return ["(missing)"]


def main(d):
    print("== GPU tests"); print("\n".join(tail(os.path.join(d, "gpu_tests.log"), 2)))
    print("== GPU tests, experimental HC"); print("\n".join(tail(os.path.join(d, "gpu_tests_experimental.log"), 1)))
    for name in ("bench_reference_arm.json", "bench_full.json"):
        try:
            j = json.loads(open(os.path.join(d, name)).read().strip().splitlines()[-1])
            keep = {k: j.get(k) for k in ("impl", "value", "compress_gibs", "decompress_gibs", "ratio", "ms_per_step", "gpu_launches")}
            keep["e2e"] = (j.get("e2e") or {}).get("value"); keep["cpu"] = (j.get("cpu_baseline") or {}).get("value")
            keep["roofline_frac"] = (j.get("roofline") or {}).get("frac"); keep["clocks"] = j.get("clocks")
            print("==", name, json.dumps(keep))
        except Exception as e:      # noqa: BLE001
            print("==", name, "unreadable:", e, tail(os.path.join(d, name.replace(".json", ".err")), 3))
    print("== compressor variants (probe lines)"); print("\n".join(tail(os.path.join(d, "runs_ab.log"), 60)))
    print("== compressor variants (ncu, 8192 x 64 KiB blocks)")
    for so in ("libb200lz4", "libb200lz4_runs", "libb200lz4_split"):
        m = ncu_metrics(os.path.join(d, f"runs_ab_{so}.csv"))
        if not m:
            print(f"  {so}: (missing)"); continue
        inst = m.get("smsp__inst_executed.sum", 0) / 8192
        print(f"  {so:18s} {m.get('gpu__time_duration.sum', 0) / 1e6:8.3f} ms   {inst / 1e3:7.1f} K warp-instructions / block   "
              f"issue active {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} %")
    print("== secondary configs (incl. second HC design)"); print("\n".join(tail(os.path.join(d, "secondary_configs.log"), 14)))
    print("== e2e chunk sweep"); print("\n".join(tail(os.path.join(d, "e2e_chunk_sweep.log"), 9)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02a")
