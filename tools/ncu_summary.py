"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_softmin_partial.ncu-rep gpurun_out/launches.csv r01
    python tools/ncu_summary.py gpurun_out/prof_tc_fwd.ncu-rep - r01 tc_fwd        # any other kernel capture

writes profiles/<tag>_softmin_partial_ncu.json (+ profiles/softmin_partial_ncu_summary.json, read by
bench.py for roofline.traffic) and profiles/<tag>_launches_summary.json (per-kernel time shares).
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    name = sys.argv[4] if len(sys.argv) > 4 else "softmin_partial"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if os.path.exists(rep):
        kernels, units = raw_page(rep)
        k = kernels[0]
        summ = {"kernel": k.get("Kernel Name"), "report": os.path.basename(rep)}
        for key in KEEP:
            if key in k:
                try:
                    summ[key] = float(k[key])
                except ValueError:
                    summ[key] = k[key]
                summ[key + "__unit"] = units.get(key, "")
        stalls = {kk[len("smsp__pcsamp_warps_issue_stalled_"):]: float(v) for kk, v in k.items()
                  if kk.startswith("smsp__pcsamp_warps_issue_stalled_") and not kk.endswith("_not_issued") and v}
        tot = sum(stalls.values()) or 1.0
        summ["stall_pct"] = {s: round(100 * v / tot, 1) for s, v in sorted(stalls.items(), key=lambda t: -t[1])[:8]}

        def to_bytes(key):
            v = float(k[key])
            u = units.get(key, "byte").lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

        summ["dram_bytes_per_launch"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        json.dump(summ, open(os.path.join(ROOT, "profiles", f"{tag}_{name}_ncu.json"), "w"), indent=1)
        if name == "softmin_partial":  # read by bench.py for roofline.traffic
            json.dump(summ, open(os.path.join(ROOT, "profiles", "softmin_partial_ncu_summary.json"), "w"), indent=1)
        print(json.dumps(summ, indent=1))
    if os.path.exists(launches):
        lines = [l for l in open(launches) if not l.startswith("==")]
        rows = list(csv.DictReader(io.StringIO("".join(lines))))
        per = defaultdict(lambda: [0, 0.0])
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            u = r.get("Metric Unit", "ns")
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
            name = r["Kernel Name"].split("(")[0][:90]
            per[name][0] += 1
            per[name][1] += ns
        tot = sum(v[1] for v in per.values()) or 1.0
        out = [{"kernel": n, "launches": c, "total_ms": round(t / 1e6, 3), "share_pct": round(100 * t / tot, 3)}
               for n, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1])]
        json.dump({"command": "python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline (under ncu)",
                   "total_ms": round(tot / 1e6, 3), "kernels": out},
                  open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.json"), "w"), indent=1)
        for o in out[:8]:
            print(o)


if __name__ == "__main__":
    main()
