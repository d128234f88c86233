#!/usr/bin/env python
"""Summarise an `ncu --set full` report of bench.py into the tables kept under profiles/.

    python scripts/ncu_extract.py gpurun_out/prof.ncu-rep --tag r01c [--workload c2]

Writes profiles/ncu_<tag>_summary.csv (one row per captured kernel launch: duration, DRAM bytes, DRAM /
tensor / issue / LSU utilisation, registers, top stall reasons) and profiles/ncu_traffic_<tag>.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by bench.py stage name) and prints the
markdown table.  Needs the `ncu` CLI (no GPU): it only reads the report.
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = {"k_tc_fold": "fold", "k_tc_stft_mel": "stft_gemm", "k_dct": "dct_stats",
         "k_delta_reduce": "cmvn_delta_reduce", "k_autocorr": "autocorr"}
COLS = {
    "time_ms": ("gpu__time_duration.sum", None),
    "dram_read_GB": ("dram__bytes_read.sum", None),
    "dram_write_GB": ("dram__bytes_write.sum", None),
    "dram_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "tensor_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    "issue_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "lsu_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    "smem_wavefront_pct": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
    "regs": ("launch__registers_per_thread", 1),
    "warps_per_sm": ("sm__warps_active.avg.per_cycle_active", 1),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME_MS = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}


def num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--workload", default="c2")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out, traffic = [], {"source": f"ncu --set full, capture {a.tag} ({os.path.basename(a.report)}), bytes per launch "
                                  "(dram__bytes_read.sum + dram__bytes_write.sum)", "workload": a.workload}
    for r in data:
        name = r[ix["Kernel Name"]]
        short = name.split("::")[-1].split("(")[0]
        rec = {"kernel": short}
        for col, (metric, scale) in COLS.items():
            v = num(r[ix[metric]]) if metric in ix else None
            if v is not None and metric.startswith("dram__bytes"):
                v = v * UNIT.get(units[ix[metric]], 1.0) / 1e9
            elif v is not None and metric == "gpu__time_duration.sum":
                v = v * TIME_MS.get(units[ix[metric]], 1.0)
            rec[col] = None if v is None else round(v, 4)
        stalls = {h.split("issue_stalled_")[1]: num(r[i]) or 0.0 for h, i in ix.items()
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h}
        tot = sum(stalls.values()) or 1.0
        rec["top_stalls"] = ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in
                                      sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        out.append(rec)
        for key, stage in STAGE.items():
            if key in short and rec["dram_read_GB"] is not None:
                traffic[stage] = round((rec["dram_read_GB"] + rec["dram_write_GB"]) * 1e9)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"ncu_{a.tag}_summary.csv"), "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(out[0].keys()))
        w.writeheader()
        w.writerows(out)
    with open(os.path.join(ROOT, "profiles", f"ncu_traffic_{a.tag}.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)
    keys = list(out[0].keys())
    print("| " + " | ".join(keys) + " |")
    print("|" + "---|" * len(keys))
    for rec in out:
        print("| " + " | ".join(str(rec[k]) for k in keys) + " |")


if __name__ == "__main__":
    main()
