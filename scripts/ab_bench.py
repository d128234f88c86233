#!/usr/bin/env python
"""A/B timing of experimental library builds on one GPU (device-resident C2 or C5 batch).

    python scripts/ab_bench.py [--workload c2] name=path[,ENV=VAL...] ...

Each variant runs in its own process (NSF_LIB_PATH selects the library, extra ENV=VAL pairs are
exported), prints the whole-step time (CUDA events, 10 steps after 3 warm-ups), the per-stage times
recorded by the library, and the max-abs error of clip 0 against the CPU oracle.  Experiments only.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(workload):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import bench
    from neurosync_trainer_lite_b200 import engine
    w = bench.WORKLOADS[workload]
    sr, Fr, Hr = w["sr"], w["F"], w["H"]
    eng = engine.get_engine(sr, Fr, Hr, device=0)
    packed, off, base, _mine, _n = bench.make_inputs(workload, 0, 1, "weak")
    dev = torch.device("cuda", 0)
    flags = 0
    if os.environ.get("AB_I16"):                     # int16 PCM + peak normalisation (the WAV path's front end)
        from neurosync_trainer_lite_b200 import _native as nv
        packed = np.clip(np.round(packed * 32767.0), -32768, 32767).astype(np.int16)
        flags = nv.PEAK_NORMALIZE
    pcm = torch.from_numpy(packed).to(dev)
    rows = int(eng.row_offsets(off)[-1])
    out = torch.empty((rows, 256), dtype=torch.float32, device=dev)
    ws = torch.empty(eng.workspace_bytes(len(packed), len(off) - 1), dtype=torch.uint8, device=dev)
    for _ in range(3):
        eng.extract_device(pcm, off, flags, out=out, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.extract_device(pcm, off, flags, out=out, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    eng.set_profiling(True)
    acc = {}
    for _ in range(5):
        eng.extract_device(pcm, off, flags, out=out, workspace=ws)
        torch.cuda.synchronize()
        for k, v in eng.stage_times_ms().items():
            acc[k] = acc.get(k, 0.0) + v / 5
    if flags:
        print(json.dumps({"ms": round(ms, 4), "stages": {k: round(v, 4) for k, v in acc.items() if v > 0},
                          "checksum": float(out[::997].abs().sum().item())}))
        return
    from oracle import feature_oracle as fo
    r1 = int(eng.row_offsets(off)[1])
    got = out[:r1].cpu().numpy()
    want = fo.extract_and_combine_features(base[0], sr, Fr, Hr)
    d = np.abs(got - want)
    print(json.dumps({"ms": round(ms, 4), "stages": {k: round(v, 4) for k, v in acc.items() if v > 0},
                      "mfcc_err": float(d[:, :23].max()), "delta_err": float(d[:, 23:69].max()),
                      "ac_err": float(d[:, 69:].max()), "checksum": float(out[::997].abs().sum().item())}))


def main():
    args = sys.argv[1:]
    workload = "c2"
    if args and args[0] == "--child":
        return child(args[1])
    if args and args[0] == "--workload":
        workload, args = args[1], args[2:]
    for spec in args:
        name, _, rest = spec.partition("=")
        parts = rest.split(",")
        env = dict(os.environ)
        if parts[0]:
            env["NSF_LIB_PATH"] = os.path.join(ROOT, parts[0])
        for kv in parts[1:]:
            k, _, v = kv.partition("=")
            env[k] = v
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", workload], env=env,
                             capture_output=True, text=True)
        line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-400:]
        print(f"{name:>12s} [{workload}] {line}", flush=True)


if __name__ == "__main__":
    main()
