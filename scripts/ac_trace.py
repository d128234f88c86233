#!/usr/bin/env python
"""Phase trace of the autocorrelation kernel (debug build: scripts/build_variant.sh actrace "" -DNSF_AC_TRACE).

    NSF_LIB_PATH=neurosync_trainer_lite_b200/_lib/variants/actrace.so python scripts/ac_trace.py

Every warp records the SM clock at the start of its staging half, the start and the end of its MMA loop for its first
40 frames.  Printed: mean staging / MMA-loop / cycle times, and per scheduler (warp % 4 of the blocks resident on one
SM) the share of time with k = 0..4 of its warps inside the MMA loop.  Experiments only."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from neurosync_trainer_lite_b200 import engine, _native
    w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    eng = engine.get_engine(w["sr"], w["F"], w["H"], device=0)
    packed, off, *_ = bench.make_inputs(sys.argv[1] if len(sys.argv) > 1 else "c2", 0, 1, "weak")
    pcm = torch.from_numpy(packed).cuda()
    rows = int(eng.row_offsets(off)[-1])
    out = torch.empty((rows, 256), dtype=torch.float32, device="cuda")
    ws = torch.empty(eng.workspace_bytes(len(packed), len(off) - 1), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        eng.extract_device(pcm, off, 0, out=out, workspace=ws)
    torch.cuda.synchronize()
    lib = ctypes.CDLL(_native.LIB_PATH)
    nf = 40
    tr = np.zeros((296, 8, nf, 3), dtype=np.int64)
    sm = np.zeros(296, dtype=np.int32)
    got = lib.nsf_debug_ac_trace(tr.ctypes.data_as(ctypes.c_void_p), sm.ctypes.data_as(ctypes.c_void_p))
    assert got == nf, got
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ac_trace.npz"), trace=tr, smid=sm)
    stage = (tr[..., 1] - tr[..., 0])[:, :, 4:]
    mma = (tr[..., 2] - tr[..., 1])[:, :, 4:]
    cyc = np.diff(tr[..., 0], axis=2)[:, :, 4:]
    res = {"staging_clk": float(stage.mean()), "mma_clk": float(mma.mean()), "cycle_clk": float(cyc.mean()),
           "staging_p10_p90": [float(np.percentile(stage, 10)), float(np.percentile(stage, 90))],
           "mma_p10_p90": [float(np.percentile(mma, 10)), float(np.percentile(mma, 90))]}
    # per SM and scheduler: occupancy of the MMA phase
    hist = np.zeros(5)
    by_sm = {}
    for b in range(296):
        by_sm.setdefault(int(sm[b]), []).append(b)
    res["blocks_per_sm"] = sorted(set(len(v) for v in by_sm.values()))
    res["block_pairs_sample"] = [by_sm[k] for k in sorted(by_sm)[:4]]
    for s, blocks in by_sm.items():
        for sched in range(4):
            ivs = []
            for b in blocks:
                for wp in (sched, sched + 4):
                    ivs.append(tr[b, wp, 4:36, 1:3])
            t0 = max(iv[0, 0] for iv in ivs)
            t1 = min(iv[-1, 1] for iv in ivs)
            if t1 <= t0:
                continue
            ev = []
            for iv in ivs:
                for a, e in iv:
                    a, e = max(a, t0), min(e, t1)
                    if e > a:
                        ev.append((a, 0, 1))
                        ev.append((e, 1, -1))
            ev.sort()
            k, last = 0, t0
            for tt, _, d in ev:
                hist[max(0, min(k, 4))] += tt - last
                last = tt
                k += d
    res["share_of_time_with_k_warps_in_mma"] = (hist / hist.sum()).round(4).tolist()
    res["mean_warps_in_mma"] = float((hist * np.arange(5)).sum() / hist.sum())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
