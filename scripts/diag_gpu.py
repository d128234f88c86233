"""First-contact GPU diagnostics: per-block errors of each path vs the oracle + stage timings.
Writes gpurun_out/diag.json.  Run as: timeout 600 python scripts/diag_gpu.py"""
import faulthandler
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
faulthandler.enable()
faulthandler.dump_traceback_later(500, exit=True)

import __graft_entry__ as g  # noqa: E402

g.build_library()
from neurosync_trainer_lite_b200 import _native as nv  # noqa: E402
from neurosync_trainer_lite_b200 import engine, synth  # noqa: E402
from oracle import feature_oracle as fo  # noqa: E402

out = {}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def save():
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def blocks(got, want):
    d = np.abs(np.asarray(got, dtype=np.float64) - want)
    r = {"mfcc": float(d[:, :23].max()), "d1": float(d[:, 23:46].max()), "d2": float(d[:, 46:69].max())}
    if d.shape[1] > 69:
        r["ac"] = float(d[:, 69:].max())
    r["nan"] = int(np.isnan(got).sum())
    return r


def step(name, fn):
    t0 = time.time()
    try:
        out[name] = fn()
    except Exception as e:  # noqa: BLE001
        out[name] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
    out[name + "_s"] = round(time.time() - t0, 3)
    print(name, json.dumps(out[name])[:600], flush=True)
    save()


print("devices", nv.lib.nsf_device_count(), flush=True)
for sr, F, H, secs in [(88200, 1470, 735, 1.0), (16000, 266, 133, 2.0), (44100, 735, 367, 1.0)]:
    for kind in ("voiced", "gated"):
        y = synth.synth_clip(secs, sr, seed=1, kind=kind)
        want = fo.extract_and_combine_features(y, sr, F, H)
        eng = engine.get_engine(sr, F, H, device=0)
        step(f"simt_{sr}_{kind}", lambda: blocks(eng.extract_host(y, [0, len(y)], nv.DEBUG_SIMT_DFT), want))
        step(f"tc_{sr}_{kind}", lambda: blocks(eng.extract_host(y, [0, len(y)], 0), want))

# stage timings at C2 scale, device resident
import torch  # noqa: E402


def c2_times():
    eng = engine.get_engine(88200, 1470, 735, device=0)
    base = [synth.synth_clip(30.0, 88200, seed=s, kind="voiced") for s in range(4)]
    clips = [base[i % 4] for i in range(60)]
    packed, off = engine.pack_clips(clips)
    dev = torch.device("cuda", 0)
    pcm = torch.from_numpy(packed).to(dev)
    eng.set_profiling(True)
    res = {}
    ws = None
    outt = None
    for it in range(4):
        outt, ws = eng.extract_device(pcm, off, 0, out=outt, workspace=ws)
        torch.cuda.synchronize()
        res[f"iter{it}"] = eng.stage_times_ms()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.set_profiling(False)
    e0.record()
    for it in range(5):
        eng.extract_device(pcm, off, 0, out=outt, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res["ms_per_batch"] = ms
    res["audio_s_per_s"] = 1800.0 / (ms * 1e-3)
    # host path
    t0 = time.perf_counter()
    rows = eng.extract_host(packed, off)
    t1 = time.perf_counter()
    rows = eng.extract_host(packed, off)
    t2 = time.perf_counter()
    res["host_first_s"] = t1 - t0
    res["host_second_s"] = t2 - t1
    res["host_audio_s_per_s"] = 1800.0 / (t2 - t1)
    want = fo.extract_and_combine_features(base[0], 88200, 1470, 735)
    res["err"] = blocks(rows[:1801], want)
    return res


step("c2", c2_times)
save()
print("done", flush=True)
