"""Dev tool: the two trajectory kernels (sg_traj_add_noise, sg_traj_channel_stats) timed alone with CUDA events on the
current stream, L2 flushed between repetitions, reported as achieved GB/s of ALGORITHMIC bytes against the measured HBM
roof of MEASURED_PEAKS.json (noise: read + write = 2*s per element, stats: s per element).
usage: dev_traj_bench.py [worlds=65536] [rows=200] [reps=10]   -> one JSON line per (kernel, dtype)"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
fn = importlib.import_module("soft-grip_b200.functions")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 7700.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2


CALLS = 4      # launches per timed region: the tensor (629 MB and up) is several times the 126 MB L2, so back-to-back passes over it
               # stream from HBM every time; the host-side cost of a call (ctypes, argument checks) overlaps the previous launch


def timed(f):
    best, tot = 1e9, 0.0
    for r in range(reps + 3):
        flush.fill_(r & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(CALLS):
            f()
        e1.record(); torch.cuda.synchronize()
        if r >= 3:
            ms = e0.elapsed_time(e1) / CALLS; best = min(best, ms); tot += ms
    return best, tot / reps


for dt in (torch.float32, torch.float64):
    s = 4 if dt == torch.float32 else 8
    x = torch.randn(W, T, 12, device="cuda", dtype=dt)
    out = torch.empty_like(x)
    mean, std = fn.channel_mean_std(x)
    ws, st_out = fn.stats_workspace(x), torch.empty((2, 12), dtype=torch.float64, device="cuda")     # allocated outside the timed region
    cases = {"noise": (lambda: fn.noised_modality(x, seed=1, out=out), 2 * s),
             "noise+standardise": (lambda: fn.noised_modality(x, seed=1, mean=mean, std=std, out=out), 2 * s),
             "stats": (lambda: fn.channel_mean_std(x, workspace=ws, out=st_out), s)}
    for name, (f, bpe) in cases.items():
        best, avg = timed(f)
        gb = x.numel() * bpe / 1e9
        print(json.dumps({"kernel": name, "dtype": str(dt).split(".")[-1], "elements": x.numel(), "algorithmic_GB": round(gb, 4),
                          "ms_best": round(best, 4), "ms_avg": round(avg, 4), "GBps_avg": round(gb / avg * 1e3, 1),
                          "hbm_peak_GBps": peak, "frac": round(gb / avg * 1e3 / peak, 4)}), flush=True)
    del x, out
