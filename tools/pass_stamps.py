"""Phase stamps (%globaltimer) of the three per-pass kernels of one fit iteration, from a -DFGP_TIMING build:
    FGP_LIB_DIR=$PWD/fastgaussianprocesses_b200/lib_rt FGP_BUILD_DEFS=-DFGP_TIMING python -m fastgaussianprocesses_b200.build
    FGP_B200_LIB=$PWD/fastgaussianprocesses_b200/lib_rt/libfgp_b200.so python tools/pass_stamps.py [log2n] [d] [lattice|net] [cold]
"cold": 256 MiB are written before every stamped iteration (the L2 flush of bench.py's `value`).
Prints, in microseconds since the first pass-A CTA started: per stamp the min / mean / max over CTAs, and the per-phase durations."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp
from fastgaussianprocesses_b200 import _lib as L
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
fam = sys.argv[3] if len(sys.argv) > 3 else "lattice"
cold = len(sys.argv) > 4 and sys.argv[4] == "cold"
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev) if cold else None
gp = fgp.FastGPLattice(fgp.Lattice(d, seed=7), device=dev) if fam == "lattice" else fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, seed=7), device=dev)
x = gp.get_x_next(1 << log2n)
gp.add_y_next(torch.cos(2 * np.pi * x).sum(1))
st = gp.fit_stepper()
for _ in range(20):
    st.step()
torch.cuda.synchronize()
fn = ctypes.CDLL(L.LIB_PATH).fgp_debug_stamps
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
names = {16: "A entry", 17: "A prologue done", 18: "A k1 in smem", 19: "A block fft done", 20: "A exit",
         21: "B entry", 22: "B loaded + column fft", 23: "B spectral done", 24: "B exit",
         25: "C entry", 26: "C prologue done", 27: "C packed in smem", 28: "C inverse fft done", 29: "C contraction done", 30: "C partials stored", 31: "C exit (fit tail)"}
runs = []
for rep in range(5):
    if cold:
        flush.zero_()
        torch.cuda.synchronize()
    st.step()
    torch.cuda.synchronize()
    buf = np.zeros((1024, 48), dtype=np.int64)
    fn(buf.ctypes.data, 1024)
    live = buf[buf[:, 16] != 0]
    t0 = live[:, 16].min()
    runs.append((live[:, 16:32] - t0) * 1e-3)
    tail = buf[buf[:, 34] > t0]  # the CTA that ran the fit step of THIS iteration
    if len(tail):
        tl = tail[np.argmax(tail[:, 34])]
        tail_us = {"ticket won": (tl[32] - t0) * 1e-3, "partials reduced": (tl[33] - t0) * 1e-3, "fit step done": (tl[34] - t0) * 1e-3, "that CTA's partials stored": (tl[30] - t0) * 1e-3}
rel = runs[-1]
out = {"family": fam, "log2n": log2n, "d": d, "l2": "flushed" if cold else "warm", "ctas": int(rel.shape[0]), "stamps_us": {}}
for k in range(16, 32):
    col = rel[:, k - 16]
    col = col[col > -1e6]
    out["stamps_us"][names[k]] = {"min": round(float(col.min()), 2), "mean": round(float(col.mean()), 2), "max": round(float(col.max()), 2)}
dur = lambda a, b: rel[:, b - 16] - rel[:, a - 16]
ph = {"A prologue": dur(16, 17), "A eval": dur(17, 18), "A fft": dur(18, 19), "A unpack+store": dur(19, 20),
      "B load+fft": dur(21, 22), "B spectral": dur(22, 23), "B ifft+store": dur(23, 24),
      "C prologue": dur(25, 26), "C load+pack": dur(26, 27), "C ifft": dur(27, 28), "C contraction": dur(28, 29), "C reduce+store": dur(29, 30), "C tail": dur(30, 31)}
out["phase_us"] = {k: {"min": round(float(v.min()), 2), "mean": round(float(v.mean()), 2), "max": round(float(v.max()), 2)} for k, v in ph.items()}
out["kernel_span_us"] = {"A": [round(float(rel[:, 0].min()), 2), round(float(rel[:, 4].max()), 2)], "B": [round(float(rel[:, 5].min()), 2), round(float(rel[:, 8].max()), 2)],
                         "C": [round(float(rel[:, 9].min()), 2), round(float(rel[:, 15].max()), 2)]}
out["fit_tail_us"] = {k: round(float(v), 2) for k, v in tail_us.items()}
out["iteration_span_us_5runs"] = [round(float(r[:, 15].max()), 2) for r in runs]
print(json.dumps(out, indent=1))
st.close()
