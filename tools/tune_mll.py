"""Tuning probe: per-iteration time of the fused fit loop (warm, back-to-back graph replays) and per-kernel times."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
fam = sys.argv[3] if len(sys.argv) > 3 else "lattice"
dev = torch.device("cuda:0")
n = 1 << log2n
if fam == "lattice":
    gp = fgp.FastGPLattice(fgp.Lattice(d, seed=7), device=dev)
else:
    gp = fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, seed=7), device=dev)
x = gp.get_x_next(n)
gp.add_y_next(torch.cos(2 * np.pi * x).sum(1))
st = gp.fit_stepper()
for _ in range(5):
    st.step()
torch.cuda.synchronize()
K = 100
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    st.step()
e1.record()
torch.cuda.synchronize()
warm = e0.elapsed_time(e1) / K * 1e3
G = st.GRAPH_ITERS  # chunk graphs of fit(): G iterations per launch, the C -> A edges inside the graph
st.replay(G)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    st.replay(G)
e1.record()
torch.cuda.synchronize()
warm_chunk = e0.elapsed_time(e1) / (10 * G) * 1e3
flush = torch.empty(256 * 1024 * 1024 // 8, device=dev)
kern = st.kernel_times(reps=10, flush=flush)
kern_w = st.kernel_times(reps=10, flush=None)
print(json.dumps({"cfg": {k: v for k, v in os.environ.items() if k.startswith("FGP_")}, "fam": fam, "log2n": log2n, "d": d, "warm_us_per_iter": round(warm, 2), "warm_us_per_iter_chunk_graph": round(warm_chunk, 2),
                  "cold_kernels_us": {k["name"]: round(k["ms"] * 1e3, 1) for k in kern}, "eager_warm_kernels_us": {k["name"]: round(k["ms"] * 1e3, 1) for k in kern_w}}))
