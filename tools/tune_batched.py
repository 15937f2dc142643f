"""Batched fits (BASELINE.json configs[4]): B lattice GPs of n = 2^log2n, d = 8 in one object; microseconds per batched iteration and GP-iterations/s.
    python tools/tune_batched.py [B] [log2n]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = int(sys.argv[2]) if len(sys.argv) > 2 else 18
d, dev = 8, torch.device("cuda:0")
gp = fgp.FastGPLattice(fgp.Lattice(d, seed=100, generating_vector=bench.gen_vec(d)), device=dev, shape_batch=torch.Size([B]),
                       shape_scale=torch.Size([B, 1]), shape_lengthscales=torch.Size([B, d]), shape_noise=torch.Size([B, 1]))
x = gp.get_x_next(1 << m)
y = bench.f_synth(x)
gp.add_y_next(torch.stack([y * (1.0 + 0.05 * k) + 0.01 * k for k in range(B)]))
st = gp.fit_stepper()
for _ in range(3):
    st.step()
torch.cuda.synchronize()
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    st.step()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / K * 1e3
kern = st.kernel_times(reps=5, flush=None)
print(json.dumps({"cfg": {k: v for k, v in os.environ.items() if k.startswith("FGP_")}, "B": B, "log2n": m, "us_per_batched_iteration": round(us, 1),
                  "gp_iterations_per_s": round(B * 1e6 / us), "kernels_us": {k["name"]: round(k["ms"] * 1e3, 1) for k in kern}}))
st.close()
