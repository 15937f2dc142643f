import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
import fastgaussianprocesses_b200 as fgp
dev = "cuda:0"
case = sys.argv[1] if len(sys.argv) > 1 else "mb_dnb2_T2_d3_n64_a2_b2x3_hyper3"
g = dict(np.load("tests/golden/%s.npz" % case))
T, d, alpha = int(g["T"]), int(g["d"]), int(g["alpha"])
ns = [int(v) for v in g["ns"]]; batch = [int(v) for v in g["batch"]]
def mk(kw):
    if str(g["family"]) == "lattice":
        seqs = [fgp.Lattice(d, generating_vector=g["z"][l], shift=g["shift"][l]) for l in range(T)]
        return fgp.FastGPLattice(seqs, num_tasks=T, alpha=alpha, device=dev, **kw)
    seqs = [fgp.DigitalNetB2(d, generating_matrices=g["C"][l], dshift=g["dshift"][l], t=int(g["t"])) for l in range(T)]
    return fgp.FastGPDigitalNetB2(seqs, num_tasks=T, alpha=alpha, device=dev, **kw)
kw = {"shape_batch": batch, "scale": torch.from_numpy(g["scale0"]), "lengthscales": torch.from_numpy(g["lengthscales0"]), "noise": torch.from_numpy(g["noise0"])}
if "deriv_0" in g:
    kw["derivatives"] = [torch.from_numpy(g["deriv_%d" % l]) for l in range(T)]
else:
    kw.update(factor_task_kernel=torch.from_numpy(g["factor_task_kernel0"]), noise_task_kernel=torch.from_numpy(g["noise_task_kernel0"]))
gp = mk(kw)
gp.get_x_next(ns); gp.add_y_next([torch.from_numpy(g["y_%d" % l]) for l in range(T)])
with torch.no_grad():
    norm, logdet = gp.get_inv_log_det_cache().get_norm_term_logdet_term()
print("norm ours", norm.reshape(-1).cpu().numpy(), "\nnorm ref ", g["norm_term0"].reshape(-1))
print("logdet ours", logdet.reshape(-1).cpu().numpy(), "\nlogdet ref ", g["logdet0"].reshape(-1))
print("kt ours", gp.gram_matrix_tasks.detach().cpu().numpy().reshape(-1)[:8])
H = g["scale0"].shape[0]
for h in range(H):
    kw1 = {"scale": torch.from_numpy(g["scale0"][h]), "lengthscales": torch.from_numpy(g["lengthscales0"][h]), "noise": torch.from_numpy(g["noise0"][h])}
    if "deriv_0" in g:
        kw1["derivatives"] = kw["derivatives"]
    else:
        kw1.update(factor_task_kernel=torch.from_numpy(g["factor_task_kernel0"][h]), noise_task_kernel=torch.from_numpy(g["noise_task_kernel0"][h]))
    g1 = mk(kw1)
    g1.get_x_next(ns)
    yb = [torch.from_numpy(g["y_%d" % l]).reshape(-1, H, ns[l])[0, h] for l in range(T)]
    g1.add_y_next(yb)
    with torch.no_grad():
        n1, l1 = g1.get_inv_log_det_cache().get_norm_term_logdet_term()
    print("h", h, "single-set norm", float(n1), "logdet", float(l1))
    L1 = g1._mt.lam_system(); Lb = gp._mt.lam_system()
    print("   lam diff", float((Lb.reshape((H,) + L1.shape)[h] - L1).abs().max()), float(L1.abs().max()))
