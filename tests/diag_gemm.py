"""GPU diagnostic (not a pytest): checks the two tcgen05 GEMM kernels against torch and, when the
MN-major (wgrad) descriptor guess is wrong, sweeps LBO/SBO/K-advance variants in subprocesses."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(kind):
    import torch
    import npp_b200
    nat = npp_b200._native
    torch.manual_seed(0)
    st = nat.current_stream()
    if kind == "km":
        for (m, n, k) in [(128, 256, 64), (128, 256, 128), (256, 512, 512), (1000, 512, 1472)]:
            a = torch.randn(m, k, device="cuda").half()
            b = torch.randn(n, k, device="cuda").half()
            c = torch.full((m, n), float("nan"), device="cuda")
            nat.check(nat.lib().npp_debug_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, st))
            ref = a.float() @ b.float().t()
            err = ((c - ref).abs().max() / ref.abs().max()).item()
            print(f"KM m={m} n={n} k={k} rel_err={err:.3e} nan={int(torch.isnan(c).sum())}", flush=True)
    else:
        for (rows, m, n, s) in [(64, 256, 256, 1), (128, 256, 256, 1), (1000, 256, 512, 3)]:
            a = torch.randn(rows, m, device="cuda").half()
            b = torch.randn(rows, n, device="cuda").half()
            c = torch.full((s, m, n), float("nan"), device="cuda")
            nat.check(nat.lib().npp_debug_wgrad(a.data_ptr(), b.data_ptr(), c.data_ptr(), rows, m, n, s, st))
            ref = a.float().t() @ b.float()
            got = c.sum(0)
            err = ((got - ref).abs().max() / ref.abs().max()).item()
            print(f"MN rows={rows} m={m} n={n} s={s} rel_err={err:.3e} nan={int(torch.isnan(got).sum())}", flush=True)


def run(kind, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    try:
        r = subprocess.run([sys.executable, __file__, kind], env=env, capture_output=True, text=True, timeout=180)
        print(f"--- {kind} {env_extra} rc={r.returncode}\n{r.stdout}{r.stderr[-1500:]}", flush=True)
    except subprocess.TimeoutExpired:
        print(f"--- {kind} {env_extra} TIMEOUT", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1])
        sys.exit(0)
    run("km", {})
    run("mn", {})
    if os.environ.get("NPP_DIAG_SWEEP", "1") == "1":
        for lbo, sbo, kadv in [(1024, 8192, 2048), (8192, 1024, 256), (128, 1024, 2048), (8192, 128, 2048), (1024, 8192, 256)]:
            run("mn", {"NPP_DEBUG_MN_LBO": str(lbo), "NPP_DEBUG_MN_SBO": str(sbo), "NPP_DEBUG_MN_KADV": str(kadv)})
