"""Check the tcgen05 GEMM kernels against the CUDA-core (SIMT) GEMM on the same bf16 inputs.
Run under `timeout` (a wrong descriptor can deadlock an mbarrier wait)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acm_gnn_b200 import _lib

def st(): return torch.cuda.current_stream().cuda_stream
BF, SIMT, TC = _lib.ACM_BF16, _lib.GEMM_SIMT, _lib.GEMM_TCGEN05

def relerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))

def run(n, fin, fp, relu):
    torch.manual_seed(n + fin + fp)
    dev = "cuda"
    ldx = (fin + 7) // 8 * 8
    x = torch.zeros(n, ldx, device=dev, dtype=torch.bfloat16); x[:, :fin] = torch.randn(n, fin, device=dev)
    wcat = (torch.randn(fin, 3 * fp, device=dev) / fin ** 0.5).to(torch.bfloat16)
    wcat_t = torch.zeros(3 * fp, ldx, device=dev, dtype=torch.bfloat16); wcat_t[:, :fin] = wcat.t()
    out = {}
    for impl in (SIMT, TC):
        h_lh = torch.full((n, 2 * fp), 7.0, device=dev, dtype=torch.bfloat16); h_i = torch.full((n, fp), 7.0, device=dev, dtype=torch.bfloat16)
        _lib.call("acm_gemm_xw_fwd", impl, BF, x.data_ptr(), ldx, wcat.data_ptr(), wcat_t.data_ptr(), h_lh.data_ptr(), h_i.data_ptr(), n, fin, fp, relu, st())
        dh = (torch.randn(n, 3 * fp, device=dev)).to(torch.bfloat16)
        torch.manual_seed(1); dh = torch.randn(n, 3 * fp, device=dev).to(torch.bfloat16)
        dw = torch.zeros(fin, 3 * fp, device=dev)
        _lib.call("acm_gemm_bwd_dw", impl, BF, x.data_ptr(), ldx, dh.data_ptr(), dw.data_ptr(), n, fin, fp, st())
        dx = torch.full((n, fin), 7.0, device=dev)
        _lib.call("acm_gemm_bwd_dx", impl, BF, dh.data_ptr(), wcat.data_ptr(), wcat_t.data_ptr(), ldx, dx.data_ptr(), fin, n, fin, fp, st())
        torch.cuda.synchronize()
        out[impl] = (h_lh, h_i, dw, dx)
    ref = x[:, :fin].float() @ wcat.float()
    names = ["h_lh", "h_i", "dw", "dx"]
    errs = [relerr(a, b) for a, b in zip(out[TC], out[SIMT])]
    e_ref = relerr(torch.cat([out[TC][0] if not relu else out[TC][0], out[TC][1]], 1), torch.cat([ref[:, :2*fp].relu() if relu else ref[:, :2*fp], ref[:, 2*fp:]], 1))
    ok = errs[0] < 1e-2 and errs[1] < 1e-2 and errs[2] < 2e-3 and errs[3] < 2e-3 and e_ref < 1e-2
    print(f"n={n} fin={fin} fp={fp} relu={relu}: " + " ".join(f"{k}={e:.2e}" for k, e in zip(names, errs)) + f" vs_torch={e_ref:.2e} {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok

if __name__ == "__main__":
    shapes = [(1000, 64, 256, 0), (333, 1433, 64, 1), (5000, 7, 8, 0), (4096, 256, 16, 0), (130, 24, 32, 1), (20000, 256, 256, 0), (777, 100, 128, 0)]
    ok = all([run(*s) for s in shapes])
    print("TC_PROBE", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
