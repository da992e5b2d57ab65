"""Development check: RVQ stage >= 1 with the residual update fused into the fp16 assignment kernel
(vqb200_vq_assign_residual) against the two stand-alone calls with the exact kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes
import vqb200
from vqb200 import _lib
from vqb200._lib import ptr, stream_ptr, check
from ctypes import c_size_t

lib = _lib.load()
dev = torch.device("cuda:0")


def case(B, T, K, perm=False, seed=0):
    torch.manual_seed(seed)
    D = 64
    Wp = 0.3 * torch.randn(K, D, device=dev)
    W = 0.2 * torch.randn(K, D, device=dev)
    stp, st = vqb200.QuantizerState(K, D, dev), vqb200.QuantizerState(K, D, dev)
    st.refresh(W)
    if perm:
        r_in = (0.5 * torch.randn(B, T, D, device=dev)).permute(0, 2, 1)
    else:
        r_in = 0.5 * torch.randn(B, D, T, device=dev)
    idx_prev = vqb200.vq_assign(r_in, Wp, stp, _lib.ASSIGN_SIMT)
    N = B * T
    sB, sC, sT = r_in.stride()
    s = stream_ptr(dev)
    # reference: stand-alone gather (residual) + exact assignment
    r_ref = torch.empty(B, D, T, device=dev)
    check(lib.vqb200_vq_gather_st(ptr(r_in), B, D, T, sB, sC, sT, ptr(Wp), ptr(idx_prev), K, None, ptr(r_ref), None, 0, None, s), "g")
    i_ref = vqb200.vq_assign(r_ref, W, st, _lib.ASSIGN_SIMT)
    out = {}
    for name, algo in (("f16", _lib.ASSIGN_TC), ("split", _lib.ASSIGN_TC_SPLIT)):
        r_out = torch.empty(B, D, T, device=dev)
        if perm:
            r_out = torch.empty(B, T, D, device=dev).permute(0, 2, 1)
        idx = torch.empty(B, T, dtype=torch.int32, device=dev)
        ws = st.assign_workspace(N)
        check(lib.vqb200_vq_assign_residual(ptr(r_in), B, D, T, sB, sC, sT, ptr(Wp), ptr(idx_prev), K, ptr(r_out), ptr(W), ptr(st.ee),
                                            ptr(st.image), ptr(st.info), K, ptr(idx), ptr(ws), c_size_t(ws.numel()), algo, s), "ar")
        torch.cuda.synchronize()
        out[name] = dict(idx_mismatch=int((idx != i_ref).sum()), r_bits_equal=bool(torch.equal(r_out.contiguous(), r_ref)),
                         flagged=int(ws.view(torch.int32)[0]), err=int(ws.view(torch.int32)[1]))
    print(json.dumps(dict(B=B, T=T, K=K, perm=perm, **out)), flush=True)
    return out["f16"]["idx_mismatch"] + (0 if out["f16"]["r_bits_equal"] else 1)


if __name__ == "__main__":
    bad = 0
    bad += case(4096, 10, 1024)
    bad += case(3001, 7, 300)
    bad += case(20000, 1, 512, perm=True)
    bad += case(200000, 10, 1024)
    print("FUSED BAD", bad)
