"""First-contact GPU diagnostics: each stage runs in its own subprocess (a trapped kernel poisons
only its own CUDA context) and prints numbers instead of asserting. Usage: python tests/gpu_diag.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = {}


def stage(fn):
    STAGES[fn.__name__] = fn
    return fn


def _gemm(precision, shapes):
    import torch
    from insenticap_model_b200 import _lib
    lib = _lib.load()
    prec = _lib.PRECISIONS[precision]
    for (M, N, K) in shapes:
        g = torch.Generator().manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) / K ** 0.5
        bias = torch.randn(N, generator=g)
        out = torch.full((M, N), float("nan"), device="cuda")
        ws = torch.empty(lib.isc_gemm_workspace_bytes(prec, M, N, K), dtype=torch.uint8, device="cuda")
        Ad, Wd, bd = A.cuda(), W.cuda(), bias.cuda()
        rc = lib.isc_gemm_tn(prec, _lib.ptr(Ad), K, _lib.ptr(Wd), K, _lib.ptr(bd), _lib.ptr(out), N,
                             M, N, K, 0, _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
        torch.cuda.synchronize()
        ref = A.double() @ W.double().t() + bias.double()
        refb = A.bfloat16().double() @ W.bfloat16().double().t() + bias.double()
        o = out.cpu().double()
        print("  %s %s rc=%d  max|err| vs fp64 %.3e  vs bf16-rounded %.3e  nan=%d" % (
            precision, (M, N, K), rc, (o - ref).abs().max(), (o - refb).abs().max(), int(torch.isnan(o).sum())), flush=True)
        if (o - ref).abs().max() > 0.05:
            bad = ((o - ref).abs() > 0.05).nonzero()
            print("   first bad entries:", bad[:6].tolist(), " count", len(bad), " rows bad:",
                  sorted(set((bad[:, 0] // 8 * 8).tolist()))[:12], "cols bad:", sorted(set((bad[:, 1] // 8 * 8).tolist()))[:12])


@stage
def gemm_fp32():
    _gemm("fp32", [(300, 200, 136), (64, 64, 16)])


@stage
def gemm_bf16():
    _gemm("bf16", [(128, 128, 64), (128, 128, 128), (256, 256, 512), (300, 200, 136)])


@stage
def gemm_bf16x3():
    _gemm("bf16x3", [(128, 128, 64), (256, 256, 512), (300, 200, 136), (3072, 2048, 1536)])


def _decode(precision):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from insenticap_model_b200 import synthetic as syn
    from oracle import captioner_oracle as O
    from tests._common import model, params, to_cuda
    gd = np.load(os.path.join(ROOT, "tests", "golden", "decode_golden.npz"))
    V, B, T = 10000, 8, 16
    fc, att, cpts, sentis, labels = syn.synthetic_inputs(B, V, seed=1)
    m = model(V, 0, precision)
    p = params(V, 0)
    with torch.no_grad():
        f = O.prologue(p, fc, att, cpts, sentis, labels)
    t, _ = m.prologue(*to_cuda(fc, att, cpts, sentis, labels))
    torch.cuda.synchronize()
    for name in ("fc", "att", "p_att", "sw", "p_sw", "sl", "cpt_feats"):
        print("  prologue %-9s max|err| %.3e" % (name, (t[name].float().cpu() - f[name]).abs().max()), flush=True)
    with torch.no_grad():
        f2 = dict(f)
    g = torch.Generator().manual_seed(7)
    h0 = torch.randn(2, B, 512, generator=g) * 0.3
    c0 = torch.randn(2, B, 512, generator=g) * 0.3
    it = torch.randint(0, V, (B,), generator=g)
    lp, (h1, c1) = m.forward_step(it.cuda(), (h0.cuda(), c0.cuda()), t["fc"], t["att"], f["p_att"].cuda(), t["sw"], f["p_sw"].cuda(), t["sl"])
    torch.cuda.synchronize()
    with torch.no_grad():
        lp_o, (h_o, c_o), (cw, sw, gw) = O.step(p, it, (h0, c0), f, want_weights=True)
    print("  step h_att %.3e h_lang %.3e c_att %.3e c_lang %.3e logprobs %.3e" % (
        (h1[0].cpu() - h_o[0]).abs().max(), (h1[1].cpu() - h_o[1]).abs().max(), (c1[0].cpu() - c_o[0]).abs().max(),
        (c1[1].cpu() - c_o[1]).abs().max(), (lp.cpu() - lp_o).abs().max()), flush=True)
    mcw, msw, mgw = m._step_weights
    print("  step weights cont %.3e senti %.3e gate %.3e" % ((mcw.cpu() - cw).abs().max(), (msw.cpu() - sw).abs().max(),
                                                              (mgw.cpu() - gw).abs().max()), flush=True)
    seq, lps, mask = m(*to_cuda(fc, att, cpts, sentis, labels), T, 1, mode="rl")
    torch.cuda.synchronize()
    print("  greedy tokens equal golden: %s  (%d/%d)  lp err %.3e" % (
        np.array_equal(seq.cpu().numpy(), gd["cfg1_greedy_seq"]), int((seq.cpu().numpy() == gd["cfg1_greedy_seq"]).sum()),
        seq.numel(), np.abs(lps.cpu().numpy() - gd["cfg1_greedy_lp"]).max()), flush=True)
    tk, sc, ln = m.beam_search(*to_cuda(fc, att, sentis, labels), beam_size=3, max_seq_len=T)
    torch.cuda.synchronize()
    print("  beam3 tokens equal golden: %s (%d/%d) score err %.3e" % (
        np.array_equal(tk.cpu().numpy(), gd["cfg1_beam3_tokens"]), int((tk.cpu().numpy() == gd["cfg1_beam3_tokens"]).sum()),
        tk.numel(), np.abs(sc.cpu().numpy() - gd["cfg1_beam3_scores"]).max()), flush=True)
    print("  beam img0 scores", sc[0].tolist(), "golden", gd["cfg1_beam3_scores"][0].tolist())


@stage
def decode_fp32():
    _decode("fp32")


@stage
def decode_bf16x3():
    _decode("bf16x3")


@stage
def decode_bf16():
    _decode("bf16")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.path.insert(0, ROOT)
        STAGES[sys.argv[1]]()
        sys.exit(0)
    for name in STAGES:
        print("== stage", name, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], cwd=ROOT, timeout=420,
                               capture_output=True, text=True)
            print(r.stdout[-4000:])
            if r.returncode != 0:
                print("  [exit %d] stderr tail:\n%s" % (r.returncode, r.stderr[-2500:]))
        except subprocess.TimeoutExpired:
            print("  [timeout]")
