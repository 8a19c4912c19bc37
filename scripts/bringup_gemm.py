"""GPU bring-up for the tcgen05 GEMM: each case runs in its own subprocess under a timeout so a
deadlocked mbarrier protocol cannot hang the box.  Usage: python scripts/bringup_gemm.py [case ...]"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    # name: (M, N, K, a_major, b_major, epilogue, split_k, dbg_flags)
    "kk_small": (256, 256, 128, 0, 0, "store", 1, 0),
    "kk_small_lbo16": (256, 256, 128, 0, 0, "store", 1, 4),
    "kk_bn128": (256, 128, 192, 0, 0, "store", 1, 0),
    "kk_ragged": (300, 200, 136, 0, 0, "store", 1, 0),
    "kk_qkv": (16000, 2304, 768, 0, 0, "bias", 1, 0),
    "kk_ffn1_gelu": (16000, 3072, 768, 0, 0, "gelu", 1, 0),
    "kk_ffn2": (16000, 768, 3072, 0, 0, "bias", 1, 0),
    "kk_dgelu": (4096, 3072, 768, 0, 0, "dgelu", 1, 0),
    "kk_vocab": (2048, 30522, 768, 0, 0, "bias", 1, 0),
    "kk_f32": (512, 768, 768, 0, 0, "f32", 1, 0),
    "mnmn_small": (256, 256, 128, 1, 1, "f32", 1, 0),
    "mnmn_small_swap": (256, 256, 128, 1, 1, "f32", 1, 2),
    "mnk_small": (256, 256, 128, 1, 0, "f32", 1, 0),
    "kmn_small": (256, 256, 128, 0, 1, "f32", 1, 0),
    "mnmn_wgrad": (768, 3072, 16000, 1, 1, "atomic", 8, 0),
    "mnmn_wgrad_ragged": (768, 74, 6400, 1, 1, "atomic", 4, 0),
    "kmn_dgrad": (4096, 768, 3072, 0, 1, "store", 1, 0),
    # single-CTA kernel forced (dbg bit 3) for A/B timing against the CTA-pair kernel
    "1cta_qkv": (16000, 2304, 768, 0, 0, "bias", 1, 8),
    "1cta_ffn1_gelu": (16000, 3072, 768, 0, 0, "gelu", 1, 8),
    "1cta_ffn2": (16000, 768, 3072, 0, 0, "bias", 1, 8),
    "kk_big": (16384, 4096, 4096, 0, 0, "store", 1, 0),
    "1cta_big": (16384, 4096, 4096, 0, 0, "store", 1, 8),
    "kk_dec_dgrad": (16000, 768, 30522, 0, 1, "store", 1, 0),
    "mnmn_dec_wgrad": (30522, 768, 16000, 1, 1, "atomic", 1, 0),
    "kk_c3_ffn1": (73600, 3072, 768, 0, 0, "gelu", 1, 0),
    # dbg bit 6 (64) flips the 8 / 16 epilogue-warp choice: same-box A/B
    "ab_qkv_8": (16000, 2304, 768, 0, 0, "bias", 1, 0), "ab_qkv_16": (16000, 2304, 768, 0, 0, "bias", 1, 64),
    "ab_gelu_16": (16000, 3072, 768, 0, 0, "gelu", 1, 0), "ab_gelu_8": (16000, 3072, 768, 0, 0, "gelu", 1, 64),
    "ab_ffn2_8": (16000, 768, 3072, 0, 0, "bias", 1, 0), "ab_ffn2_16": (16000, 768, 3072, 0, 0, "bias", 1, 64),
    "ab_wgrad_8": (768, 3072, 16000, 1, 1, "atomic", 5, 0), "ab_wgrad_16": (768, 3072, 16000, 1, 1, "atomic", 5, 64),
    "ab_c3qkv_8": (73600, 2304, 768, 0, 0, "bias", 1, 0), "ab_c3qkv_16": (73600, 2304, 768, 0, 0, "bias", 1, 64),
    "ab_c3gelu_16": (73600, 3072, 768, 0, 0, "gelu", 1, 0), "ab_c3gelu_8": (73600, 3072, 768, 0, 0, "gelu", 1, 64),
    "y_ffn1_plain": (16000, 3072, 768, 0, 0, "bias", 1, 0),
    "y_ffn1_gelu_noaux": (16000, 3072, 768, 0, 0, "gelu_noaux", 1, 0),
    "y_ffn1_gelu_half": (8000, 3072, 768, 0, 0, "gelu", 1, 0),
    "y_ffn1_gelu_quarter": (4000, 3072, 768, 0, 0, "gelu", 1, 0),
    "y_ffn1_plain_c3": (73600, 3072, 768, 0, 0, "bias", 1, 0),
    # epilogue dissection (results are wrong by construction): 16 = no global stores, 32 = no TMEM loads either
    "x_qkv_nostore": (16000, 2304, 768, 0, 0, "bias", 1, 16),
    "x_qkv_nold": (16000, 2304, 768, 0, 0, "bias", 1, 32),
    "x_ffn1_nostore": (16000, 3072, 768, 0, 0, "gelu", 1, 16),
    "x_ffn1_nold": (16000, 3072, 768, 0, 0, "gelu", 1, 32),
    "x_ffn2_nold": (16000, 768, 3072, 0, 0, "bias", 1, 32),
    "x_big_nold": (16384, 4096, 4096, 0, 0, "store", 1, 32),
    # round 2: the MOSEI-shape epilogue-bound GEMMs, dissected the same way
    "z_ffn1_gg": (73600, 3072, 768, 0, 0, "gelu_grad", 1, 0),
    "z_ffn1_gg_nostore": (73600, 3072, 768, 0, 0, "gelu_grad", 1, 16),
    "z_ffn1_gg_nold": (73600, 3072, 768, 0, 0, "gelu_grad", 1, 32),
    "z_ffn2d_mul": (73600, 3072, 768, 0, 1, "mul_aux", 1, 0),
    "z_ffn2d_mul_nostore": (73600, 3072, 768, 0, 1, "mul_aux", 1, 16),
    "z_ffn2d_mul_nold": (73600, 3072, 768, 0, 1, "mul_aux", 1, 32),
    "z_qkv": (73600, 2304, 768, 0, 0, "bias", 1, 0),
    "z_qkv_nostore": (73600, 2304, 768, 0, 0, "bias", 1, 16),
    "z_qkv_nold": (73600, 2304, 768, 0, 0, "bias", 1, 32),
    "z_ffn1_plain": (73600, 3072, 768, 0, 0, "bias", 1, 0),
    "z_ffn2": (73600, 768, 3072, 0, 0, "bias", 1, 0),
    "z_ffn2_nold": (73600, 768, 3072, 0, 0, "bias", 1, 32),
    "z_wo": (73600, 768, 768, 0, 0, "bias", 1, 0),
    "z_wo_nold": (73600, 768, 768, 0, 0, "bias", 1, 32),
    # dbg bit 8 (256): st.global epilogue instead of the TMA stores
    "s_ffn1_gg": (73600, 3072, 768, 0, 0, "gelu_grad", 1, 256),
    "s_ffn2d_mul": (73600, 3072, 768, 0, 1, "mul_aux", 1, 256),
    "s_qkv": (73600, 2304, 768, 0, 0, "bias", 1, 256),
    "s_wo": (73600, 768, 768, 0, 0, "bias", 1, 256),
    "s_ffn2": (73600, 768, 3072, 0, 0, "bias", 1, 256),
    "z_c2_qkv": (16000, 2304, 768, 0, 0, "bias", 1, 0), "s_c2_qkv": (16000, 2304, 768, 0, 0, "bias", 1, 256),
    "z_c2_ffn1_gg": (16000, 3072, 768, 0, 0, "gelu_grad", 1, 0), "s_c2_ffn1_gg": (16000, 3072, 768, 0, 0, "gelu_grad", 1, 256),
    "z_ragged_gg": (1000, 1000, 136, 0, 0, "gelu_grad", 1, 0),
    "z_ragged_bias": (1000, 1000, 136, 0, 0, "bias", 1, 0),
    "z_ragged_mul": (1000, 1024, 136, 0, 1, "mul_aux", 1, 0),
    "z_vocab": (2048, 30522, 768, 0, 0, "bias", 1, 0),
    "y_ffn1_gelu_1stream": (73600, 3072, 768, 0, 0, "gelu_noaux", 1, 0),
    "y_ffn1_gg_8warps": (73600, 3072, 768, 0, 0, "gelu_grad", 1, 64),
}


def run_case(name):
    import torch
    from msa_b200 import capi
    M, N, K, am, bm, epi, split, flags = CASES[name]
    torch.manual_seed(0)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device=dev) * 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    ref = A.float() @ B.float().t()
    A_in = A.t().contiguous() if am else A
    B_in = B.t().contiguous() if bm else B
    # pad leading dims to multiples of 8 where needed
    def pad_ld(t):
        r, c = t.shape
        if c % 8 == 0:
            return t
        buf = torch.zeros(r, (c + 7) // 8 * 8, device=dev, dtype=t.dtype)
        buf[:, :c] = t
        return buf[:, :c]
    A_in, B_in = pad_ld(A_in), pad_ld(B_in)
    ldc = (N + 7) // 8 * 8
    out_f32 = epi in ("f32", "atomic")
    Cbuf = torch.zeros(M, ldc, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    C = Cbuf[:, :N]
    aux = None
    kw = {}
    if epi == "store":
        e = capi.EPI_STORE_BF16
    elif epi == "bias":
        e = capi.EPI_STORE_BF16; kw["bias"] = bias; ref = ref + bias
    elif epi == "gelu":
        e = capi.EPI_GELU_BF16; kw["bias"] = bias
        aux = torch.zeros(M, ldc, device=dev, dtype=torch.bfloat16)[:, :N]
        pre = (ref + bias).to(torch.bfloat16).float()
        ref = torch.nn.functional.gelu(pre)
    elif epi == "gelu_noaux":
        e = capi.EPI_GELU_BF16; kw["bias"] = bias
        pre = (ref + bias).to(torch.bfloat16).float()
        ref = torch.nn.functional.gelu(pre)
    elif epi == "gelu_grad":
        e = capi.EPI_GELU_GRAD_BF16; kw["bias"] = bias
        aux = torch.zeros(M, ldc, device=dev, dtype=torch.bfloat16)[:, :N]
        ref = torch.nn.functional.gelu(ref + bias)
    elif epi == "mul_aux":
        e = capi.EPI_MUL_AUX_BF16
        aux = torch.randn(M, ldc, device=dev).to(torch.bfloat16)[:, :N]
        ref = ref * aux.float()
    elif epi == "dgelu":
        e = capi.EPI_DGELU_BF16
        aux = torch.randn(M, ldc, device=dev).to(torch.bfloat16)[:, :N]
        u = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(u).sum().backward()
        ref = ref * u.grad
    elif epi == "f32":
        e = capi.EPI_STORE_F32
    elif epi == "atomic":
        e = capi.EPI_ATOMIC_ADD_F32
        Cbuf.fill_(1.0); ref = ref + 1.0
    def launch():
        capi.gemm(A_in, B_in, C, M, N, K, a_major=am, b_major=bm, epilogue=e, aux=aux, split_k=split,
                  dbg_flags=flags, **kw)
    launch()
    torch.cuda.synchronize()
    got = C.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    res = {"case": name, "max_abs_err": err, "ref_max": scale, "rel": err / max(scale, 1e-9)}
    if epi == "gelu":
        res["aux_err"] = (aux.float() - pre).abs().max().item()
    if epi == "gelu_grad":
        u = (A.float() @ B.float().t() + bias).requires_grad_(True)
        torch.nn.functional.gelu(u).sum().backward()
        res["aux_err"] = (aux.float() - u.grad).abs().max().item()
    if True:
        for _ in range(3):
            launch()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        s.record()
        for _ in range(iters):
            launch()
        t.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(t) / iters
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    os.makedirs("gpurun_out", exist_ok=True)
    results = []
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                results.append(json.loads(line[-1][7:]))
            else:
                results.append({"case": n, "error": (r.stderr or r.stdout)[-600:]})
        except subprocess.TimeoutExpired:
            results.append({"case": n, "error": "TIMEOUT (hang)"})
        print(json.dumps(results[-1]), flush=True)
    json.dump(results, open("gpurun_out/bringup_gemm.json", "w"), indent=1)
