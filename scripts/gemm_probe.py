"""Bring-up probe for the tcgen05 GEMM (csrc/gemm_tc.cu): every case runs in its own process (a trapped launch poisons the
CUDA context) under a timeout; prints one JSON line per case.  Usage: python scripts/gemm_probe.py [--perf] [--sweep]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (M, N, K, a_mn, b_mn, splits, epilogue)
    "nn_small": (256, 256, 128, 0, 0, 1, ""),
    "nn_tile": (128, 256, 64, 0, 0, 1, ""),
    "nn_tails": (1000, 264, 200, 0, 0, 1, ""),
    "nn_n128": (300, 128, 256, 0, 0, 1, ""),
    "nn_n64": (300, 64, 256, 0, 0, 1, ""),
    "nn_big": (4096, 2048, 1024, 0, 0, 1, ""),
    "dgrad_small": (256, 256, 128, 0, 1, 1, ""),
    "dgrad_tails": (1000, 264, 200, 0, 1, 1, ""),
    "wgrad_small": (256, 256, 128, 1, 1, 1, ""),
    "wgrad_tails": (264, 200, 1000, 1, 1, 1, ""),
    "wgrad_split": (2048, 1024, 40320, 1, 1, 8, ""),
    "amn_only": (256, 256, 128, 1, 0, 1, ""),
    "epi_bias_gelu": (1000, 512, 256, 0, 0, 1, "bias,gelu,pre"),
    "epi_bias_res": (1000, 512, 256, 0, 0, 1, "bias,res"),
    "epi_res32_seg": (1000, 2048, 256, 0, 0, 1, "bias,res32,seg"),
    "epi_dact": (1000, 512, 256, 0, 1, 1, "dact"),
    "epi_lnfold": (1000, 512, 256, 0, 0, 1, "lnfold,bias"),
    "epi_f32out": (1000, 512, 256, 0, 0, 1, "f32out"),
}
PERF = {
    "perf_kv": (542080, 2048, 1024, 0, 0, 1, "bias,seg"),
    "perf_q": (40320, 2048, 1024, 0, 0, 1, "bias,seg"),
    "perf_proj": (40320, 1024, 1024, 0, 0, 1, "bias,res"),
    "perf_proj_bias": (40320, 1024, 1024, 0, 0, 1, "bias"),
    "perf_proj_plain": (40320, 1024, 1024, 0, 0, 1, ""),
    "perf_proj_res32": (40320, 1024, 1024, 0, 0, 1, "bias,res32"),
    "perf_mlp1": (40320, 4096, 1024, 0, 0, 1, "bias,gelu,pre"),
    "perf_mlp2": (40320, 1024, 4096, 0, 0, 1, "bias,res"),
    "perf_dgrad": (542080, 1024, 2048, 0, 1, 1, ""),
    "perf_wgrad": (2048, 1024, 542080, 1, 1, 16, ""),
    "perf_edge512": (327660, 512, 512, 0, 0, 1, "bias,silu,pre"),
    "perf_sq8k": (8192, 8192, 8192, 0, 0, 1, ""),
}


def run_case(name):
    import torch

    from anemoi_models_b200 import gemm as G

    M, N, K, a_mn, b_mn, splits, epi = {**CASES, **PERF}[name]
    perf = name.startswith("perf")
    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(1)
    A = torch.randn(M, K, generator=gen).to(dev).bfloat16() if not perf else torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, generator=gen).to(dev).bfloat16() if not perf else torch.randn(N, K, device=dev).bfloat16()
    a_arg = A.t().contiguous() if a_mn else A
    b_arg = B.t().contiguous() if b_mn else B
    kw = {}
    opts = set(epi.split(",")) if epi else set()
    ref = None
    if not perf or M * N <= 2 ** 28:
        ref = A.float() @ B.float().t()
    bias = torch.randn(N, device=dev) if "bias" in opts else None
    if "lnfold" in opts:
        rs, rt, cv = torch.rand(M, device=dev) + 0.5, torch.randn(M, device=dev), torch.randn(N, device=dev)
        kw.update(row_scale=rs, row_shift=rt, col_vec=cv)
        if ref is not None:
            ref = rs[:, None] * ref + rt[:, None] * cv[None, :]
    if bias is not None:
        kw["bias"] = bias
        if ref is not None:
            ref = ref + bias
    pre = None
    if "dact" in opts:
        pre = torch.randn(M, N, device=dev).bfloat16()
        kw.update(dact_pre=pre, act=1)
        if ref is not None:
            x = pre.float().requires_grad_(True)
            torch.nn.functional.gelu(x).sum().backward()
            ref = ref * x.grad
    for nm, code in (("gelu", 1), ("silu", 0)):
        if nm in opts:
            kw["act"] = code
            if "pre" in opts:
                pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
                kw["pre_out"] = pre
            if ref is not None:
                pre_ref = ref
                r = ref.bfloat16().float() if "pre" in opts else ref
                ref = torch.nn.functional.gelu(r) if code == 1 else torch.nn.functional.silu(r)
    res = None
    if "res" in opts or "res32" in opts:
        res = torch.randn(M, N, device=dev)
        if "res" in opts:
            res = res.bfloat16()
        kw["residual"] = res
        if ref is not None:
            ref = ref + res.float()
    outs = None
    if "seg" in opts:
        seg = N // 2
        outs = [torch.empty(M, seg, device=dev, dtype=torch.bfloat16) for _ in range(2)]
        kw.update(out=outs, seg_cols=seg)
    if "f32out" in opts:
        kw["out_dtype"] = torch.float32
    if splits > 1:
        kw["splits"] = splits
        kw["out_dtype"] = torch.float32

    def call():
        return G.gemm(a_arg, b_arg, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), **kw)

    out = call()
    torch.cuda.synchronize()
    rec = {"case": name, "M": M, "N": N, "K": K, "a_mn": a_mn, "b_mn": b_mn, "splits": splits, "epi": epi,
           "desc": os.environ.get("AB2_GEMM_DESC", "default"), "cg": int(os.environ.get("AB2_GEMM_CG", "2"))}
    if ref is not None:
        got = torch.cat([o.float() for o in out], dim=1) if isinstance(out, list) else out.float()
        err = (got - ref).abs()
        scale = max(1.0, float(ref.abs().max()))
        rec["max_rel_err"] = float(err.max()) / scale
        rec["frac_bad"] = float((err > 2e-2 * scale).float().mean())
        if "pre" in opts and pre is not None and "dact" not in opts:
            rec["pre_rel_err"] = float((pre.float() - pre_ref).abs().max()) / max(1.0, float(pre_ref.abs().max()))
        rec["ok"] = bool(rec["max_rel_err"] < 1e-2 and rec.get("pre_rel_err", 0.0) < 1e-2)
    if perf:
        def timeit(fn, n=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
            ev[0].record()
            for i in range(n):
                fn()
                ev[i + 1].record()
            torch.cuda.synchronize()
            ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
            return ts[len(ts) // 2]

        ms = timeit(call)
        rec["ms"] = ms
        rec["tflops"] = 2.0 * M * N * K / ms / 1e9
        a2 = A if not a_mn else a_arg.t()
        if a_mn and b_mn:
            ms_ref = timeit(lambda: a_arg.t() @ b_arg)  # dY^T x
        elif b_mn:
            ms_ref = timeit(lambda: A @ b_arg)
        elif bias is not None:
            ms_ref = timeit(lambda: torch.nn.functional.linear(A, B, bias.bfloat16()))
        else:
            ms_ref = timeit(lambda: A @ B.t())
        rec["cublas_ms"] = ms_ref
        rec["cublas_tflops"] = 2.0 * M * N * K / ms_ref / 1e9
    print("PROBE " + json.dumps(rec), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        return
    names = list(CASES)
    if "--perf" in sys.argv:
        names += list(PERF)
    if "--only-perf" in sys.argv:
        names = list(PERF)
    results = {}

    def launch(name, env=None):
        e = dict(os.environ)
        if env:
            e.update(env)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=180, env=e)
            line = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            if line:
                rec = json.loads(line[-1][6:])
            else:
                rec = {"case": name, "ok": False, "error": (r.stderr or r.stdout)[-600:]}
        except subprocess.TimeoutExpired:
            rec = {"case": name, "ok": False, "error": "timeout"}
        rec["wall_s"] = round(time.time() - t0, 1)
        if env:
            rec["env"] = env
        print(json.dumps(rec), flush=True)
        return rec

    for n in names:
        results[n] = launch(n)
    if "--cg1" in sys.argv:  # the same cases on single-CTA tiles (cta_group::1), for the A/B table
        for n in names:
            launch(n, {"AB2_GEMM_CG": "1"})
    if "--sweep" in sys.argv:
        # descriptor byte offsets k_lbo,k_sbo,mn_lbo,mn_sbo: alternatives in case the defaults are wrong
        for n, alts in (("nn_small", ["0,1024,8192,1024", "1024,1024,8192,1024", "16,128,8192,1024"]),
                        ("dgrad_small", ["16,1024,1024,8192", "16,1024,8192,128", "16,1024,128,8192", "16,1024,8192,2048", "16,1024,2048,8192"]),
                        ("wgrad_small", ["16,1024,1024,8192", "16,1024,8192,128", "16,1024,128,8192", "16,1024,8192,2048", "16,1024,2048,8192"])):
            if not results.get(n, {}).get("ok", False):
                for alt in alts:
                    launch(n, {"AB2_GEMM_DESC": alt})


if __name__ == "__main__":
    main()
