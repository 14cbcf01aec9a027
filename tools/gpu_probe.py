"""GPU probe: run building-block checks one by one, survive sticky CUDA errors by restarting.
Usage (on the GPU box):  python tools/gpu_probe.py            # driver: runs all cases, writes gpurun_out/probe.json
                         python tools/gpu_probe.py --from i   # worker
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'gpurun_out')


def cases():
    import torch
    from osu_dreamer_b200 import lib
    dev = 'cuda'
    g = torch.Generator(device='cpu').manual_seed(0)

    def rnd(*s, dtype=torch.bfloat16):
        return torch.randn(*s, generator=g).to(dtype).to(dev)

    def rel(a, b):
        return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20))

    def gemm_case(M, N, K, dt, amaj, bmaj, epi=0, split=1, cfp32=False, bias=True):
        A = rnd(M, K, dtype=dt)
        B = rnd(N, K, dtype=dt)
        Aop = A if amaj == 0 else A.t().contiguous()
        Bop = B if bmaj == 0 else B.t().contiguous()
        bs = rnd(N, dtype=torch.float32) if bias and epi != 2 else None
        ref = A.float() @ B.float().t()
        if epi == 2:
            C = torch.ones(M, N, dtype=torch.float32, device=dev)
            ref = ref + 1
        else:
            C = torch.full((M, N), float('nan'), dtype=torch.float32 if cfp32 else torch.bfloat16, device=dev)
            if bs is not None:
                ref = ref + bs
            if epi == 1:
                ref = torch.nn.functional.silu(ref)
        lib.gemm(Aop, Bop, C, bias=bs, a_major=amaj, b_major=bmaj, epi=epi, split_k=split)
        torch.cuda.synchronize()
        return rel(C, ref)

    bf, f32 = torch.bfloat16, torch.float32
    cs = []
    cs.append(('bf16_KK_small', lambda: gemm_case(128, 128, 64, bf, 0, 0, cfp32=True, bias=False)))
    cs.append(('bf16_KK_k512', lambda: gemm_case(256, 256, 512, bf, 0, 0, cfp32=True)))
    cs.append(('bf16_KK_bn256', lambda: gemm_case(128 * 160, 512, 512, bf, 0, 0)))
    cs.append(('bf16_KK_tail', lambda: gemm_case(128 * 150 + 37, 3072, 512, bf, 0, 0)))
    cs.append(('bf16_KK_silu', lambda: gemm_case(1000, 128, 128, bf, 0, 0, epi=1)))
    cs.append(('bf16_MNK_small', lambda: gemm_case(128, 128, 64, bf, 1, 0, cfp32=True, bias=False)))
    cs.append(('bf16_KMN_small', lambda: gemm_case(128, 128, 64, bf, 0, 1, cfp32=True, bias=False)))
    cs.append(('bf16_MNMN_small', lambda: gemm_case(128, 128, 64, bf, 1, 1, cfp32=True, bias=False)))
    cs.append(('bf16_MNMN_k256', lambda: gemm_case(256, 256, 256, bf, 1, 1, cfp32=True, bias=False)))
    cs.append(('bf16_wgrad_split', lambda: gemm_case(3072, 512, 128 * 64 + 19 * 8, bf, 1, 1, epi=2, split=6, cfp32=True)))
    cs.append(('bf16_dgrad_KMN', lambda: gemm_case(128 * 150, 512, 3072, bf, 0, 1, cfp32=False, bias=False)))
    cs.append(('tf32_KK_small', lambda: gemm_case(128, 128, 32, f32, 0, 0, cfp32=True, bias=False)))
    cs.append(('tf32_KK_k512', lambda: gemm_case(512, 512, 512, f32, 0, 0, cfp32=True)))
    cs.append(('tf32_MNMN', lambda: gemm_case(256, 256, 256, f32, 1, 1, cfp32=True, bias=False)))
    cs.append(('tf32_KMN', lambda: gemm_case(256, 256, 256, f32, 0, 1, cfp32=True, bias=False)))

    def qkv_case():
        T, L = 1024, 512
        x = rnd(T, 512)
        w = (rnd(3072, 512).float() * 512 ** -0.5).to(bf)
        b = rnd(3072, dtype=f32) * 0.1
        qw = 1 + 0.1 * rnd(64, dtype=f32)
        kw = 1 + 0.1 * rnd(64, dtype=f32)
        rope = lib.rope_table(L, dev)
        raw = torch.empty(T, 3072, dtype=bf, device=dev)
        out = lib.qkv_proj(x, w, b, qw, kw, rope, L, raw_out=raw)
        torch.cuda.synchronize()
        ref = x.float() @ w.float().t() + b
        e_raw = rel(raw, ref)
        q, k, v = ref.view(T, 3, 16, 64).unbind(1)
        eps = torch.finfo(torch.float32).eps

        def nr(t, wt):
            t = t * (t.pow(2).mean(-1, keepdim=True) + eps).rsqrt() * wt
            pos = (torch.arange(T, device=dev) % L).float()
            inv = (10000 ** (torch.arange(0, 64, 2).float() / -64)).to(dev)
            fr = torch.outer(pos, inv)[:, None, :]
            t1, t2 = t.chunk(2, -1)
            return torch.cat([t1 * fr.cos() - t2 * fr.sin(), t1 * fr.sin() + t2 * fr.cos()], -1)
        refo = torch.stack([nr(q, qw), nr(k, kw), v], 1).reshape(T, 3072)
        return max(e_raw, rel(out, refo))
    cs.append(('qkv_epilogue', qkv_case))

    def attn_case(B, L):
        def f():
            qkv = rnd(B * L, 3072)
            y, lse = lib.attn_fwd(qkv, B, L)
            torch.cuda.synchronize()
            q, k, v = qkv.float().view(B, L, 3, 16, 64).permute(2, 0, 3, 1, 4)
            sc = (q @ k.transpose(-1, -2)) / 8
            ref = (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, 1024)
            e2 = float((lse - torch.logsumexp(sc, -1)).abs().max())
            return max(rel(y, ref), e2)
        return f
    cs.append(('attn_B1_L128', attn_case(1, 128)))
    cs.append(('attn_B2_L512', attn_case(2, 512)))
    cs.append(('attn_B1_L200', attn_case(1, 200)))
    cs.append(('attn_B1_L4096', attn_case(1, 4096)))

    def attn_bound_case(B, L, variant):
        def f():
            qkv = rnd(B * L, 3072)
            bound = torch.tensor([14.0], device=dev)
            y, lse = lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=variant)
            torch.cuda.synchronize()
            q, k, v = qkv.float().view(B, L, 3, 16, 64).permute(2, 0, 3, 1, 4)
            sc = (q @ k.transpose(-1, -2)) / 8
            ref = (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, 1024)
            e2 = float((lse - torch.logsumexp(sc, -1)).abs().max())
            return max(rel(y, ref), e2)
        return f
    for v in (0, 1, 2, 3, 4):
        cs.append((f'attn_fixed_v{v}_L200', attn_bound_case(1, 200, v)))
        cs.append((f'attn_fixed_v{v}_L1024', attn_bound_case(2, 1024, v)))
        cs.append((f'attn_online_v{v}_L1000', (lambda vv: (lambda: attn_case_v(1, 1000, vv)))(v)))

    def attn_case_v(B, L, variant):
        qkv = rnd(B * L, 3072)
        y, lse = lib.attn_fwd(qkv, B, L, variant=variant)
        torch.cuda.synchronize()
        q, k, v = qkv.float().view(B, L, 3, 16, 64).permute(2, 0, 3, 1, 4)
        sc = (q @ k.transpose(-1, -2)) / 8
        ref = (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, 1024)
        return max(rel(y, ref), float((lse - torch.logsumexp(sc, -1)).abs().max()))

    def attn_speed(variant, fixed):
        def f():
            B, L = 8, 8192
            qkv = rnd(B * L, 3072)
            bound = torch.tensor([14.0], device=dev) if fixed else None
            for _ in range(2):
                lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=variant)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=variant)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            return {'ms': ms, 'tflops': 4 * B * 16 * L * L * 64 / ms / 1e9}
        return f
    for v in (2, 4):
        for fx in (0, 1):
            cs.append((f'attn_speed_B8_L8192_v{v}_fixed{fx}', attn_speed(v, fx)))

    def attn_bwd_speed():
        B, L = 8, 8192
        qkv = rnd(B * L, 3072)
        dy = rnd(B * L, 1024)
        y, lse = lib.attn_fwd(qkv, B, L)
        lib.attn_bwd(qkv, y, dy, lse, B, L)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            lib.attn_bwd(qkv, y, dy, lse, B, L)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        return {'ms': ms, 'tflops_algorithmic': 8 * B * 16 * L * L * 64 / ms / 1e9}
    cs.append(('attn_bwd_speed_B8_L8192', attn_bwd_speed))

    def gemm_speed():
        M, N, K = 131072, 3072, 512
        A, Bm = rnd(M, K), rnd(N, K)
        C = torch.empty(M, N, dtype=bf, device=dev)
        for _ in range(2):
            lib.gemm(A, Bm, C)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            lib.gemm(A, Bm, C)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        return {'ms': ms, 'tflops': 2 * M * N * K / ms / 1e9}
    cs.append(('gemm_speed_qkv_shape', gemm_speed))

    def timeit(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def speed_qkv_epi():
        T, L = 131072, 8192
        x = rnd(T, 512)
        w = rnd(3072, 512)
        b = rnd(3072, dtype=f32)
        qw, kw = rnd(64, dtype=f32), rnd(64, dtype=f32)
        rope = lib.rope_table(L, dev)
        raw = torch.empty(T, 3072, dtype=bf, device=dev)
        ms1 = timeit(lambda: lib.qkv_proj(x, w, b, qw, kw, rope, L, raw_out=raw))
        ms0 = timeit(lambda: lib.qkv_proj(x, w, b, qw, kw, rope, L))
        return {'ms_with_raw': ms1, 'ms_no_raw': ms0, 'tflops_with_raw': 2 * T * 3072 * 512 / ms1 / 1e9}
    cs.append(('speed_qkv_epilogue', speed_qkv_epi))

    def speed_wgrad():
        T = 131072
        dy, x = rnd(T, 3072), rnd(T, 512)
        C = torch.zeros(3072, 512, dtype=f32, device=dev)
        ms = timeit(lambda: lib.gemm(dy, x, C, a_major=1, b_major=1, epi=2, split_k=12))
        return {'ms': ms, 'tflops': 2 * T * 3072 * 512 / ms / 1e9}
    cs.append(('speed_wgrad_qkv', speed_wgrad))

    def speed_dgrad():
        T = 131072
        dy, w = rnd(T, 3072), rnd(3072, 512)
        C = torch.empty(T, 512, dtype=bf, device=dev)
        ms = timeit(lambda: lib.gemm(dy, w, C, b_major=1))
        return {'ms': ms, 'tflops': 2 * T * 3072 * 512 / ms / 1e9}
    cs.append(('speed_dgrad_qkv', speed_dgrad))

    def speed_f32out():
        T = 131072
        y, w = rnd(T, 1024), rnd(512, 1024)
        b = rnd(512, dtype=f32)
        C = torch.empty(T, 512, dtype=f32, device=dev)
        ms = timeit(lambda: lib.gemm(y, w, C, bias=b))
        return {'ms': ms, 'tflops': 2 * T * 512 * 1024 / ms / 1e9}
    cs.append(('speed_outproj_f32', speed_f32out))
    return cs


def worker(start):
    import torch
    cs = cases()
    res_path = os.path.join(OUT, 'probe.json')
    res = json.load(open(res_path)) if os.path.exists(res_path) else {}
    for i in range(start, len(cs)):
        name, fn = cs[i]
        t0 = time.time()
        try:
            err = fn()
            res[name] = {'rel_err': err, 's': round(time.time() - t0, 2)}
        except Exception as e:  # noqa
            res[name] = {'error': repr(e)[:400]}
            json.dump(res, open(res_path, 'w'), indent=1)
            print(name, res[name], flush=True)
            sys.exit(100 + i)  # sticky CUDA error likely: restart after this case
        print(name, res[name], flush=True)
        json.dump(res, open(res_path, 'w'), indent=1)
    sys.exit(0)


def main():
    os.makedirs(OUT, exist_ok=True)
    if '--from' in sys.argv:
        worker(int(sys.argv[sys.argv.index('--from') + 1]))
        return
    res_path = os.path.join(OUT, 'probe.json')
    if os.path.exists(res_path):
        os.remove(res_path)
    start = 0
    for _ in range(40):
        try:
            r = subprocess.run([sys.executable, __file__, '--from', str(start)], timeout=300)
            code = r.returncode
        except subprocess.TimeoutExpired:
            print('worker timeout at', start, flush=True)
            code = 100 + start
            res = json.load(open(res_path)) if os.path.exists(res_path) else {}
            res[f'case_{start}_timeout'] = True
            json.dump(res, open(res_path, 'w'), indent=1)
        if code == 0:
            break
        if code >= 100:
            start = code - 100 + 1
        else:
            start += 1
    print(open(res_path).read())


if __name__ == '__main__':
    main()
