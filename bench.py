#!/usr/bin/env python
"""bench.py -- denoiser training-step throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one full fit-denoiser training step (reference: DiffusionTrainer.training_step + Lightning's
clip / AdamW / LambdaLR / EMA, osu_dreamer/models/diffusion/train.py:69-126, model.yml:39) on one batch of
synthetic data at BASELINE.json configs[1]: batch 16 per GPU, seq_len 8192, 128 audio-feature channels, bf16
tensor-core operands (fp32 accumulate / residual / statistics).  Rank 0 prints ONE JSON line.

`--impl reference` times the reference's own CPU arithmetic for the same step (the oracle port of the
reference's DiffusionModel / trainer loss, torch CPU, all host threads) on a bounded sample -- there is no
GPU work on that arm.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ = 8192
BATCH_PER_GPU = 16


def f_fwd(L: int) -> float:
    """algorithmic forward FLOPs per sample."""
    return L * (8 * (8523776 + 4096 * L) + 54436)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smmax, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_train_step_seconds(L: int, B: int, iters: int, warmup: int, threads: int):
    """the oracle port of the reference's training arithmetic on the host cores: loss fwd + autograd bwd +
    clip + AdamW + EMA on a bounded shape."""
    import torch
    from oracle import denoiser_oracle as O
    torch.set_num_threads(threads)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.make_state_dict(1234).items()}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v), v.detach().clone()) for k, v in sd.items()}
    inp = O.make_inputs(B, L, seed=7)
    times = []
    for it in range(warmup + iters):
        t0 = time.perf_counter()
        for v in sd.values():
            v.grad = None
        loss, _ = O.trainer_loss(sd, inp['h'], inp['x1'], inp['s'], inp['x0'], inp['t'])
        loss.backward()
        gn = math.sqrt(sum(float(v.grad.double().pow(2).sum()) for v in sd.values()))
        coef = min(1.0, 1.0 / (gn + 1e-6))
        with torch.no_grad():
            for k, v in sd.items():
                m, vv, ema = state[k]
                O.adamw_ema_step(v, v.grad, m, vv, ema, it + 1, 3e-4 * O.lr_lambda(it), clip_coef=coef, ema_first=(it == 0))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    # torch's CPU kernels stop scaling (and on a shared 128-thread host get much slower) beyond a few dozen
    # threads: use up to 32.  Shape = BASELINE configs[0], the reference's own CPU-runnable case (B=2, L=512).
    threads = min(os.cpu_count() or 1, 32)
    Ls, Bs = 512, 2
    sec = cpu_train_step_seconds(Ls, Bs, max(1, min(args.steps, 5)), max(0, min(args.warmup, 1)), threads)
    # samples/s measured at L=512; the same arithmetic at L=8192 costs F(8192)/F(512) more per sample
    v_sample = Bs / sec
    scale = f_fwd(Ls) / f_fwd(SEQ)
    value = v_sample * scale
    sample = (f'oracle port (torch CPU fp32, {threads} threads) of the reference train step at B={Bs}, L={Ls}: '
              f'{v_sample:.4f} samples/s measured, scaled by F_fwd({Ls})/F_fwd({SEQ})={scale:.4f} to the L={SEQ} workload '
              f'(the reference cannot run L={SEQ} training on CPU: 8 x 4.3 GB of saved attention scores per sample)')
    line = {
        'impl': 'reference', 'metric': 'denoiser train samples/sec @ seq8192', 'value': value, 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'fit-denoiser train step, batch {BATCH_PER_GPU}/GPU, seq_len {SEQ}, 128 audio channels '
                               f'(CPU arm: bounded sample, see cpu_baseline.sample)'},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ CUDA arm
def time_kernel(fn, iters=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from osu_dreamer_b200 import lib
    from osu_dreamer_b200.denoiser import default_args
    from osu_dreamer_b200.trainer import DiffusionTrainer, LRScheduleArgs
    from oracle import denoiser_oracle as O  # only for the seeded synthetic weights/inputs + cpu_baseline leg

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus or world == 1, f'WORLD_SIZE {world} != --gpus {args.gpus}'
    lib.load()

    B, L = BATCH_PER_GPU, SEQ
    tr = DiffusionTrainer(val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
                          schedule_args=LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
                          osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32,
                          diffusion_args=default_args())
    sd = O.make_state_dict(1234)  # identical non-degenerate weights on every rank
    tr.diffusion.load_state_dict(sd)
    tr.diffusion_ema.module.load_state_dict(sd)
    tr = tr.to(dev)
    g = torch.Generator().manual_seed(1000 + rank)
    h_host = torch.randn(B, 128, L, generator=g).pin_memory()
    x1 = torch.randn(B, 6, L, generator=g)
    x1_host = (x1 * x1.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()).pin_memory()
    s = torch.randn(B, 32, generator=g)
    s_host = (s * s.pow(2).mean(1, keepdim=True).add(1e-6).rsqrt()).pin_memory()
    lab_host = torch.zeros(B, 5).pin_memory()
    batch_dev = tuple(t.to(dev) for t in (h_host, x1_host, s_host, lab_host))
    torch.manual_seed(1234 + rank)

    def step_resident():
        return tr.training_step(batch_dev, world_size=world)

    def step_e2e():
        batch = tuple(t.to(dev, non_blocking=True) for t in (h_host, x1_host, s_host, lab_host))
        loss, _ = tr.training_step(batch, world_size=world)
        return float(loss)  # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / 1e3, wall

    for _ in range(max(3, args.warmup)):
        step_resident()
    n0 = lib.launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    sec, _ = timed(step_resident, args.steps)
    clk = clocks.stop() if rank == 0 else None
    launches = (lib.launch_count() - n0) // max(1, args.steps)
    step_e2e()
    sec_e2e, wall_e2e = timed(step_e2e, args.steps)

    value = world * B * args.steps / sec
    e2e_value = world * B * args.steps / max(sec_e2e, wall_e2e if world == 1 else sec_e2e)
    peaks = load_peaks()
    step_tf = 3 * f_fwd(L) * B / (sec / args.steps) / 1e12  # per GPU, algorithmic (fwd + dgrad + wgrad)

    line = {
        'metric': 'denoiser train samples/sec @ seq8192', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': sec / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': f'fit-denoiser train step (fwd + bwd + NCCL grad allreduce + clip + AdamW + EMA), '
                               f'batch {B}/GPU, seq_len {L}, 128 audio channels, 46.9M params (BASELINE configs[1]'
                               f'{"; weak-scaled DDP" if world > 1 else ""})',
                   'global_batch': world * B, 'seq_len': L, 'parallelism': f'dp{world}',
                   'l2': 'per-step activations (>30 GB) far exceed the 126 MB L2; no explicit flush'},
        'e2e': {'value': e2e_value, 'unit': 'samples/s',
                'h2d_bytes_per_step': int(sum(t.numel() * 4 for t in (h_host, x1_host, s_host, lab_host))),
                'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'step_algorithmic_tflops_per_gpu': step_tf,
        'step_tensor_frac_of_sustained_peak': step_tf / peaks['tf_sustained'],
    }
    if rank == 0:
        line['clocks'] = clk

    # ---- roofline of the dominant kernels (attention), timed alone with CUDA events on the launch stream
    if rank == 0:
        try:
            Ba = B
            qkv = torch.randn(Ba * L, 3072, device=dev).to(torch.bfloat16)
            y, lse = lib.attn_fwd(qkv, Ba, L)
            dy = torch.randn(Ba * L, 1024, device=dev).to(torch.bfloat16)
            bound = torch.tensor([14.0], device=dev)  # randn scores / 8 stay far below 2^14: same kernel path the model runs
            ms_f = time_kernel(lambda: lib.attn_fwd(qkv, Ba, L, bound_log2=bound, variant=7), iters=3, warm=1)  # the model's variant
            ms_b = time_kernel(lambda: lib.attn_bwd_fused(qkv, y, dy, lse, Ba, L), iters=3, warm=1)
            fl_f = 4.0 * Ba * 16 * L * L * 64
            kern = {'attn_fwd': {'ms': ms_f, 'tflops': fl_f / ms_f / 1e9},
                    'attn_bwd_fused': {'ms': ms_b, 'tflops': 2 * fl_f / ms_b / 1e9}}
            dom = 'attn_bwd_fused' if 8 * ms_b > 8 * ms_f else 'attn_fwd'
            ach = kern[dom]['tflops']
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this exact shape (B=16, L=8192), from the
            # latest committed ncu --set full capture (profiles/ncu_attention_summary_latest.json, written by
            # tools/ncu_summary.py from tools/gpu_round.sh's capture of tools/prof_attn.py 16 8192)
            ncu_traffic = {'attn_bwd_fused': 1.622335e9 + 1.030786e9, 'attn_fwd': 0.805537e9 + 0.259025e9}
            ncu_src = 'profiles/r01h_ncu_attention_summary.json'
            try:
                latest = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'ncu_attention_summary_latest.json')
                gb = lambda v: float(v.split()[0]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[v.split()[1]]
                for kk in json.load(open(latest))['kernels']:
                    tot = gb(kk['dram__bytes_read.sum']) + gb(kk['dram__bytes_write.sum'])
                    if 'attn_bwd_fused_kernel' in kk['Kernel Name']:
                        ncu_traffic['attn_bwd_fused'] = tot
                    elif 'attn_fwd_db_kernel' in kk['Kernel Name']:
                        ncu_traffic['attn_fwd'] = tot
                ncu_src = 'profiles/ncu_attention_summary_latest.json'
            except Exception:  # noqa: keep the constants of the r01h capture
                pass
            alg_bytes = {'attn_bwd_fused': Ba * L * (3072 * 2 + 1024 * 2 + 2048 * 2 + 1024 * 4),  # q,k,v + dO~ in; dk,dv + fp32 dq out
                         'attn_fwd': Ba * L * (3072 * 2 + 1024 * 2)}
            line['roofline'] = {'bound': 'tensor', 'kernel': dom, 'achieved': ach, 'peak': peaks['tf_burst'],
                                'unit': 'TFLOP/s', 'frac': ach / peaks['tf_burst'],
                                'traffic': ncu_traffic[dom] if (Ba, L) == (16, 8192) else None,
                                'peak_source': peaks['source'] + ', burst (kernel timed alone)',
                                'traffic_note': f'DRAM bytes per launch from ncu ({ncu_src}) vs '
                                                f'{alg_bytes[dom] / 1e9:.2f} GB algorithmic: the fp32 dQ accumulator is partly evicted and '
                                                're-read between its TMA reduce-adds (1.2x); K/V/Q/dO re-reads are served by L2',
                                'flop_convention': 'algorithmic = 2 x forward (SURVEY 8(d)); the single-pass kernel executes 2.5 x forward '
                                                   '(5 GEMMs: S, dP, dV, dK, dQ), so its tensor pipe runs at 1.25 x the quoted rate',
                                'limits_note': 'd=64 attention (ncu, profiles/ncu_attention_summary_latest.json): forward SFU pipe 76 % busy '
                                               '(16384 exponentials per 128x128 tile = 1024 clk of 16-lane SFU), tensor pipe 38 %; backward tensor '
                                               'pipe 48 %, shared-memory pipe 65-80 %; TMEM->register bandwidth is NOT the bound (tools/micro/'
                                               'ldtm_bench.cu: 475-910 B/clk/SM vs 59 used); under load the B200 sits at its 1000 W power cap '
                                               '(sm clock 1.6-1.65 GHz of 1.965), see DESIGN.md 5',
                                'algorithmic_flops_per_launch': (2 * fl_f if dom != 'attn_fwd' else fl_f),
                                'kernels': kern,
                                'share_of_step': {k: 8 * v['ms'] / (sec / args.steps * 1e3) for k, v in kern.items()}}
            del qkv, y, lse, dy
        except Exception as e:  # noqa
            line['roofline'] = {'error': repr(e)[:200]}

    # ---- secondary metric: 64-step sampling latents/s (BASELINE config 3: B=32, L=8192): bf16 path and the
    #      fp32-grade path (precision='fp32': 3x-bf16 split products, 1e-3 tolerance class)
    if rank == 0 and world == 1 and not args.no_sampling:
        try:
            tr.zero_grad()
            tr.diffusion._rt.ws.clear()
            torch.cuda.empty_cache()
            Bs = 32
            m = tr.diffusion_ema.module.eval()
            hs = torch.randn(Bs, 128, L, device=dev)
            ss = torch.randn(Bs, 32, device=dev)
            line['sampling'] = {'metric': '64-step sample latents/sec @ seq8192', 'unit': 'latents/s', 'batch': Bs}
            for prec in ('bf16', 'fp32'):
                m.precision = prec
                m._rt.reset()
                torch.cuda.empty_cache()
                m.sample(hs[:2], ss[:2], 1)  # warm-up (weight packing, workspaces of this shape are re-made below)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                xs = m.sample(hs, ss, 64)
                e1.record()
                torch.cuda.synchronize()
                ssec = e0.elapsed_time(e1) / 1e3
                line['sampling'][prec] = {'value': Bs / ssec, 'seconds': ssec,
                                          'algorithmic_tflops': 65 * f_fwd(L) * Bs / ssec / 1e12,
                                          'finite': bool(torch.isfinite(xs).all())}
            m.precision = 'bf16'
            line['sampling']['value'] = line['sampling']['bf16']['value']
        except Exception as e:  # noqa
            line['sampling'] = {'error': repr(e)[:300]}

    # ---- CPU baseline (oracle port on this box's host cores), rank 0 at N=1 only, bounded sample
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = min(os.cpu_count() or 1, 32)  # torch CPU stops scaling beyond a few dozen threads
        Ls, Bs = 512, 2                          # BASELINE configs[0]: the reference's CPU-runnable case
        sec_cpu = cpu_train_step_seconds(Ls, Bs, 3, 1, threads)
        scale = f_fwd(Ls) / f_fwd(L)
        line['cpu_baseline'] = {
            'value': Bs / sec_cpu * scale, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
            'sample': f'oracle port (torch CPU fp32) train step at B={Bs}, L={Ls}: {Bs / sec_cpu:.4f} samples/s, scaled by '
                      f'F_fwd({Ls})/F_fwd({L})={scale:.5f} to seq_len {L}'}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-sampling', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == '__main__':
    main()
