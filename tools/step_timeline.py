"""CUPTI timeline of LIVE training steps at the bench shape (nsys is not in the image; torch.profiler's CUDA activity
records every kernel of the process, including the ones libosd_b200.so launches through ctypes).

Answers VERDICT r01 'prove or drop "power-bound"': for each profiled step, the span from the first kernel start to the last
kernel end, the sum of kernel durations inside it, the idle gaps between kernels, and the in-step duration of every kernel
name -- next to the same attention kernels timed ALONE in the same process, and the SM clock / power trace of the run.

    python tools/step_timeline.py [B] [L] [steps] > gpurun_out/step_timeline.json
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from oracle import denoiser_oracle as O  # noqa: E402  (seeded synthetic weights only)
from osu_dreamer_b200 import lib  # noqa: E402
from osu_dreamer_b200.denoiser import default_args  # noqa: E402
from osu_dreamer_b200.trainer import DiffusionTrainer, LRScheduleArgs  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device('cuda', 0)
    tr = DiffusionTrainer(val_batches=8, opt_args=dict(lr=3e-4, weight_decay=0.01),
                          schedule_args=LRScheduleArgs(warmup_steps=1000, warmup_init=0.3, decay_start=30000),
                          osl_weight=1.0, del_weight=30.0, emb_dim=6, a_dim=128, style_dim=32, diffusion_args=default_args())
    sd = O.make_state_dict(1234)
    tr.diffusion.load_state_dict(sd)
    tr.diffusion_ema.module.load_state_dict(sd)
    tr = tr.to(dev)
    inp = O.make_inputs(B, L, seed=7)
    batch = (inp['h'].to(dev), inp['x1'].to(dev), inp['s'].to(dev), torch.zeros(B, 5, device=dev))
    for _ in range(4):
        tr.training_step(batch)
    torch.cuda.synchronize()
    # un-profiled reference time of the same steps (events), with the clock sampler running
    clocks = bench.ClockSampler(0)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tr.training_step(batch)
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms_plain = e0.elapsed_time(e1) / 10
    marks = []
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            tr.training_step(batch)
            torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'trace.json')
        prof.export_chrome_trace(path)
        trace = json.load(open(path))
    ks = [(e['ts'], e['dur'], e['name']) for e in trace['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy') and 'dur' in e]
    ks.sort()
    # split into steps at the largest gaps (each step ends with a host synchronize)
    gaps = sorted(((ks[i + 1][0] - (ks[i][0] + ks[i][1]), i) for i in range(len(ks) - 1)), reverse=True)[:steps - 1]
    cuts = sorted(i for _, i in gaps)
    bounds = [0] + [c + 1 for c in cuts] + [len(ks)]
    per_step = []
    agg = {}
    for a, b in zip(bounds[:-1], bounds[1:]):
        seg = ks[a:b]
        span = seg[-1][0] + seg[-1][1] - seg[0][0]
        busy, end, idle_gaps, big = 0.0, seg[0][0], 0.0, 0
        for ts, dur, name in seg:
            if ts > end:
                idle_gaps += ts - end
                big += (ts - end) > 5.0
            busy += dur
            end = max(end, ts + dur)
            short = name.split('(')[0].replace('void ', '').replace('osd::', '')[:60]
            d = agg.setdefault(short, [0, 0.0])
            d[0] += 1
            d[1] += dur
        per_step.append({'launches': len(seg), 'span_ms': span / 1e3, 'sum_kernel_ms': busy / 1e3, 'idle_gap_ms': idle_gaps / 1e3,
                         'gaps_over_5us': int(big), 'gap_frac': idle_gaps / span})
    # attention kernels alone (same process, warm GPU)
    qkv = torch.randn(B * L, 3072, device=dev).to(torch.bfloat16)
    y, lse = lib.attn_fwd(qkv, B, L)
    dy = torch.randn(B * L, 1024, device=dev).to(torch.bfloat16)
    bound = torch.tensor([14.0], device=dev)
    alone = {'attn_fwd_ms': bench.time_kernel(lambda: lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=18), iters=5, warm=2),
             'attn_bwd_fused_ms': bench.time_kernel(lambda: lib.attn_bwd_fused(qkv, y, dy, lse, B, L), iters=5, warm=2)}
    top = sorted(((v[1] / steps / 1e3, v[0] // steps, k) for k, v in agg.items()), reverse=True)
    fl = 4.0 * B * 16 * L * L * 64
    in_step = {}
    for ms, n, k in top:
        if 'attn_fwd_db_kernel' in k or 'attn_fwd_pp3_kernel' in k:
            in_step['attn_fwd'] = {'ms_per_launch': ms / n, 'tflops': fl / (ms / n) / 1e9}
        if 'attn_bwd_fused_kernel' in k:
            in_step['attn_bwd_fused'] = {'ms_per_launch': ms / n, 'tflops': 2 * fl / (ms / n) / 1e9}
    out = {'shape': {'B': B, 'L': L}, 'step_ms_unprofiled_events': ms_plain, 'clocks_during_unprofiled_steps': clk,
           'profiled_steps': per_step,
           'kernels_in_step_ms_per_step': [{'kernel': k, 'launches': n, 'ms': round(ms, 4)} for ms, n, k in top[:40]],
           'attention_in_step': in_step, 'attention_alone_same_process': alone,
           'method': 'torch.profiler (CUPTI activity records) over whole training steps; span = first kernel start .. last kernel end; '
                     'idle gaps = time inside the span with no kernel running'}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
