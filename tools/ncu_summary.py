"""Summarise an ncu report (`ncu --set full ... -o x`) into the JSON kept under profiles/: per kernel the duration,
tensor / SFU / issue utilisation, registers, shared memory, DRAM bytes and the shared-memory wavefront counts.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json "source description" [launch-skip]"""
import csv
import io
import json
import subprocess
import sys

rep, out, desc = sys.argv[1], sys.argv[2], sys.argv[3]
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'lts__t_sectors.avg.pct_of_peak_sustained_elapsed']
idx = {h: i for i, h in enumerate(hdr)}
kernels = []
for r in rows[2 + skip:]:
    k = {'Kernel Name': r[idx['Kernel Name']]}
    for m in keep:
        if m in idx:
            k[m] = f'{r[idx[m]]} {units[idx[m]]}'.strip()
    kernels.append(k)
json.dump({'source': desc, 'kernels': kernels}, open(out, 'w'), indent=1)
print(f'{len(kernels)} kernels -> {out}')
