"""Turns the ncu output of a gpurun job into the tracked summaries under profiles/.

    python scratch/summarize_ncu.py <tag> <launch_list.csv> <full.ncu-rep>

writes profiles/<tag>_launches.csv (the launch list as captured), profiles/<tag>_launches_summary.md
(per kernel: launches, time, share, DRAM / L2 bytes per launch), profiles/<tag>_ncu_full_raw.csv
(`ncu --page raw --csv` of the full capture), profiles/<tag>_ncu_full_summary.md (the metrics the
roofline discussion in DESIGN.md quotes) and profiles/<tag>_chain_traffic.json (DRAM bytes per
launch by C entry point: bench.py's `roofline.traffic`)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENTRY = {'tc_mlp_train_kernel': 'tb_tc_mlp_train', 'tc_wgrad_all_kernel': 'tb_mlp_wgrad_fused',
         'tc_mlp_forward_kernel': 'tb_tc_mlp_forward', 'tc_mlp_backward_kernel': 'tb_tc_mlp_backward'}


def read_csv_after_banner(path):
    lines = open(path).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    return list(csv.reader(lines[start:]))


def short(name):
    return name.split('(')[0].replace('void ', '').replace('tb::', '').strip()


def launches(tag, path):
    rows = read_csv_after_banner(path)
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        key = r[col['ID']]
        d = per.setdefault(key, dict(kernel=short(r[col['Kernel Name']]), grid=r[col['Grid Size']],
                                     block=r[col['Block Size']]))
        value = float(r[col['Metric Value']].replace(',', ''))
        unit = r[col['Metric Unit']]
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
        d[r[col['Metric Name']]] = value * scale
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d['kernel'], dict(n=0, us=0.0, rd=0.0, wr=0.0, l2=0.0, grid=d['grid'], block=d['block']))
        a['n'] += 1
        a['us'] += d.get('gpu__time_duration.sum', 0.0)
        a['rd'] += d.get('dram__bytes_read.sum', 0.0)
        a['wr'] += d.get('dram__bytes_write.sum', 0.0)
        a['l2'] += d.get('lts__t_bytes.sum', 0.0)
    total = sum(a['us'] for a in agg.values())
    out = [f'# {tag}: launch list of one PPO iteration with one 16384-row minibatch per network',
           '', 'Command: see scratch/gpu_job*.sh (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,'
           'dram__bytes_write.sum,lts__t_bytes.sum --clock-control none, graphs off, cudaProfilerStart/Stop '
           'around the iteration).  Times under ncu are serialised single-launch times: use the SHARES.', '',
           '| kernel | launches | grid | block | us / launch | share | DRAM rd MB / launch | DRAM wr MB / launch | L2 MB / launch |',
           '|---|---|---|---|---|---|---|---|---|']
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        n = a['n']
        out.append(f"| {k} | {n} | {a['grid']} | {a['block']} | {a['us'] / n:.2f} | {a['us'] / total:.3f} | "
                   f"{a['rd'] / n / 1e6:.2f} | {a['wr'] / n / 1e6:.2f} | {a['l2'] / n / 1e6:.2f} |")
    out.append('')
    out.append(f'total kernel time {total:.1f} us over {sum(a["n"] for a in agg.values())} launches')
    open(os.path.join(ROOT, 'profiles', f'{tag}_launches_summary.md'), 'w').write('\n'.join(out) + '\n')
    with open(os.path.join(ROOT, 'profiles', f'{tag}_launches.csv'), 'w') as f:
        csv.writer(f).writerows(rows)
    traffic = {}
    for k, a in agg.items():
        base = k.split('<')[0]
        if base in ENTRY:
            traffic.setdefault(ENTRY[base], []).append((a['rd'] + a['wr']) / a['n'])
    return {k: round(sum(v) / len(v)) for k, v in traffic.items()}


METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
    'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_uniform.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
]


def full(tag, rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_full_raw.csv'), 'w').write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f'# {tag}: ncu --set full of the two tensor-core kernels of the update chain (16384 rows)', '',
           'Raw page: ' + f'{tag}_ncu_full_raw.csv' + ' (same directory).  Metric names as ncu prints them; where a metric has',
           'several collection sections the first match is listed.', '']
    for r in rows[2:]:
        name = short(r[hdr.index('Kernel Name')])
        out += [f'## {name}  (launch id {r[0]}, grid {r[hdr.index("Grid Size")]}, block {r[hdr.index("Block Size")]})', '',
                '| metric | value | unit |', '|---|---|---|']
        for m in METRICS:
            hit = [i for i, h in enumerate(hdr) if h == m or h.endswith('.' + m)]
            if hit:
                out.append(f'| {m} | {r[hit[0]]} | {units[hit[0]]} |')
        out.append('')
    open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_full_summary.md'), 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    tag, launch_csv, rep = sys.argv[1:4]
    traffic = launches(tag, launch_csv)
    full(tag, rep)
    json.dump(traffic, open(os.path.join(ROOT, 'profiles', f'{tag}_chain_traffic.json'), 'w'), indent=1)
    print(traffic)
