"""Turn the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01      # reads gpurun_out/launches_r01.csv, prof_kernels_r01.ncu-rep, microbench_r01.jsonl
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
os.makedirs(P, exist_ok=True)

# ---- launch list: share of every kernel in the capture -------------------------------------------------------
src = os.path.join(G, 'launches_%s.csv' % tag)
if os.path.isfile(src):
    lines = [l for l in open(src) if l.startswith('"')]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1e3 if row['Metric Unit'] in ('nsecond', 'ns') else (v * 1e3 if row['Metric Unit'] in ('msecond', 'ms') else v)
        name = re.sub(r'^void ', '', re.sub(r'\(.*', '', row['Kernel Name']))[:100]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(os.path.join(P, '%s_launches_summary.txt' % tag), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline\n')
        f.write('# %d launches, %.1f ms of kernel time in the capture (5 training steps + set-up); times are cold-cache and serialised:\n'
                '# compare SHARES.  us = microseconds.\n' % (sum(cnt.values()), T / 1e3))
        mine = sorted((k for k in tot if k.startswith('bh::')), key=lambda k: -tot[k])
        f.write('\n## custom kernels (libbihome_b200.so)\n')
        for k in mine:
            f.write('%-60s n=%5d total %10.1f us  avg %9.2f us  share %6.3f %%\n' % (k, cnt[k], tot[k], tot[k] / cnt[k], 100 * tot[k] / T))
        f.write('custom kernels together: %.2f %% of GPU time\n' % (100 * sum(tot[k] for k in mine) / T))
        f.write('\n## top 30 overall (cuDNN / ATen = backbone + frozen extractor + Adam)\n')
        for k, v in tot.most_common(30):
            f.write('%-100s n=%5d total %10.1f us  share %6.2f %%\n' % (k, cnt[k], v, 100 * v / T))
    print('wrote launches summary')

# ---- ncu --set full: a few metrics per captured kernel --------------------------------------------------------
rep = os.path.join(G, 'prof_kernels_%s.ncu-rep' % tag)
raw_csv = os.path.join(G, 'prof_kernels_raw_%s.csv' % tag)
if os.path.isfile(rep) or os.path.isfile(raw_csv):
    if os.path.isfile(raw_csv) and os.path.getsize(raw_csv) > 1000:      # exported on the GPU box (tools/gpu_round.sh)
        raw = open(raw_csv).read()
    else:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct' ]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(P, '%s_ncu_kernels.csv' % tag), 'w') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([re.sub(r'\(.*', '', r[i]) if hdr[i] == 'Kernel Name' else r[i] for i in idx])
    print('wrote ncu kernel summary')
    # dram traffic of the loss launch at B=256 (stream + finish kernels), for bench.py's roofline.traffic
    ik, ir, iw = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    per = {}
    for r in rows[2:]:
        for key in ('bihome_stream_kernel', 'bihome_finish_kernel'):
            if key in r[ik]:
                per.setdefault(key, []).append(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
    if 'bihome_stream_kernel' in per:
        total = sum(sum(v) / len(v) for v in per.values())
        with open(os.path.join(P, 'loss_traffic.json'), 'w') as f:
            json.dump({'dram_bytes_per_launch': total, 'kernels': {k: sum(v) / len(v) for k, v in per.items()},
                       'source': 'ncu --set full, %s, tools/microbench.py --once (B=256, C=64, h=w=32, channels-last)' % os.path.basename(rep)}, f, indent=1)
        print('wrote loss_traffic.json', total)

    # dram traffic per launch of the other entry points bench.py may name in its roofline block (first captured geometry of
    # each kernel = the B = 256 north-star shape of tools/microbench.py --once)
    ENTRY = {'bh_bihome_fwd_bwd': ('bihome_stream_kernel', 'bihome_finish_kernel'),
             'bh_fieldhead_bwd': ('fieldhead_gx_mma_kernel', 'fieldhead_gw_mma_kernel'), 'bh_fieldhead_fwd': ('fieldhead_fwd_mma_kernel',),
             'bh_fieldhead_moments': ('moments_mma_kernel',), 'bh_fieldhead_affine': ('affine_acc_kernel',),
             'bh_warp_fwd': ('warp_fwd_tile_kernel',), 'bh_warp_bwd': ('warp_bwd_tile_kernel', 'warp_bwd_finish_kernel'),
             'bh_pairgen_apply': ('pairgen_apply_kernel',)}
    ig = hdr.index('launch__grid_size')
    first = {}
    for r in rows[2:]:
        for keys in ENTRY.values():
            for key in keys:
                if key in r[ik]:
                    geo = first.setdefault(key, (r[ig], []))
                    if geo[0] == r[ig]:
                        geo[1].append(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
    traffic = {}
    for entry, keys in ENTRY.items():
        if all(k in first for k in keys):
            traffic[entry] = sum(sum(first[k][1]) / len(first[k][1]) for k in keys)
    if traffic:
        traffic['source'] = 'ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum), %s, tools/microbench.py --once, B = 256' % tag
        with open(os.path.join(P, 'kernel_traffic.json'), 'w') as f:
            json.dump(traffic, f, indent=1)
        print('wrote kernel_traffic.json', traffic)

mb = os.path.join(G, 'microbench_%s.jsonl' % tag)
if os.path.isfile(mb):
    with open(os.path.join(P, '%s_microbench.txt' % tag), 'w') as f:
        f.write('# python tools/microbench.py  (CUDA events, L2 flushed before every launch, mean of 20; peak = measured copy bandwidth)\n')
        for l in open(mb):
            try:
                d = json.loads(l)
            except Exception:  # noqa: BLE001
                continue
            if 'kernel' in d:
                f.write('%-34s B=%-5s P=%-4s C=%-4s %9.4f ms %9.1f MB %8.1f GB/s  frac %.3f\n' % (
                    d['kernel'], d.get('B'), d.get('P', '-'), d.get('C', '-'), d['ms'], d['algorithmic_MB'], d['GB/s'], d['frac_of_measured_peak']))
            else:
                f.write('# %s\n' % json.dumps(d))
    print('wrote microbench summary')
