"""Turns what tools/collect_profiles.sh left in gpurun_out/prof/ into the tracked summaries under profiles/.

  launches.csv.gz -> profiles/<tag>_launches_b16.csv.gz (copy) + <tag>_launches_b16_summary.txt (one forward, per kernel)
  layers.txt      -> profiles/<tag>_layers_b16.txt (copy)
  ops_<name>.csv  -> profiles/<tag>_ncu_ops_summary.txt (key metrics of each representative op, one launch each)
"""
import collections
import csv
import gzip
import io
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'gpurun_out', 'prof')
DST = os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name


def launches():
    path = os.path.join(SRC, 'launches.csv.gz')
    if not os.path.exists(path):
        print('no launch list')
        return
    text = gzip.open(path, 'rt').read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    seq = []
    for r in rows:
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ms = v / 1e6 if unit in ('ns', 'nsecond') else v / 1e3 if unit in ('us', 'usecond') else v
        seq.append((short(r['Kernel Name']), ms))
    # forwards are delimited by the single argmax_gather launch each contains
    marks = [i for i, (k, _) in enumerate(seq) if 'argmax_gather' in k]
    if len(marks) < 2:
        print('cannot delimit forwards')
        return
    period = marks[-1] - marks[-2]
    one = seq[marks[-1] - period + 1: marks[-1] + 1]
    agg = collections.OrderedDict()
    for k, ms in one:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    out = ['ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` '
           '(16 clips 3x512x512).',
           'Launches of ONE forward (the last of %d in the run): %d launches, %.2f ms summed (cold-cache, serialised: compare '
           'SHARES, not absolutes)' % (len(marks), len(one), tot), '',
           '%-62s %5s %10s %7s' % ('kernel', 'n', 'ms', 'share')]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('%-62s %5d %10.3f %6.1f%%' % (k[:62], a[0], a[1], 100 * a[1] / tot))
    tc = sum(a[1] for k, a in agg.items() if re.search(r'gemm_tc|conv_halo|swin_mlp|rgb_conv|ln_linear', k))
    out.append('')
    out.append('tcgen05 GEMM / conv class (gemm_tc, conv_halo, conv_halo2, swin_mlp, ln_linear, rgb_conv): %.2f ms = %.1f%% of the '
               'launch-list time' % (tc, 100 * tc / tot))
    open(os.path.join(DST, '%s_launches_b16_summary.txt' % tag), 'w').write('\n'.join(out) + '\n')
    shutil.copy(path, os.path.join(DST, '%s_launches_b16.csv.gz' % tag))
    print('\n'.join(out[:14]))


KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__cluster_size', 'cluster'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/smem throughput %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor operand-feed active %'),
    ('sm__inst_issued.avg.per_cycle_active', 'issued IPC / SM'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('sm__cycles_elapsed.max', 'SM cycles'),
]

KEYS += [('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
         ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %')]

OPS = [('window_tc', 'window_attn_tc_kernel<32>: 4 clips x 1024 shifted windows of 48 tokens, C=256, 8 heads (core only: HBM-bound, 24 FLOP/B)'),
       ('mha_tc', 'mha_tc_kernel: 4 clips, L=3072, 8 heads x d=64 (flash attention, O / L in TMEM)'),
       ('argmax', 'argmax_gather_kernel: T=49152 rows x 1024 fp32 logits + gather of 512-float codes'),
       ('l2_argmin', 'l2_argmin_pair_kernel (after z_pack_kernel): T=49152 tokens x 1024 codes x 512, random z (bf16 tensor-core scores on CTA pairs + certified window + exact resolution)'),
       ('ln_linear', 'ln_linear_kernel: LayerNorm + q/kv projection, T=196608, C=256 -> 768'),
       ('conv_out', 'conv_out_gn_kernel: norm_out + SiLU + conv 64->3, 12 frames 512^2, fp32 NCHW output'),
       ('halo64', 'conv_halo_kernel<64>: 3x3, 12 frames 512^2, Cin=Cout=64, residual (weights resident)'),
       ('halo128', 'conv_halo2_kernel<128> (CTA pairs): 3x3, 12 frames 256^2, Cin=Cout=128, residual'),
       ('conv256', 'gemm_tc_kernel<256, pair>: 3x3, 12 frames 128^2, Cin=Cout=256 (K=2304), residual'),
       ('linear256', 'gemm_tc_kernel<256>: linear M=196608 N=256 K=256 + residual (HBM / L2 bound)'),
       ('linear512', 'gemm_tc_kernel<256>: linear M=49152 N=512 K=512 + bf16 residual (the 32^2-level Swin / global-transformer shape)'),
       ('linear512_f32', 'gemm_tc_kernel<256>: linear M=49152 N=512 K=512 + fp32 residual stream (global transformer out_proj / linear2)'),
       ('up128', 'conv_halo_kernel<128>: one 2x2 phase conv of the folded Upsample, 12 frames 256^2 -> 512^2, Cin=Cout=128'),
       ('swin_mlp', 'swin_mlp_kernel: LN+fc1+GELU+fc2+residual, T=196608, C=256'),
       ('rgb', 'rgb_conv_kernel<3,1,1>: conv_in 3->64 on 12 frames 512^2 with GroupNorm statistics'),
       ('gn', 'gn_apply_kernel (last of stats/finalize/apply): GroupNorm+SiLU, 12 frames 512^2 x 64')]


def ops():
    out = ['ncu --set full --clock-control none, one launch each (third call, warm), tools/ncu_ops.py <op> run in isolation.',
           'Shapes are the model\'s at 12 frames instead of 48.  Values as ncu prints them (units in brackets).', '']
    for op, title in OPS:
        path = os.path.join(SRC, 'ops_%s.csv' % op)
        if not os.path.exists(path) or os.path.getsize(path) == 0:
            continue
        rows = list(csv.reader(open(path)))
        head, units, vals = rows[0], rows[1], rows[-1]
        d = {h: (v, u) for h, u, v in zip(head, units, vals)}
        out.append('== %s' % title)
        out.append('   kernel: %s' % short(d.get('Kernel Name', ('?', ''))[0]))
        for key, label in KEYS:
            if key in d and d[key][0] != '':
                out.append('   %-32s %s [%s]' % (label, d[key][0], d[key][1]))
        out.append('')
    open(os.path.join(DST, '%s_ncu_ops_summary.txt' % tag), 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[:40]))


def traffic():
    """dram bytes per launch of the dominant (tcgen05 GEMM / conv) class over one forward -> profiles/<tag>_traffic.json"""
    import json
    path = os.path.join(SRC, 'traffic.csv.gz')
    if not os.path.exists(path):
        print('no traffic capture')
        return
    text = gzip.open(path, 'rt').read()
    start = text.index('"ID"')
    per = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO(text[start:])):
        d = per.setdefault(r['ID'], {'k': short(r['Kernel Name'])})
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        if r['Metric Name'].startswith('dram__bytes'):
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            d['bytes'] = d.get('bytes', 0.0) + v * mult
        elif r['Metric Name'] == 'gpu__time_duration.sum':
            d['ms'] = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
    seq = list(per.values())
    marks = [i for i, d in enumerate(seq) if 'argmax_gather' in d['k']]
    if len(marks) < 2:
        print('cannot delimit forwards')
        return
    one = seq[marks[-2] + 1: marks[-1] + 1]
    cls = [d for d in one if re.search(r'gemm_tc|conv_halo|swin_mlp|rgb_conv|ln_linear|conv_out_gn', d['k'])]
    tot_b = sum(d.get('bytes', 0.0) for d in cls)
    allb = sum(d.get('bytes', 0.0) for d in one)
    out = {'source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum, one forward of `bench.py --steps 1 --warmup 3` '
                     '(16 clips 3x512x512), launches of the tcgen05 GEMM / conv class',
           'launches': len(cls), 'dram_bytes_per_launch': tot_b / max(len(cls), 1), 'dram_bytes_class_per_step': tot_b,
           'dram_bytes_all_kernels_per_step': allb, 'launches_all': len(one),
           'class_ms_serialised': sum(d.get('ms', 0.0) for d in cls)}
    # algorithmic bytes of the class from the per-launch shapes of layers.txt (inputs + outputs + residual + weights, bf16)
    lay = os.path.join(SRC, 'layers.txt')
    if os.path.exists(lay):
        alg = 0.0
        n = 0
        for line in open(lay):
            m = re.match(r'\s*(\S.*?)\s+n=\s*(\d+)\s', line)
            if not m:
                continue
            desc, cnt = m.group(1), int(m.group(2))
            g = {k: int(v) for k, v in re.findall(r'\b([FHWKNM])(\d+)\b', desc)}
            if 'M' in g:
                M, N, K = g['M'], g.get('N', 0), g.get('K', 0)
                b = M * K * 2 + M * N * 2 + N * K * 2
            elif 'F' in g and 'H' in g:
                M = g['F'] * g['H'] * g['W']
                N, K = g.get('N', 0), g.get('K', 0)
                taps = 9 if desc.startswith(('halo3', 'conv3', 'conv_out')) else 4 if desc.startswith('conv2') else 1
                b = M * (K // taps) * 2 + M * N * 2 + N * K * 2 + (M * N * 2 if ' r1' in desc else 0)
            else:
                continue
            alg += b * cnt
            n += cnt
        out['algorithmic_bytes_class_per_step'] = alg
        out['algorithmic_bytes_per_launch'] = alg / max(n, 1)
    json.dump(out, open(os.path.join(DST, '%s_traffic.json' % tag), 'w'), indent=1)
    print(json.dumps(out, indent=1))


launches()
traffic()
if os.path.exists(os.path.join(SRC, 'layers.txt')):
    shutil.copy(os.path.join(SRC, 'layers.txt'), os.path.join(DST, '%s_layers_b16.txt' % tag))
ops()
