"""ncu CSV logs (gpurun_out/) -> the summaries committed under profiles/.

    python tools/summarise_ncu.py launches gpurun_out/launches.csv profiles/r1c_launches_summary.txt "<header note>"
    python tools/summarise_ncu.py traffic gpurun_out/conv1x1.csv gpurun_out/conv3x3.csv profiles/conv_traffic.json "<source note>"
"""
import csv
import json
import sys
from collections import defaultdict


def rows_of(path):
    lines = [l for l in open(path, errors='replace') if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if 'Kernel Name' in r:
            hdr = r
            break
    iN, iM, iV = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
    iI = hdr.index('ID')
    out = defaultdict(dict)
    names = {}
    for r in rd:
        if len(r) <= iV:
            continue
        names[r[iI]] = r[iN]
        out[r[iI]][r[iM]] = float(r[iV].replace(',', ''))
    return names, out


def launches(path, dst, note):
    names, vals = rows_of(path)
    agg = defaultdict(lambda: [0, 0.0])
    for i, n in names.items():
        short = n.split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        agg[short][0] += 1
        agg[short][1] += vals[i].get('gpu__time_duration.sum', 0.0) / 1e3      # ns -> us
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        f.write('# %s\n# %d launches, total %.1f us (cold-cache, serialised: compare SHARES, not absolutes)\n' % (note, len(names), tot))
        f.write('%-46s %5s %12s %7s %9s\n' % ('kernel', 'n', 'total_us', 'share', 'avg_us'))
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-46s %5d %12.1f %6.1f%% %9.1f\n' % (k[:46], n, us, 100 * us / tot, us / n))
    print(open(dst).read())


def traffic(p1, p3, dst, note):
    out = {}
    for key, path in (('1x1', p1), ('3x3', p3)):
        names, vals = rows_of(path)
        n = len(names)
        rd = sum(v.get('dram__bytes_read.sum', 0.0) for v in vals.values())
        wr = sum(v.get('dram__bytes_write.sum', 0.0) for v in vals.values())
        us = sum(v.get('gpu__time_duration.sum', 0.0) for v in vals.values()) / 1e3
        out[key] = dict(launches=n, dram_bytes_per_launch=(rd + wr) / max(n, 1), dram_read_bytes=rd, dram_write_bytes=wr,
                        total_us_under_ncu=us, source=note)
    json.dump(out, open(dst, 'w'), indent=1)
    print(json.dumps(out, indent=1))


FAMILY_OF_TAG = {'sh_conv_fwd_1x1': 'conv1x1', 'sh_conv_fwd_3x3': 'conv3x3', 'sh_conv_wgrad_1x1': 'wgrad1x1', 'sh_conv_wgrad_3x3': 'wgrad3x3',
                 'sh_conv_wgrad3x3': 'wgrad3x3', 'sh_gn_relu_fwd': 'gn_relu_fwd', 'sh_gn_relu_bwd_prezeroed': 'gn_relu_bwd', 'sh_gn_relu_bwd': 'gn_relu_bwd',
                 'sh_maxpool_fwd': 'pool_upsample_add', 'sh_maxpool_bwd': 'pool_upsample_add', 'sh_upsample_add_fwd': 'pool_upsample_add',
                 'sh_upsample_bwd': 'pool_upsample_add', 'sh_add': 'pool_upsample_add', 'sh_mvproj_loss_fwdbwd': 'mvproj'}


def families(path, dst, note):
    """Launch list taken with `--nvtx --print-nvtx-rename kernel` (kernels carry the name of the C-ABI call that launched them) and the
    metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum -> per-family DRAM traffic per launch, the file
    bench.py reads for `roofline.traffic` (keys = bench.FAMILY_KERNELS)."""
    names, vals = rows_of(path)
    out = {}
    for i, n in names.items():
        tag = n.split('/')[0].strip()               # "<NVTX range = C-ABI call>/<kernel>"
        fam = FAMILY_OF_TAG.get(tag)
        if fam is None:
            continue
        d = out.setdefault(fam, dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, total_us_under_ncu=0.0))
        if tag == 'sh_mvproj_loss_fwdbwd' and 'mvproj_main_kernel' not in n:
            pass                                  # prep / finish kernels of the call: their bytes count, not their launch
        else:
            d['launches'] += 1
        d['dram_read_bytes'] += vals[i].get('dram__bytes_read.sum', 0.0)
        d['dram_write_bytes'] += vals[i].get('dram__bytes_write.sum', 0.0)
        d['total_us_under_ncu'] += vals[i].get('gpu__time_duration.sum', 0.0) / 1e3
    for fam, d in out.items():
        d['dram_bytes_per_launch'] = (d['dram_read_bytes'] + d['dram_write_bytes']) / max(d['launches'], 1)
        d['source'] = note
    json.dump(out, open(dst, 'w'), indent=1)
    print(json.dumps(out, indent=1))


def reps(dst, pairs):
    """`ncu --set full` reports (.ncu-rep, read with `ncu -i ... --page raw --csv`) -> one CSV row per profiled launch."""
    import subprocess
    cols = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'smsp__inst_executed.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct']
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['report', 'what'] + cols)
        for path, what in pairs:
            out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
            rows = list(csv.reader(out.splitlines()))
            hdr, units = rows[0], rows[1]
            unit = dict(zip(hdr, units))
            for r in rows[2:]:
                rec = dict(zip(hdr, r))
                w.writerow([path.split('/')[-1], what] + [(rec.get(c, '')[:70] + (' ' + unit[c] if unit.get(c) and c != 'Kernel Name' else '')).strip()
                                                           for c in cols])


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(*sys.argv[2:5])
    elif sys.argv[1] == 'families':
        families(*sys.argv[2:5])
    elif sys.argv[1] == 'reps':
        reps(sys.argv[2], [tuple(a.split('=', 1)) for a in sys.argv[3:]])
    else:
        traffic(*sys.argv[2:6])
