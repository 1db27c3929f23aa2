#!/usr/bin/env python
"""Turn the ncu artefacts of a gpurun trip (gpurun_out/) into the committed evidence under profiles/:
   r2_ncu_full_<kernel>.txt (tools/ncu_summary.py), r2_launches_<mode>_summary.txt (tools/launch_summary.py),
   and the `train` / `kan` / `infer` entries of profiles/roofline_traffic.json (DRAM bytes per launch from the --set full captures)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r2'


def dram_bytes(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        tot = 0.0
        for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            tot += float(r[col[key]].replace(',', '')) * scale[units[col[key]]]
        dur = float(r[col['gpu__time_duration.sum']].replace(',', ''))
        dur *= {'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3}.get(units[col['gpu__time_duration.sum']], 1.0)
        out.append((r[col['Kernel Name']], tot, dur))
    return out


def main():
    names = {}
    for f in sorted(os.listdir(OUT)):
        if f.startswith('prof_') and f.endswith('.ncu-rep'):
            k = f[len('prof_'):-len('.ncu-rep')]
            rep = os.path.join(OUT, f)
            txt = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
            if txt.strip():
                open(os.path.join(PROF, f'{TAG}_ncu_full_{k}.txt'), 'w').write(txt)
                names[k] = dram_bytes(rep)
                print(k, [(n[:40], round(b / 1e6, 1), round(d, 1)) for n, b, d in names[k]])
    for mode in ('train', 'kan', 'infer'):
        src = os.path.join(OUT, f'launches_{mode}.csv')
        if os.path.exists(src):
            txt = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'launch_summary.py'), src], capture_output=True, text=True).stdout
            head = {'train': 'ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline\n'
                             '(RoViT-KAN stage-4 train steps, batch 256, B200; cold-cache serialised launch times: compare SHARES)\n',
                    'kan': 'ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --mode kan --steps 2 --warmup 3 --no-cpu-baseline\n'
                           '(KANSeverityModule [192,64,1] and [192,64,16,1], batch 65536, forward and forward+backward; cold-cache serialised)\n',
                    'infer': 'ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 100 : python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline\n'
                             '(RoViT-KAN inference forwards, batch 1024, B200; cold-cache serialised launch times: compare SHARES)\n'}[mode]
            open(os.path.join(PROF, f'{TAG}_launches_{mode}_summary.txt'), 'w').write(head + txt)
    tj_path = os.path.join(PROF, 'roofline_traffic.json')
    tj = json.load(open(tj_path))
    if 'gemm_tn' in names and names['gemm_tn']:
        n, b, d = names['gemm_tn'][0]
        tj['train'] = {'batch_per_gpu': 256, 'kernel': 'gemm_tn_kernel<192,4>', 'dram_bytes_per_launch': b,
                       'captured_launch_us': d,
                       'source': f'profiles/{TAG}_ncu_full_gemm_tn.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the one '
                                 'captured launch of the batch-256 train step; the 49 launches per step differ in shape)'}
    kan_parts = ['kan_fwd_tc', 'kan_small_fwd', 'kan_small_bwd', 'kan_bwd_x_tc', 'kan_bwd_w_tc']
    if all(k in names and names[k] for k in kan_parts):
        tot = sum(names[k][0][1] for k in kan_parts)
        tj['kan'] = {'batch_per_gpu': 65536, 'kernel': 'KANSeverityModule([192,64,1]) fwd+bwd: ' + ' + '.join(kan_parts),
                     'dram_bytes_per_launch': tot, 'per_kernel_bytes': {k: names[k][0][1] for k in kan_parts},
                     'algorithmic_bytes_per_launch': 152.0e6,
                     'source': f'profiles/{TAG}_ncu_full_kan_*.txt (ncu --set full, dram read + write of one launch of each of the five kernels '
                               'of one forward+backward at batch 65536; "launch" = one fwd+bwd of the stack)'}
    mlp_key = 'mlp2' if names.get('mlp2') else 'mlp'       # mlp2 = the two-tile kernel the inference path runs by default
    if names.get(mlp_key):
        n, b, d = names[mlp_key][0]
        tj['infer'] = {'batch_per_gpu': 1024, 'kernel': n.split('(')[0].replace('void ', ''), 'dram_bytes_per_launch': b,
                       'algorithmic_bytes_per_launch': 465444864,
                       'source': f'profiles/{TAG}_ncu_full_{mlp_key}.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch at batch 1024)'}
    json.dump(tj, open(tj_path, 'w'), indent=1)
    print(json.dumps(tj, indent=1)[:1500])


if __name__ == '__main__':
    main()
