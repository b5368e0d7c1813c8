#!/usr/bin/env python3
"""Print the key figures of bench.py JSON lines: python tools/bench_summary.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, 'unreadable:', e)
        continue
    e2e = d.get('e2e', {})
    print('{}: n_gpus {} value {:.0f} e2e {:.0f} (predict_f32 {}, n256_f64 {}) frac {} sustained {} scaling {}'.format(
        path, d.get('n_gpus'), d['value'], e2e.get('value', 0),
        '{:.0f}'.format(e2e['predict_f32']['value']) if e2e.get('predict_f32') else None,
        '{:.0f}'.format(e2e['predict_n256_f64']) if e2e.get('predict_n256_f64') else None,
        d.get('roofline', {}).get('frac'), d.get('roofline', {}).get('frac_sustained'), d.get('scaling')))
    cfg = (d.get('config') or {}).get('configs') or {}
    for k, v in cfg.items():
        if isinstance(v, dict):
            print('   {:28s} reads/s {:.0f} windows/s {:.0f} of-kernel {:.3f} cpu {}'.format(
                k, v.get('reads_per_s', 0), v.get('windows_per_s', 0), v.get('fraction_of_kernel_rate', 0),
                v.get('cpu_port_reads_per_s')))
        else:
            print('  ', k, v)
    if d.get('parity'):
        print('   parity', d['parity']['max_abs_err'], 'unsaturated', d['parity']['unsaturated'], '| clocks', d.get('clocks'))
    if d.get('cpu_baseline'):
        print('   cpu_baseline', d['cpu_baseline']['value'], 'cores', d['cpu_baseline']['cores'])
