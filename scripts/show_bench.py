import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, 'ms/step', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1),
                  'conv TF', round(d['roofline']['achieved'], 1), 'frac', round(d['roofline']['frac'], 3))
            print('  ', {k: (round(v['ms_per_step'], 2), round(v.get('GBps', v.get('TFLOPs', 0)), 0)) for k, v in d['kernel_breakdown'].items() if isinstance(v, dict)})
            for t in d['roofline']['top_shapes']:
                print('    ', t['shape'], round(t['ms'], 3), 'ms', round(t['TFLOPs'], 0), 'TF/s')
            print('   clocks', d['clocks'])
            sm = d.get('sampling')
            if sm:
                print('   sampling', round(sm['value'], 3), 'volumes/s', round(sm['ms_per_batch'], 1), 'ms; e2e', round(sm['e2e']['value'], 3), 'frac_executed', round(sm.get('frac_executed', 0), 3))
            oc = d.get('other_configs')
            if oc:
                c1 = oc.get('cfg1_default_inference')
                if c1:
                    print('   cfg1', round(c1['ms_per_volume'], 1), 'ms per volume e2e;', 'cpu', c1.get('cpu_baseline', {}).get('seconds_per_volume'), 's')
                dd = oc['ddim50_40']
                print('   ddim50_40', round(dd['value'], 3), 'volumes/s', round(dd['ms_per_batch'], 1), 'ms; e2e', round(dd['e2e']['value'], 3), 'frac_executed', round(dd['frac_executed'], 3))
                for r in oc['ae_sweep']['sizes']:
                    print('   ae', r['edge'], round(r['ms'], 2), 'ms', round(r['tflops_executed'], 0), 'TF/s executed; e2e', round(r['e2e_ms'], 2), 'ms')
