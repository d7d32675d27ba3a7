import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f, 'ms/step', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1),
                  'conv TF', round(d['roofline']['achieved'], 1), 'frac', round(d['roofline']['frac'], 3))
            print('  ', {k: (round(v['ms_per_step'], 2), round(v.get('GBps', v.get('TFLOPs', 0)), 0)) for k, v in d['kernel_breakdown'].items()})
            for t in d['roofline']['top_shapes']:
                print('    ', t['shape'], round(t['ms'], 3), 'ms', round(t['TFLOPs'], 0), 'TF/s')
            print('   clocks', d['clocks'])
