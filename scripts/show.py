"""prints the headline numbers of bench JSON files: python scripts/show.py gpurun_out/r01z_*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        txt = [l for l in open(f) if l.startswith('{')][-1]
        d = json.loads(txt)
        lp = d.get('lookup_pass') or {}
        e2e = d.get('e2e') or {}
        print('%-44s %6.2f G/s %7.1f ms' % (f.split('/')[-1], d['value'] / 1e9, d['ms_per_step']),
              {k: round(v, 1) for k, v in d.get('profile_ms_per_step', {}).items()},
              'grp', d['config'].get('table_partitions'), 'lookup %.1f G/s' % (lp.get('value', 0) / 1e9), 'e2e %.1f G/s' % (e2e.get('value', 0) / 1e9))
    except Exception as e:
        print(f, 'ERR', e)
