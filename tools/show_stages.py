import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value %.0f ms/step %.3f'%(d['value'], d['ms_per_step']), {k:round(v['ms'],3) for k,v in d['roofline']['stages'].items()}, 'frac %.3f'%d['roofline']['frac'])
