for d in 0 1 2 3; do
  PA_CONV_DEBUG=$d python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('debug $d', 'step %.3f' % d['ms_per_step'], ' '.join('%s=%.3f' % (n.split(':')[-1], v['ms_per_step']) for n,v in k.items() if 'layer1' in n or 'layer2.1' in n or 'layer3.1.conv2' in n))"
done
