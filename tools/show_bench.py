import json, sys
d = json.load(open(sys.argv[1]))
print('steps/s %.2f  ms/step %.2f  e2e %.2f  kernels/step %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d.get('kernels_per_step')))
r = d.get('roofline')
if r:
    print('TC kernel: %.1f TFLOP/s (%.1f%% of %s), %.2f ms/step in kernel, %.1f GFLOP/step' % (r['achieved'], 100 * r['frac'], r['peak'], r['ms_per_step_in_kernel'], r['algorithmic_gflop_per_step']))
for l in d.get('roofline_per_layer', []):
    print('  M=%7d K=%5d N=%4d taps=%2d x%d: %7.1f us %7.1f TF' % (l['M'], l['K'], l['N'], l['taps'], l['launches_per_step'], l['us'], l['tflops']))
print('clocks', d.get('clocks'))
print('cpu_baseline', d.get('cpu_baseline'))
print('library_baseline', d.get('library_baseline'))
print('e2e', d.get('e2e'))
