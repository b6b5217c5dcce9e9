# A/B of prebuilt library variants (scripts/_build/lib_*.so) on the headline workload: device-resident value only
cp embree_b200/csrc/libembree4_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp scripts/_build/lib_$v.so embree_b200/csrc/libembree4_b200.so
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/var_$v.json').read().strip().splitlines()[-1]); print('$v', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms', 'nodes/ray', round(d['roofline']['nodes_per_ray'],2), 'tris/ray', round(d['roofline']['tris_per_ray'],2), 'parity', d.get('parity'))"
done
cp /tmp/lib_orig.so embree_b200/csrc/libembree4_b200.so
