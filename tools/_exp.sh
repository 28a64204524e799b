python -m pytest tests -m gpu -q -x -k "fused or screen or rowdot or gloo or variant" 2>&1 | tail -2
timeout 300 python tools/bench_fused.py 2>&1 | tail -3
python bench.py --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('value',d['value'],'e2e',e['value'],'packed',e['packed_wire_variant']['value'],'u8',e['u8_host_variant']['value'],'parity',d['cpu_baseline']['parity_vs_gpu_arm'])"
