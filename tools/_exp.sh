python -m pytest tests -m gpu -q -x -k "host_blocks or fused or screen" 2>&1 | tail -3
for rep in 1 2; do python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/b3.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/b3.json').read()); e=d['e2e']
print('e2e %.0f (%.2f ms) pack_ms %.2f | dense %.0f packed %.0f u8 %.0f'%(e['value'],e['ms_per_step'],e['host_pack']['pack_ms'],e['dense_h2d_variant']['value'],e['packed_wire_variant']['value'],e['u8_host_variant']['value']))"; done
