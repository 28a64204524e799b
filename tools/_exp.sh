python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('value',d['value'],'e2e',e['value'],'packed',e['packed_wire_variant']['value'],'u8',e['u8_host_variant'])"
timeout 120 python tools/umma_debug.py umma | tail -1
timeout 120 python tools/profile_umma.py --engine umma --rows 16384 --train 100000 --launches 4 | tail -1
