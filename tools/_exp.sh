python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>&1 | tail -1 > gpurun_out/bench_mxf4x1_v3.json; cat gpurun_out/bench_mxf4x1_v3.json | cut -c1-400
python bench.py --engine umma2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_umma2_v3.json; cut -c1-300 gpurun_out/bench_umma2_v3.json
python tools/bench_configs.py 2>&1 | tail -12
python tools/bench_minmax.py 2>&1 | tail -6
