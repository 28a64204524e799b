python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for e in mxf4x1 mxf4 umma2 mxf4mc; do echo "== $e"; timeout 120 python tools/profile_umma.py --engine $e --rows 16384 --train 100000 --launches 6 | tail -1; done
for m in 8 12; do echo "mxf4x1 mode $m"; GAUCHE_B200_DEBUG_MODE=$m timeout 120 python tools/profile_umma.py --engine mxf4x1 --rows 16384 --train 100000 --launches 4 | tail -1; done
timeout 300 python tools/bench_fused.py 2>&1 | tail -2
