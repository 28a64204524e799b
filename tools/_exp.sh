export GAUCHE_B200_EW=16
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for e in mxf4x1 mxf4 umma2; do echo "== EW16 $e"; timeout 120 python tools/profile_umma.py --engine $e --rows 16384 --train 100000 --launches 5 | tail -1; done
for m in 12 8; do echo "EW16 mxf4x1 mode $m"; GAUCHE_B200_DEBUG_MODE=$m timeout 120 python tools/profile_umma.py --engine mxf4x1 --rows 16384 --train 100000 --launches 4 | tail -1; done
timeout 300 python tools/bench_fused.py 2>&1 | tail -2
unset GAUCHE_B200_EW
timeout 300 python tools/bench_fused.py 2>&1 | tail -1
