for e in mxf4x1 mxf4 umma2 umma2x1 mxf4mc mxf4r2 umma2r2; do timeout 120 python tools/umma_debug.py $e 2>&1 | tail -1; done
for e in mxf4x1 mxf4 umma2 umma2x1; do echo "== $e"; timeout 120 python tools/profile_umma.py --engine $e --rows 16384 --train 100000 --launches 5; done
for m in 1 4 8 9 12; do echo "mxf4x1 mode $m"; GAUCHE_B200_DEBUG_MODE=$m timeout 120 python tools/profile_umma.py --engine mxf4x1 --rows 16384 --train 100000 --launches 4; done
timeout 300 python tools/bench_fused.py 2>&1 | tail -6
