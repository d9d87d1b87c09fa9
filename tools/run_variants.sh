for v in minb2 minb3 minb4 strict; do
  echo "=== $v"
  MPASB_LIB=$PWD/mpas_model_b200/csrc/libmpasb_$v.so timeout 300 python tools/quick_bench.py 40962 55 10 2>&1 | grep -E "ms/step|k:k2" | head -5
done
MPASB_LIB=$PWD/mpas_model_b200/csrc/libmpasb_minb2.so timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -s 2>&1 | tail -12
MPASB_LIB=$PWD/mpas_model_b200/csrc/libmpasb_strict.so timeout 600 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
