run() {
  for i in 1 2 3 4 5 6 7 8; do
    env "$@" T2H_TRAIN_DEBUG=1 timeout 120 python -m pytest tests/test_gpu_vqgan_train.py -q -m gpu -s -x -k "fixture and fp32" 2>&1 | grep -E "train debug. .rg" | sed 's/.*dw \([0-9.]*\).*/\1/'
  done | tr '\n' ' '
  echo
}
echo "baseline:"; run A=1
echo "WS_ZEROS:"; run T2H_WS_ZEROS=1
echo "PDL=0:"; run T2H_PDL=0
