for v in h32c5 h16c6 h16c7 h16c8; do
  export DTR_B200_LIB=/root/repo/variants/libdtr_$v.so
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  python bench.py --steps 50 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['roofline']['ms_per_launch'], d['ms_per_step'])"
done
