for v in head tmpl head tmpl; do
  export DTR_B200_LIB=/root/repo/variants/libdtr_$v.so
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -1
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v mesh1080', d['roofline']['ms_per_launch'], d['ms_per_step'])"
  python bench.py --workload views1080_tex --views 64 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v views1080_tex64', d['roofline']['ms_per_launch'], d['ms_per_step'])"
  python bench.py --workload mesh4k_tex --views 16 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v mesh4k_tex', d['roofline']['ms_per_launch'], d['ms_per_step'])"
done
