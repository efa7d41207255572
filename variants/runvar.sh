for v in base expens base expens; do
  export DTR_B200_LIB=/root/repo/variants/libdtr_$v.so
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v mesh1080', d['roofline']['ms_per_launch'], d['ms_per_step'])"
done
