for v in s10t25 s20t25 s10t40 s5t10 s30t40; do
  export DTR_B200_LIB=/root/repo/variants/libdtr_$v.so
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v mesh1080', d['roofline']['ms_per_launch'], d['ms_per_step'])"
  python bench.py --workload views1080_tex --views 16 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v views1080_tex', d['roofline']['ms_per_launch'], d['ms_per_step'])"
done
