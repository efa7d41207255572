# first a small, time-limited correctness run of the warp-pair kernel (a hang must not eat the box), then A/B
for v in "$@"; do
DTR_B200_LIB=/root/repo/variants/libdtr_$v.so timeout 300 python -X faulthandler -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "pytest $v rc=$?"
done
WL="views1080_tex mesh1080 fill4k" bash tools/ab_variants.sh "$@"
