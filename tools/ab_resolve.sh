# A/B of resolve-kernel variants: raster stage time of 64 textured 1080p views (parity checked by bench.py)
cd /root/repo
export WL="${WL:-views1080_tex}"
bash tools/ab_variants.sh "$@" 2>&1
for c in 8 16 64 128; do
  echo "== DTR_B200_RESOLVE_CTAS=$c with $1"
  DTR_B200_RESOLVE_CTAS=$c bash tools/ab_variants.sh $1 2>&1 | grep -v "^base"
done
