# like gpu_variants.sh, but runs the whole GPU test tier against each build
for v in "$@"; do
  export MDOPT_B200_LIB=$PWD/build/variants/lib$v.so
  timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
done
bash scripts/gpu_variants.sh "$@"
