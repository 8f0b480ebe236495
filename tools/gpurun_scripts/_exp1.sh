set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv
echo "== default routing"; python tools/bench_sustained.py 2 0 1
echo "== NO_T (pixels-as-M single CTA BN=128)"; D3_CONV_NO_T=1 python tools/bench_sustained.py 2 0 1
echo "== NO_T + 2CTA"; D3_CONV_NO_T=1 D3_CONV_2CTA=1 python tools/bench_sustained.py 2 0 1 2
echo "== NO_T + 2CTA, no epilogue"; D3_CONV_DEBUG_SKIP=4 D3_CONV_NO_T=1 D3_CONV_2CTA=1 python tools/bench_sustained.py 2 0 1 2
echo "== NO_T + 2CTA, MMA only"; D3_CONV_DEBUG_SKIP=7 D3_CONV_NO_T=1 D3_CONV_2CTA=1 python tools/bench_sustained.py 2 0 1 2
echo "== NO_T + 2CTA, no loads"; D3_CONV_DEBUG_SKIP=3 D3_CONV_NO_T=1 D3_CONV_2CTA=1 python tools/bench_sustained.py 2 0 1 2
echo "== default, no epilogue"; D3_CONV_DEBUG_SKIP=4 python tools/bench_sustained.py 2 0 1
echo "== fp16 bench"; python bench.py --dtype fp16 --steps 5 --no-cpu-baseline
echo "== bf16 bench"; python bench.py --dtype bf16 --steps 5 --no-cpu-baseline
