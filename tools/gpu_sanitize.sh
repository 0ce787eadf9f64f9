# compute-sanitizer passes over the final build: gpu_sanitize.sh <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1
OUT=gpurun_out/sanitizer_$TAG.txt
CS=/usr/local/cuda/bin/compute-sanitizer
{
echo "compute-sanitizer (CUDA 12.9) on a B200, final build of round 2 (TMA loaders, programmatic dependent launch, monotonic alignment search):"
echo "memcheck, smoke() (x_low, fp32 + bf16, + 3 alignments):"
timeout 900 $CS --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke\[|ERROR SUMMARY|Invalid|error" | head -12
echo "memcheck, GPU parity tests: golden fixtures, TMA-fed tiles (medium), fused vs unfused stage, summed second convs:"
timeout 1500 $CS --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "golden or (tma_fed and medium) or fused or summed" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
echo "memcheck, monotonic alignment search tests:"
timeout 900 $CS --tool memcheck python -m pytest tests/test_gpu_mas.py -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
echo "synccheck, smoke():"
timeout 900 $CS --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|error" | head -4
echo "initcheck, smoke():"
timeout 900 $CS --tool initcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|error" | head -4
} > $OUT 2>&1
cat $OUT
