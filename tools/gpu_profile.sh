# ncu captures behind profiles/: launch list of one short bench run + `--set full` of the dominant kernels, summarised on the box
# (the .ncu-rep files are deleted there: gpurun copies back at most 64 MiB).
# usage (from this container):  gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh <tag>'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-x}
B="python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv | tee gpurun_out/launch_shares_$TAG.txt
summarise() {   # <rep stem> <number of launches for the hot-line listing>
    python tools/ncu_metrics.py gpurun_out/$1.ncu-rep > gpurun_out/$1.txt 2>&1
    for i in $(seq 0 $(($2 - 1))); do python tools/ncu_lines.py gpurun_out/$1.ncu-rep $i 25 2>&1 | cut -c1-260 > gpurun_out/$1.lines$i.txt; done
    rm -f gpurun_out/$1.ncu-rep
}
# the two fused MRF kernels (a full chunk of the timed pass), then the decoder's k_conv_tc launches of the same pass, then attention
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mrf3' -s ${MRF_SKIP:-6} -c 2 -o gpurun_out/mrf_full_$TAG -f $B > gpurun_out/ncu_mrf_$TAG.log 2>&1
summarise mrf_full_$TAG 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc' -s ${CONV_SKIP:-75} -c 9 -o gpurun_out/conv_dec_$TAG -f $B > gpurun_out/ncu_conv_$TAG.log 2>&1
summarise conv_dec_$TAG 0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rel_attention' -s 2 -c 1 -o gpurun_out/attn_$TAG -f $B > gpurun_out/ncu_attn_$TAG.log 2>&1
summarise attn_$TAG 1
ls -la gpurun_out | grep $TAG
