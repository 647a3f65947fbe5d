#!/bin/bash
# One full ncu capture of the pruned max-min kernel on a mid-lattice slice (run under gpurun):
#   gpurun --timeout 600 -- 'bash dev/ncu_bp.sh r02i [cfg3] [count] [i0]'
tag=${1:-rXX}; cfg=${2:-cfg3}; cnt=${3:-4194304}; i0=${4:-50000000}
out=gpurun_out; mkdir -p $out
python dev/prof_run.py pruned $cnt $cfg $i0 > $out/${tag}_prof_run.log 2>&1 || { tail -5 $out/${tag}_prof_run.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_maxmin_bp -c 1 -s 1 -f -o $out/${tag}_prof_bp \
    python dev/prof_run.py pruned $cnt $cfg $i0 > $out/${tag}_ncu_bp.log 2>&1
ncu -i $out/${tag}_prof_bp.ncu-rep --page details > $out/${tag}_ncu_details_k_maxmin_bp.txt 2>&1
ncu -i $out/${tag}_prof_bp.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_k_maxmin_bp.csv 2>&1
ncu -i $out/${tag}_prof_bp.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $out/${tag}_ncu_source_k_maxmin_bp.csv.gz
[ $(stat -c %s $out/${tag}_prof_bp.ncu-rep) -gt 20000000 ] && rm -f $out/${tag}_prof_bp.ncu-rep
cat $out/${tag}_prof_run.log
