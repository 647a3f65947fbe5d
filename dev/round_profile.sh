#!/bin/bash
# Round measurement pass on one B200 (run under gpurun): GPU tests, bench lines, launch list, one full ncu capture
# of the dominant kernel, kernel sweep.  Outputs go to gpurun_out/; copy the summaries to profiles/.
#   gpurun --timeout 1500 -- 'bash dev/round_profile.sh r02h [notest]'
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
if [ "$2" != notest ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gputests.log
fi
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench_n1.err
python bench.py --steps 2 --warmup 3 --kernel exact > $out/${tag}_bench_n1_exact_kernel.json 2>> $out/${tag}_bench_n1.err
python dev/perf_sweep.py cfg2 cfg3 cfg4 > $out/${tag}_kernel_sweep.log 2>&1
python dev/perf_sweep.py cfg5 exact pruned >> $out/${tag}_kernel_sweep.log 2>&1
# launch list of the bench command (serialised, cold cache: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu_bench.log 2>&1
# full capture of the dominant kernel on a mid-lattice slice of cfg3
ncu --set full --clock-control none --import-source on -k regex:k_maxmin_bp -c 1 -s 1 -f -o $out/${tag}_prof_bp \
    python dev/prof_run.py pruned 4194304 cfg3 50000000 > $out/${tag}_ncu_bp.log 2>&1
ncu -i $out/${tag}_prof_bp.ncu-rep --page details > $out/${tag}_ncu_details_k_maxmin_bp.txt 2>&1
ncu -i $out/${tag}_prof_bp.ncu-rep --page raw --csv > $out/${tag}_ncu_raw_k_maxmin_bp.csv 2>&1
ncu -i $out/${tag}_prof_bp.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $out/${tag}_ncu_source_k_maxmin_bp.csv.gz
# gpurun copies back at most 64 MiB: the report itself stays on the box when it is large
[ $(stat -c %s $out/${tag}_prof_bp.ncu-rep) -gt 20000000 ] && rm -f $out/${tag}_prof_bp.ncu-rep
du -sh $out
tail -1 $out/${tag}_bench_n1.json | cut -c1-400
tail -3 $out/${tag}_gputests.log
