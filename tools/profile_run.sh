#!/bin/bash
# Round-2 profiling pass on ONE GPU box (run under gpurun): launch list of the bench command + `ncu --set full` captures of the
# dominant kernels. Numbers printed by runs under ncu are never bench values. Raw reports go to gpurun_out/ (scratch); the
# summaries are extracted here with tools/ncu_extract.py and copied into profiles/ by hand.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ntt --no-sizes --no-prove"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_msm_bench.csv $B > $O/r02_launches_bench.log 2>&1
for k in msm_accumulate msm_group msm_stitch msm_scatter; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o $O/r02_prof_$k $B > $O/r02_prof_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 4 -c 2 -f -o $O/r02_prof_ntt22 python tools/microbench.py ntt k=22 > $O/r02_prof_ntt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:permutation_constraints\|graph_evaluate\|lookup_constraints -c 3 -f -o $O/r02_prof_quotient python tools/quotient_probe.py > $O/r02_prof_quotient.log 2>&1
for f in $O/r02_prof_*.ncu-rep; do
  ncu -i $f --page raw --csv 2>/dev/null | python tools/ncu_extract.py > ${f%.ncu-rep}.md 2>/dev/null
done
ls -la $O/r02_prof_* | head -30
