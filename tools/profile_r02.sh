#!/bin/bash
# Round-2 evidence run (one GPU): timing probes, the bench launch list and one `ncu --set full` capture per kernel family.
# Run under gpurun; outputs land in gpurun_out/ and are summarised here afterwards (tools/ncu_summary.py) into profiles/.
set -u
O=gpurun_out
T="timeout 600"
$T python tools/angle_sweep.py 17 33 51 63 64 66 91 96 101 127 128 129 160 192 199 200 201 202 224 255 256 320 384 511 512 > $O/r02_angle_sweep.log 2>&1
$T python tools/fast_ab.py 33 64 91 100 128 200 256 512 > $O/r02_fastmath_ab.log 2>&1
K2_N=4e7 $T python tools/k2_time.py 91 256 512 > $O/r02_k2_time.log 2>&1
$T python tools/bench_rows.py > $O/r02_rows_bench.jsonl 2> $O/r02_rows_bench.err
# launch list of the bench command (plain run first: ncu only after the same command exited 0 without it)
$T python bench.py --quick --no-cpu-baseline --steps 20 > $O/plain_bench.log 2>&1 && \
  $T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv \
     python bench.py --quick --no-cpu-baseline --steps 20 > $O/ncu_launches.log 2>&1
for k in k1u k1q k2s k2s512 k3 k4; do
  case $k in k1u|k1q) pat=eval_uniform;; k2s|k2s512) pat=moments_kernel;; k3) pat=loglike_kernel;; k4) pat=latent_kernel;; esac
  tgt=$k; export K2_A=256
  if [ $k = k2s512 ]; then tgt=k2s; export K2_A=512; fi
  $T python tools/prof_target.py $tgt > $O/plain_$k.log 2>&1 && \
    $T ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -o $O/prof_r02_$k -f python tools/prof_target.py $tgt > $O/ncu_$k.log 2>&1
  tail -1 $O/ncu_$k.log
  # summarise on the box (the reports together exceed what gpurun copies back); keep the two headline reports
  python tools/ncu_summary.py $O/prof_r02_$k.ncu-rep --top 14 > $O/r02_${k}_ncu_summary.txt 2>&1
  python tools/ncu_lines.py $O/prof_r02_$k.ncu-rep --top 25 >> $O/r02_${k}_ncu_summary.txt 2>&1
  case $k in k1u|k2s) ;; *) rm -f $O/prof_r02_$k.ncu-rep;; esac
done
# the bench lines themselves (our arm with the CPU baseline, then the reference arm)
$T python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
$T python bench.py --impl reference > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
echo done
