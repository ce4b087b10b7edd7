#!/bin/bash
# parameter sweep (tuning builds under variants/): config 2 and config 4, headline only
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-10s %-14s Mrays/s %7.1f  ms/step %7.3f  stages %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}))'
python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "benchtiles or validation or benchmark_script" 2>&1 | tail -5
run() { PTINA_B200_LIB=$2 python bench.py --quick --no-cpu --scene $1 --steps $3 --warmup $4 2>/dev/null | tail -1 | python -c "$summ" $1 $5; }
run cornell_monkey $PWD/ptina_b200/libptina_b200.so 10 3 default
for v in vote_l1 vote_l3 fetch4 fetch16 sblk256 sblk1024; do run cornell_monkey $PWD/variants/$v.so 10 3 $v; done
run cornell_monkey $PWD/ptina_b200/libptina_b200.so 10 3 default-again
run mega $PWD/ptina_b200/libptina_b200.so 3 1 default
for v in vote_l1 vote_l3 fetch4 fetch16 minb6 minb8; do run mega $PWD/variants/$v.so 3 1 $v; done
run matball $PWD/ptina_b200/libptina_b200.so 5 2 default
for v in vote_l1 vote_l3 fetch4 fetch16; do run matball $PWD/variants/$v.so 5 2 $v; done
