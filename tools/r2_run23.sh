#!/bin/bash
# small wavefronts: resident vs global-memory tree kernel (MLT, and config 2 at 4 / 1 spp per step)
for env in "" "PTB_NO_RESIDENT_BVH=1" "PTB_NO_RESIDENT_BVH=1 PTB_NO_WIDE4=1" "PTB_QUANT_RESIDENT_BVH=1"; do
  echo "== ${env:-default}"
  env $env python tools/mlt_bench.py 64 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mlt ms/render %.3f  Mprop/s %.1f' % (d['ms_per_render'], d['proposals_per_s']/1e6))"
  for spp in 4 1; do env $env python bench.py --quick --no-cpu --scene cornell_monkey --spp $spp 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 spp $spp: Mrays/s %.1f ms/step %.3f' % (d['value'], d['ms_per_step']), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})"; done
done
