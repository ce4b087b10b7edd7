"""The L2 read-bandwidth probe behind `roofline_l2.peak` (ptb_measure_l2 -> k_l2_read): 148 x 8 blocks x 256 threads stream a 64 MiB
buffer `iters` times with 16-byte ld.global.cg loads after two warm passes.  Run plain for the number, under ncu for the evidence that
the bytes come from L2 (lts__t_sector_hit_rate ~ 100 %, dram__bytes_read ~ 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptina_b200 import _native, worker
worker.init()
ctx = _native.context()
for mb, it in ((64, 20), (32, 40), (96, 14)):
    print(f'{mb} MiB x {it}: {ctx.measure_l2(mb, it):.0f} GB/s')
