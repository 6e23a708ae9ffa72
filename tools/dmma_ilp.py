"""Print the DMMA.8x8x4 issue rate (TFLOP/s, whole chip) versus accumulators per warp and warps per SM."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp

lib = gp._lib.load()
gp._lib.require_cuda()
print('%8s' % 'acc\\warps' + ''.join('%9d' % w for w in (4, 8, 16, 32)))
for nacc in (1, 2, 4, 8, 16, 32):
    row = []
    for w in (4, 8, 16, 32):
        if nacc * w > 512:          # would exceed the register file (64 accumulator registers x 1024 threads)
            row.append(float('nan'))
            continue
        tf = ctypes.c_double()
        gp._lib.check(lib.gpmc_bench_dmma_ilp(nacc, w, 20000 // nacc, ctypes.byref(tf)), 'dmma_ilp')
        row.append(tf.value)
    print('%8d ' % nacc + ''.join('%9.2f' % v for v in row))
